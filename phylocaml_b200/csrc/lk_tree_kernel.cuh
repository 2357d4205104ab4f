// Tree-fused 4-state pruning: ONE launch evaluates the whole post-order schedule.
//
// Sites are independent, so a thread can carry its (pattern, rate class) through every node
// of the tree. The running CLV lives in registers; only the first child of a node with two
// interior children is parked on a small shared-memory stack (depth = Strahler number of the
// tree, <= log2 T); tips are 1 byte per pattern. Compared with one streaming kernel per node
// (3 CLV moves per update), HBM sees at most ONE write per update (RETAIN: every interior
// CLV is still produced in its slot for later incremental use) or nothing but the tips.
//
// Execution model (v5, after profiling v1..v4 -- see profiles/README.md):
//   * thread = one rate class of R=2 patterns; CTA = 128 threads = a tile of 256/K patterns;
//     four CTAs per SM. Two patterns per thread halve the transition-matrix traffic into
//     registers (the L1 write-back pipe was the limiter at R=1) and double the ILP;
//   * warps run FREE through the schedule: there is no CTA barrier per step. The only
//     barriers are one per tile (tip buffer hand-over) and one per 1024-pattern block
//     (canonical site-sum fold);
//   * tip masks of a tile ([T rows] x [TILE patterns], two 4-bit masks per byte) are staged by bulk-TMA
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), double buffered so the
//     next tile's tips land while this tile is computed;
//   * the transition matrices of a step (1 KB for K=4, k-interleaved so the K lanes of a
//     pattern read adjacent 16-byte chunks) are copied by each warp into its own
//     double-buffered shared-memory slot with cp.async one step ahead. (They were first read
//     through L1 with prefetch.global.L1 -- but with 4 CTAs x 52 KB the carve-out leaves no L1:
//     29 % of the sectors missed to L2 and the first use of each matrix stalled.)
//   * retained CLVs leave as 256-bit stores, 1 KB contiguous per warp.
//
// Arithmetic and its order are those of prune4_kernel / root4_kernel (=> same bits); a
// one-hot tip mask takes column j of P directly (0 + P[i][j] == P[i][j]).
#pragma once
#include "common.cuh"
#include "lk_kernels.cuh"  // tip_contrib

namespace phylo {

// One step of the compiled schedule: 32 bytes, read from shared memory one step ahead.
struct __align__(16) TreeInstr {
  int kinds;             // lkind | rkind << 2 | push_first << 4
  int lidx, ridx;        // tip row (OPK_TIP) or node slot (OPK_STORED)
  int out_slot;          // node slot that receives the result, or -1
  double *out_clv;       // RETAIN: where the result goes (NULL = nowhere)
  int32_t *out_sc;
};
static_assert(sizeof(TreeInstr) == 32, "TreeInstr is two 16-byte words");

constexpr int kTreeThreads = 128;
constexpr int kTreeR = 2;         // patterns per thread
constexpr int kTreePrefetch = 4;  // steps of look-ahead for the transition matrices

// Transition matrices are stored k-interleaved by pt_build_kernel(interleave=1): element
// e (= i*4+j) of rate class k sits at [e/2][k][e%2]. `pk` points at this thread's k.
// A "half" is two rows of the 4x4 matrix (8 doubles, 4 x LDG.128): h = 0 -> rows 0,1.
template <int K>
__device__ __forceinline__ void load_half(const double *__restrict__ pk, int h, double (&pm)[8]) {
#pragma unroll
  for (int ep = 0; ep < 4; ++ep) {
    const double2 t = *reinterpret_cast<const double2 *>(pk + (h * 4 + ep) * K * 2);
    pm[2 * ep] = t.x;
    pm[2 * ep + 1] = t.y;
  }
}
template <int K>
__device__ __forceinline__ void load_matrix(const double *__restrict__ pk, double (&pm)[16]) {
#pragma unroll
  for (int ep = 0; ep < 8; ++ep) {
    const double2 t = __ldg(reinterpret_cast<const double2 *>(pk + ep * K * 2));
    pm[2 * ep] = t.x;
    pm[2 * ep + 1] = t.y;
  }
}
// two rows: x[i] = sum_j P[i][j] v[j]  (same expression as prune4_kernel => same bits)
__device__ __forceinline__ void matvec_half(const double (&pm)[8], const d4 &v, double &x0, double &x1) {
  x0 = ((pm[0] * v.x + pm[1] * v.y) + pm[2] * v.z) + pm[3] * v.w;
  x1 = ((pm[4] * v.x + pm[5] * v.y) + pm[6] * v.z) + pm[7] * v.w;
}
__device__ __forceinline__ void matvec(const double (&pm)[16], const d4 &v, double (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    x[i] = ((pm[i * 4 + 0] * v.x + pm[i * 4 + 1] * v.y) + pm[i * 4 + 2] * v.z) + pm[i * 4 + 3] * v.w;
}
// a tip enters as a 0/1 vector: sum_j P[i][j]*L_j reproduces the ascending-j sum over the
// mask bit for bit. 1.0 = 0x3FF00000'00000000: built with integer ops, no I2F.
__device__ __forceinline__ d4 mask_vec(int m) {
  return d4{__hiloint2double((m & 1) * 0x3FF00000, 0), __hiloint2double(((m >> 1) & 1) * 0x3FF00000, 0),
            __hiloint2double(((m >> 2) & 1) * 0x3FF00000, 0), __hiloint2double(((m >> 3) & 1) * 0x3FF00000, 0)};
}

struct TreeArgs {
  const TreeInstr *prog;   // n_instr steps + 1 root step
  int n_instr;
  const double *P;         // [2*n_instr + 1][16*K] k-interleaved
  const uint8_t *tips4;    // [tip_stride/32][T][16]: groups of 32 patterns, two 4-bit masks per byte (even pattern = low nibble)
  int64_t tip_stride;      // patterns per row (multiple of 1024)
  int T;
  int64_t N;
  double *const *node_clv; // per slot: OPK_STORED operands
  int32_t *const *node_sc;
  const double *pi, *probs;
  double pinvar;
  const uint8_t *inv;
  const double *weights;
  double *site_lnl;
  double *groups;          // [ceil(N/32)] sums of 32 consecutive weighted site lnL (canonical level 0)
  int stack_depth;
  double2 *spill;          // warp-autonomous kernel: global scratch for deep stack levels
  int smem_levels, interleave;  // warp-autonomous kernel: stack levels in shared memory, group order
  int64_t tile_begin, tile_end;  // tiles this launch covers (a slab of the alignment)
};

template <int K, bool RETAIN>
__global__ void __launch_bounds__(kTreeThreads, 4) lk_tree4_kernel(const TreeArgs a) {
  constexpr int NT = kTreeThreads, R = kTreeR;
  constexpr int HALF = NT / K;              // patterns per r-slice of the tile
  constexpr int TILE = R * HALF;            // patterns per tile
  constexpr int GROUPS = TILE / 32;         // 32-pattern fold groups per tile
  constexpr int PM = 16 * K;                // doubles per transition matrix set (all k)
  constexpr int TROW = TILE / 2;            // bytes per tip row of a tile (nibble packed)
  static_assert(TILE % 32 == 0 && kLnlBlock % TILE == 0 && TROW % 16 == 0, "tile shape");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stage = reinterpret_cast<double *>(smem_raw);                      // [2][TILE] site values
  uint64_t *tipbar = reinterpret_cast<uint64_t *>(stage + 2 * TILE);         // [2] (+2 pad)
  d4 *stack = reinterpret_cast<d4 *>(tipbar + 4);                            // [depth][R][NT]
  int *stack_sc = reinterpret_cast<int *>(stack + (size_t)a.stack_depth * R * NT);
  int4 *sprog = reinterpret_cast<int4 *>(stack_sc + (size_t)a.stack_depth * R * NT);  // [n_steps][2]
  double *pring = reinterpret_cast<double *>(sprog + 2 * (size_t)(a.n_instr + 2));     // [warps][2][2*PM]
  uint8_t *tipbuf = reinterpret_cast<uint8_t *>(pring + (size_t)(NT / 32) * 2 * 2 * PM);
  tipbuf = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tipbuf) + 127) & ~(uintptr_t)127);
  const size_t tipbuf_bytes = (size_t)a.T * TROW;

  const int tid = threadIdx.x, k = tid % K, pl = tid / K, lane = tid & 31, warp = tid >> 5;
  const int n_steps = a.n_instr + 1;  // + root step
  const int64_t ntiles = (a.N + TILE - 1) / TILE;

  if (tid == 0) {
    mbar_init(&tipbar[0], 1);
    mbar_init(&tipbar[1], 1);
    fence_mbar_init();
  }
  // the compiled schedule lives in shared memory for the whole kernel
  for (int i = tid; i < 2 * n_steps; i += NT) sprog[i] = __ldg(reinterpret_cast<const int4 *>(a.prog) + i);
  if (tid < 2) sprog[2 * n_steps + tid] = make_int4(0, 0, 0, 0);  // harmless word past the end
  __syncthreads();

  const double pi0 = a.pi[0], pi1 = a.pi[1], pi2 = a.pi[2], pi3 = a.pi[3], pk_prob = a.probs[k];
  double *myring = pring + (size_t)warp * 2 * 2 * PM;  // this warp's two matrix slots
  // copy the two matrix sets of `step` (contiguous in a.P) into ring slot `slot`
  auto fetch_matrices = [&](int step, int slot) {
    const double *src = a.P + (size_t)(2 * step) * PM;
    double *dst = myring + (size_t)slot * 2 * PM;
    for (int c = lane; c < PM; c += 32) cp_async16(dst + 2 * c, src + 2 * c);
    cp_async_commit();
  };
  d4 *mystack = stack + tid;
  int *mystack_sc = stack_sc + tid;
  // nibble of pattern (pl + r*HALF) inside a tip row of the tile
  // (tips are group-major: [group of 32 patterns][T rows][16 bytes])
  const int tb_off0 = (pl >> 5) * a.T * 16 + ((pl & 31) >> 1);
  const int tb_off1 = ((pl + HALF) >> 5) * a.T * 16 + (((pl + HALF) & 31) >> 1), tb_sh = (pl & 1) * 4;  // HALF is even

  auto issue_tips = [&](int64_t p0) {  // one bulk-TMA copy stages the tip rows of a tile
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&tipbar[0], (uint32_t)tipbuf_bytes);
      bulk_g2s(tipbuf, a.tips4 + (size_t)(p0 / 32) * a.T * 16, (uint32_t)tipbuf_bytes, &tipbar[0]);
    }
  };
  // canonical level 0 for a finished tile: each group of 32 consecutive patterns is folded
  // by one warp. Runs at the start of the NEXT tile (after its barrier), off the critical path.
  auto fold_tile = [&](int64_t done_tile, int sbuf) {
    for (int g = warp; g < GROUPS; g += NT / 32) {
      const double v = warp_fold(stage[sbuf * TILE + g * 32 + lane]);
      if (lane == 0) a.groups[done_tile * GROUPS + g] = v;
    }
  };

  // work unit = one tile; each resident CTA owns a contiguous run of tiles (adjacent tiles
  // keep the tip reads and CLV writes of a CTA in the same DRAM pages)
  const int64_t slab_end = min(ntiles, a.tile_end);
  const int64_t per_cta = (slab_end - a.tile_begin + gridDim.x - 1) / gridDim.x;
  const int64_t tile_lo = a.tile_begin + (int64_t)blockIdx.x * per_cta, tile_hi = min(slab_end, tile_lo + per_cta);
  uint32_t tile_seq = 0;  // tiles processed by this CTA

  for (int64_t tile = tile_lo; tile < tile_hi; ++tile, ++tile_seq) {
    {
      const int64_t p0 = tile * TILE;
      const int buf = tile_seq & 1;  // site-value staging buffer of this tile
      __syncthreads();  // every warp has finished the previous tile: the tip buffer is free
      issue_tips(p0);
      if (tile_seq > 0) fold_tile(tile - 1, buf ^ 1);
      fetch_matrices(0, 0);  // matrices of step 0 while the tips are in flight
      mbar_wait(&tipbar[0], tile_seq & 1);
      const uint8_t *tb = tipbuf;

      int64_t pat[R], item_off[R];
      bool active[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        pat[r] = p0 + pl + r * HALF;
        active[r] = pat[r] < a.N;
        item_off[r] = (pat[r] * K + k) * 4;  // this thread's doubles inside any CLV array
      }
      d4 cur[R];
      int cur_sc[R], sp = 0;  // sp in units of elements (R*NT per stack level)
#pragma unroll
      for (int r = 0; r < R; ++r) { cur[r] = d4{0, 0, 0, 0}; cur_sc[r] = 0; }

      // software pipeline: while step s computes, the words of step s+1 are read from shared
      // memory and, once landed, its tip masks; a step starts with all but its matrices
      int4 iw = sprog[0];  // kinds, lidx, ridx, out_slot
      int mlb0, mlb1, mrb0, mrb1;  // raw tip bytes of (left,right) x (r=0,1)
      {
        const int lrow = ((iw.x & 3) == OPK_TIP) ? iw.y : 0, rrow = (((iw.x >> 2) & 3) == OPK_TIP) ? iw.z : 0;
        mlb0 = tb[lrow * 16 + tb_off0]; mlb1 = tb[lrow * 16 + tb_off1];
        mrb0 = tb[rrow * 16 + tb_off0]; mrb1 = tb[rrow * 16 + tb_off1];
      }
      // rolling half-matrix pipeline over the warp's shared-memory slots: slot (step & 1)
      // holds this step's two matrix sets; the copy of the next step's is issued first
      double hA[8], hB[8];
      cp_async_wait<0>();
      __syncwarp();
      {
        const double *pm0 = myring + k * 2;
        load_half<K>(pm0, 0, hA);
        load_half<K>(pm0, 1, hB);
      }

      for (int step = 0; step < a.n_instr; ++step) {
        __syncwarp();                               // all lanes are done with slot (step+1)&1
        fetch_matrices(step + 1, (step + 1) & 1);   // the root's matrix after the last step
        const double *pm_ = myring + (size_t)(step & 1) * 2 * PM + k * 2;
        const double *pn_ = myring + (size_t)((step + 1) & 1) * 2 * PM + k * 2;
        const int lkind = iw.x & 3, rkind = (iw.x >> 2) & 3, push = iw.x & 16, lidx = iw.y, ridx = iw.z;
        const int4 ow = RETAIN ? sprog[2 * step + 1] : make_int4(0, 0, 0, 0);  // out_clv, out_sc
        const int4 nw = sprog[2 * step + 2];                                   // next step's word
        if (push) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            mystack[sp + r * NT] = cur[r];
            mystack_sc[sp + r * NT] = cur_sc[r];
          }
          sp += R * NT;
        }
        // ---- left operand vectors
        d4 lv[R];
        int sc[R];
        if (lkind == OPK_TIP) {
          lv[0] = mask_vec(mlb0 >> tb_sh); lv[1] = mask_vec(mlb1 >> tb_sh);
          sc[0] = sc[1] = 0;
        } else if (lkind == OPK_CUR) {
#pragma unroll
          for (int r = 0; r < R; ++r) { lv[r] = cur[r]; sc[r] = cur_sc[r]; }
        } else if (lkind == OPK_POP) {
          sp -= R * NT;
#pragma unroll
          for (int r = 0; r < R; ++r) { lv[r] = mystack[sp + r * NT]; sc[r] = mystack_sc[sp + r * NT]; }
        } else {
          const double *src = a.node_clv[lidx];
          const int32_t *ssc = a.node_sc[lidx];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            lv[r] = d4{0, 0, 0, 0}; sc[r] = 0;
            if (active[r]) { lv[r] = ld256_stream(src + item_off[r]); sc[r] = ssc[pat[r]]; }
          }
        }
        // ---- x = P_l * lv, half by half; each finished half is refilled with P_r at once
        double x[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) matvec_half(hA, lv[r], x[r][0], x[r][1]);
        load_half<K>(pm_ + PM, 0, hA);
#pragma unroll
        for (int r = 0; r < R; ++r) matvec_half(hB, lv[r], x[r][2], x[r][3]);
        load_half<K>(pm_ + PM, 1, hB);
        // ---- right operand vectors
        d4 rv[R];
        if (rkind == OPK_TIP) {
          rv[0] = mask_vec(mrb0 >> tb_sh); rv[1] = mask_vec(mrb1 >> tb_sh);
        } else if (rkind == OPK_CUR) {
#pragma unroll
          for (int r = 0; r < R; ++r) { rv[r] = cur[r]; sc[r] += cur_sc[r]; }
        } else if (rkind == OPK_POP) {
          sp -= R * NT;
#pragma unroll
          for (int r = 0; r < R; ++r) { rv[r] = mystack[sp + r * NT]; sc[r] += mystack_sc[sp + r * NT]; }
        } else {
          const double *src = a.node_clv[ridx];
          const int32_t *ssc = a.node_sc[ridx];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            rv[r] = d4{0, 0, 0, 0};
            if (active[r]) { rv[r] = ld256_stream(src + item_off[r]); sc[r] += ssc[pat[r]]; }
          }
        }
        // next step's tip bytes (its word has landed by now); consumed one iteration later
        {
          const int lrow = ((nw.x & 3) == OPK_TIP) ? nw.y : 0, rrow = (((nw.x >> 2) & 3) == OPK_TIP) ? nw.z : 0;
          mlb0 = tb[lrow * 16 + tb_off0]; mlb1 = tb[lrow * 16 + tb_off1];
          mrb0 = tb[rrow * 16 + tb_off0]; mrb1 = tb[rrow * 16 + tb_off1];
        }
        iw = nw;
        // ---- y = P_r * rv and the products; freed halves are refilled with the NEXT step's
        // left matrix (the root matrix after the last step)
        d4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          double y0, y1;
          matvec_half(hA, rv[r], y0, y1);
          v[r].x = x[r][0] * y0; v[r].y = x[r][1] * y1;
        }
        cp_async_wait<0>();  // next step's matrices have landed in the other slot
        __syncwarp();
        load_half<K>(pn_, 0, hA);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          double y2, y3;
          matvec_half(hB, rv[r], y2, y3);
          v[r].z = x[r][2] * y2; v[r].w = x[r][3] * y3;
        }
        load_half<K>(pn_, 1, hB);
        double *oc = reinterpret_cast<double *>(((uint64_t)(uint32_t)ow.y << 32) | (uint32_t)ow.x);
        int32_t *os = reinterpret_cast<int32_t *>(((uint64_t)(uint32_t)ow.w << 32) | (uint32_t)ow.z);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          int h = max(max(hi32(v[r].x), hi32(v[r].y)), max(hi32(v[r].z), hi32(v[r].w)));
#pragma unroll
          for (int off = K / 2; off >= 1; off >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, off));
          if (h < kScaleHiThresh) {
            v[r].x *= 0x1p+256; v[r].y *= 0x1p+256; v[r].z *= 0x1p+256; v[r].w *= 0x1p+256;
            ++sc[r];
          }
          cur[r] = v[r];
          cur_sc[r] = sc[r];
          if (RETAIN) {
            if (oc != nullptr && active[r]) {
              st256(oc + item_off[r], v[r]);
              if (k == 0) os[pat[r]] = sc[r];
            }
          }
        }
      }
      // ---- root-edge join (root4_kernel's arithmetic): P applies to the b side only.
      // iw and the tip bytes already hold the root step (fetched by the last iteration).
      {
        const int akind = iw.x & 3, bkind = (iw.x >> 2) & 3;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          d4 av{0, 0, 0, 0}, bv{0, 0, 0, 0};
          int c = 0;
          // the POP operand (if any) was pushed before the CUR one was computed
          if (akind == OPK_TIP) av = mask_vec((r ? mlb1 : mlb0) >> tb_sh);
          else if (akind == OPK_CUR) { av = cur[r]; c += cur_sc[r]; }
          else if (akind == OPK_POP) { av = mystack[sp - R * NT + r * NT]; c += mystack_sc[sp - R * NT + r * NT]; }
          else if (active[r]) { av = ld256_stream(a.node_clv[iw.y] + item_off[r]); c += a.node_sc[iw.y][pat[r]]; }
          if (bkind == OPK_TIP) bv = mask_vec((r ? mrb1 : mrb0) >> tb_sh);
          else if (bkind == OPK_CUR) { bv = cur[r]; c += cur_sc[r]; }
          else if (bkind == OPK_POP) { bv = mystack[sp - R * NT + r * NT]; c += mystack_sc[sp - R * NT + r * NT]; }
          else if (active[r]) { bv = ld256_stream(a.node_clv[iw.z] + item_off[r]); c += a.node_sc[iw.z][pat[r]]; }
          double y[4];
          matvec_half(hA, bv, y[0], y[1]);  // hA/hB hold the root matrix (loaded by the last step)
          matvec_half(hB, bv, y[2], y[3]);
          const double lk = (((pi0 * av.x) * y[0] + (pi1 * av.y) * y[1]) + (pi2 * av.z) * y[2]) + (pi3 * av.w) * y[3];
          double l = pk_prob * lk;
#pragma unroll
          for (int off = 1; off < K; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
          if (k == 0) {
            double wl = 0.0;
            if (active[r]) {
              const int64_t p = pat[r];
              double lnl;
              if (a.pinvar >= 0.0) {
                const int m = a.inv[p];
                const double pv = (m & 1 ? pi0 : 0.0) + (m & 2 ? pi1 : 0.0) + (m & 4 ? pi2 : 0.0) + (m & 8 ? pi3 : 0.0);
                lnl = lnl_pinvar(l, c, a.pinvar, pv);
              } else {
                lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
              }
              if (a.site_lnl) a.site_lnl[p] = lnl;
              wl = (a.weights ? a.weights[p] : 1.0) * lnl;
            }
            stage[buf * TILE + pl + r * HALF] = wl;
          }
        }
      }
    }
  }
  __syncthreads();
  if (tile_seq > 0) fold_tile(tile_hi - 1, (tile_seq - 1) & 1);  // the last tile of this CTA
}

// canonical level 1: partial[b] = fold of the 32 group sums of block b (groups past N are 0)
__global__ void __launch_bounds__(32) fold_groups_kernel(const double *__restrict__ groups, int64_t n_groups,
                                                         double *__restrict__ partials) {
  const int64_t g = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const double v = warp_fold(g < n_groups ? groups[g] : 0.0);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

}  // namespace phylo
