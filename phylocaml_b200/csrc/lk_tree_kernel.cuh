// Tree-fused 4-state pruning: ONE launch evaluates the whole post-order schedule.
//
// Sites are independent, so a thread can carry its (pattern, rate class) through every node
// of the tree. The running CLV lives in registers; only the first child of a node with two
// interior children is parked on a small shared-memory stack (depth = Strahler number of the
// tree, <= log2 T); tips are 1 byte per pattern. Compared with one streaming kernel per node
// (3 CLV moves per update), HBM sees at most ONE write per update (RETAIN: every interior
// CLV is still produced in its slot for later incremental use) or nothing but the tips.
//
// Blackwell specifics:
//   * tip masks of a tile ([T rows] x [TILE patterns] bytes) are staged by bulk-TMA
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), double buffered so the
//     next tile's tips land while this tile is computed;
//   * the two transition matrices of each step travel through a 4-slot TMA-fed ring
//     (mbarrier full barriers), stored k-interleaved so the K lanes of a pattern read
//     adjacent 16-byte chunks (bank-conflict free);
//   * retained CLVs leave as 256-bit stores, 1 KB contiguous per warp.
//
// Arithmetic and its order are those of prune4_kernel / root4_kernel (=> same bits); a
// one-hot tip mask takes column j of P directly (0 + P[i][j] == P[i][j]).
#pragma once
#include "common.cuh"

namespace phylo {

enum : int { OPK_TIP = 0, OPK_CUR = 1, OPK_POP = 2, OPK_STORED = 3 };

struct __align__(16) TreeInstr {
  int lkind, rkind;      // OPK_*
  int lidx, ridx;        // tip row, or node slot for OPK_STORED
  int push_first;        // park the running CLV on the stack before this step
  int out_slot;          // node slot that receives the result (RETAIN), or -1
  int pad0, pad1;
};

constexpr int kRing = 4;          // transition-matrix ring slots
constexpr int kTreeThreads = 256;

// ---- mbarrier / bulk-TMA primitives (sm_90+ PTX, SASS: SYNCS / UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// element e (= i*4+j) of rate class k in the k-interleaved matrix layout written by
// pt_build_kernel(interleave=1): [e/2][k][e%2]
template <int K>
__device__ __forceinline__ int pidx(int e, int k) {
  return ((e >> 1) * K + k) * 2 + (e & 1);
}

// x[i] = sum_j P[i][j] v[j], P read from the shared-memory ring slot into registers
template <int K>
__device__ __forceinline__ void apply_inner(const double *ps, int k, const d4 &v, double (&x)[4]) {
  double pm[16];
#pragma unroll
  for (int ep = 0; ep < 8; ++ep) {
    const double2 t = *reinterpret_cast<const double2 *>(ps + (ep * K + k) * 2);
    pm[2 * ep] = t.x;
    pm[2 * ep + 1] = t.y;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    x[i] = ((pm[i * 4 + 0] * v.x + pm[i * 4 + 1] * v.y) + pm[i * 4 + 2] * v.z) + pm[i * 4 + 3] * v.w;
}

// x[i] = sum_{j in mask} P[i][j] (ascending j); one-hot masks take the column directly
template <int K>
__device__ __forceinline__ void apply_tip(const double *ps, int k, int m, double (&x)[4]) {
  m &= 15;
  if (__popc(m) == 1) {
    const int j = __ffs(m) - 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = ps[pidx<K>(i * 4 + j, k)];
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double a = (m & 1) ? ps[pidx<K>(i * 4 + 0, k)] : 0.0;
      a += (m & 2) ? ps[pidx<K>(i * 4 + 1, k)] : 0.0;
      a += (m & 4) ? ps[pidx<K>(i * 4 + 2, k)] : 0.0;
      a += (m & 8) ? ps[pidx<K>(i * 4 + 3, k)] : 0.0;
      x[i] = a;
    }
  }
}

struct TreeArgs {
  const TreeInstr *prog;   // n_instr steps + 1 root step (kinds/idx only)
  int n_instr;
  const double *P;         // [2*n_instr + 1][16*K] k-interleaved
  const uint8_t *tips;     // [T][tip_stride]
  int64_t tip_stride;
  int T;
  int64_t N;
  double *const *node_clv; // per slot (RETAIN / OPK_STORED)
  int32_t *const *node_sc;
  const double *pi, *probs;
  double pinvar;
  const uint8_t *inv;
  const double *weights;
  double *site_lnl;
  double *partials;
  int stack_depth;
};

template <int K, int R, bool RETAIN>
__global__ void __launch_bounds__(kTreeThreads, 1) lk_tree4_kernel(const TreeArgs a) {
  constexpr int NT = kTreeThreads;
  constexpr int TILE = NT * R / K;          // patterns per tile
  constexpr int TILES = kLnlBlock / TILE;   // tiles per 1024-pattern reduction block
  constexpr int PSLOT = 2 * 16 * K;         // doubles per ring slot (left + right matrices)
  static_assert(TILE >= 16 && kLnlBlock % TILE == 0, "tile must divide the reduction block");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  // carve-up (all offsets multiples of 16 B)
  double *vals = reinterpret_cast<double *>(smem_raw);                       // [1024]
  double *wsum = vals + kLnlBlock;                                           // [32]
  uint64_t *tipbar = reinterpret_cast<uint64_t *>(wsum + 32);                // [2]
  uint64_t *pbar = tipbar + 2;                                               // [kRing]
  double *pring = reinterpret_cast<double *>(pbar + kRing + 2);              // [kRing][PSLOT]
  d4 *stack = reinterpret_cast<d4 *>(pring + kRing * PSLOT);                 // [depth][R][NT]
  int *stack_sc = reinterpret_cast<int *>(stack + (size_t)a.stack_depth * R * NT);  // [depth][R][NT]
  uint8_t *tipbuf = reinterpret_cast<uint8_t *>(stack_sc + (size_t)a.stack_depth * R * NT);
  tipbuf = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tipbuf) + 127) & ~(uintptr_t)127);
  const size_t tipbuf_bytes = (size_t)a.T * TILE;                            // per buffer

  const int tid = threadIdx.x, k = tid % K, pl0 = tid / K;
  const int n_steps = a.n_instr + 1;  // + root step
  const int64_t nblocks = (a.N + kLnlBlock - 1) / kLnlBlock;

  if (tid == 0) {
    mbar_init(&tipbar[0], 1);
    mbar_init(&tipbar[1], 1);
    for (int s = 0; s < kRing; ++s) mbar_init(&pbar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const double pi0 = a.pi[0], pi1 = a.pi[1], pi2 = a.pi[2], pi3 = a.pi[3], pk = a.probs[k];

  // warp 0 stages the tip rows of one tile
  auto issue_tips = [&](int64_t p0, int buf) {
    if (tid < 32) {
      if (tid == 0) mbar_expect_tx(&tipbar[buf], (uint32_t)tipbuf_bytes);
      __syncwarp();
      for (int t = tid; t < a.T; t += 32)
        bulk_g2s(tipbuf + (size_t)buf * tipbuf_bytes + (size_t)t * TILE, a.tips + (size_t)t * a.tip_stride + p0,
                 TILE, &tipbar[buf]);
    }
  };
  auto issue_p = [&](int step, uint32_t seq) {  // thread 0 only
    const int slot = seq % kRing;
    const int bl = (step < a.n_instr) ? 2 * step : 2 * a.n_instr;
    const int br = (step < a.n_instr) ? 2 * step + 1 : 2 * a.n_instr;
    mbar_expect_tx(&pbar[slot], PSLOT * 8);
    bulk_g2s(pring + slot * PSLOT, a.P + (size_t)bl * 16 * K, 16 * K * 8, &pbar[slot]);
    bulk_g2s(pring + slot * PSLOT + 16 * K, a.P + (size_t)br * 16 * K, 16 * K * 8, &pbar[slot]);
  };

  uint32_t tile_seq = 0;  // tiles processed by this CTA (tip double buffer phase)
  uint32_t pseq = 0;      // steps processed by this CTA (ring phase)

  // first tile of this CTA
  if ((int64_t)blockIdx.x < nblocks) issue_tips((int64_t)blockIdx.x * kLnlBlock, 0);

  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    for (int i = tid; i < kLnlBlock; i += NT) vals[i] = 0.0;
    const int64_t blk_p0 = blk * kLnlBlock;
    const int tiles_here = (int)min((int64_t)TILES, (a.N - blk_p0 + TILE - 1) / TILE);
    for (int sub = 0; sub < tiles_here; ++sub, ++tile_seq) {
      const int64_t p0 = blk_p0 + (int64_t)sub * TILE;
      const int buf = tile_seq & 1;
      // prefetch the next tile's tips into the other buffer (free since the previous tile's
      // closing barrier)
      {
        int64_t np0 = -1;
        if (sub + 1 < tiles_here) np0 = p0 + TILE;
        else if (blk + gridDim.x < nblocks) np0 = (blk + gridDim.x) * kLnlBlock;
        if (np0 >= 0) {
          if (tid == 0) fence_proxy_async();
          issue_tips(np0, buf ^ 1);
        }
      }
      // ring prologue: matrices of the first kRing-1 steps
      if (tid == 0) {
        fence_proxy_async();
        for (int s = 0; s < kRing - 1 && s < n_steps; ++s) issue_p(s, pseq + s);
      }
      mbar_wait(&tipbar[buf], (tile_seq >> 1) & 1);
      const uint8_t *tb = tipbuf + (size_t)buf * tipbuf_bytes;

      d4 cur[R];
      int cur_sc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) { cur[r] = d4{0, 0, 0, 0}; cur_sc[r] = 0; }
      int sp = 0;

      for (int step = 0; step < n_steps; ++step, ++pseq) {
        __syncthreads();  // every thread is done with step-1: its ring slot may be refilled
        if (tid == 0 && step + kRing - 1 < n_steps) {
          fence_proxy_async();
          issue_p(step + kRing - 1, pseq + kRing - 1);
        }
        const TreeInstr ins = a.prog[step];
        mbar_wait(&pbar[pseq % kRing], (pseq / kRing) & 1);
        const double *psl = pring + (pseq % kRing) * PSLOT, *psr = psl + 16 * K;

        if (ins.push_first) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            stack[((size_t)sp * R + r) * NT + tid] = cur[r];
            stack_sc[((size_t)sp * R + r) * NT + tid] = cur_sc[r];
          }
          ++sp;
        }
        const bool is_root = (step == a.n_instr);
        // ---- fetch operands (kinds are uniform over the grid)
        d4 lv[R], rv[R];
        int lsc[R], rsc[R];
        const int popl = (ins.lkind == OPK_POP), popr = (ins.rkind == OPK_POP);
        if (popl || popr) --sp;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int pl = pl0 + r * (NT / K);
          const int64_t p = p0 + pl;
          lsc[r] = rsc[r] = 0;
          lv[r] = rv[r] = d4{0, 0, 0, 0};
          if (ins.lkind == OPK_CUR) { lv[r] = cur[r]; lsc[r] = cur_sc[r]; }
          else if (ins.lkind == OPK_POP) {
            lv[r] = stack[((size_t)sp * R + r) * NT + tid];
            lsc[r] = stack_sc[((size_t)sp * R + r) * NT + tid];
          } else if (ins.lkind == OPK_STORED) {
            if (p < a.N) { lv[r] = ld256_stream(a.node_clv[ins.lidx] + (p * K + k) * 4); lsc[r] = a.node_sc[ins.lidx][p]; }
          }
          if (ins.rkind == OPK_CUR) { rv[r] = cur[r]; rsc[r] = cur_sc[r]; }
          else if (ins.rkind == OPK_POP) {
            rv[r] = stack[((size_t)sp * R + r) * NT + tid];
            rsc[r] = stack_sc[((size_t)sp * R + r) * NT + tid];
          } else if (ins.rkind == OPK_STORED) {
            if (p < a.N) { rv[r] = ld256_stream(a.node_clv[ins.ridx] + (p * K + k) * 4); rsc[r] = a.node_sc[ins.ridx][p]; }
          }
        }
        if (!is_root) {
          // ---- pruning update
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int pl = pl0 + r * (NT / K);
            double x[4], y[4];
            if (ins.lkind == OPK_TIP) apply_tip<K>(psl, k, tb[(size_t)ins.lidx * TILE + pl], x);
            else apply_inner<K>(psl, k, lv[r], x);
            if (ins.rkind == OPK_TIP) apply_tip<K>(psr, k, tb[(size_t)ins.ridx * TILE + pl], y);
            else apply_inner<K>(psr, k, rv[r], y);
            d4 v{x[0] * y[0], x[1] * y[1], x[2] * y[2], x[3] * y[3]};
            int h = max(max(hi32(v.x), hi32(v.y)), max(hi32(v.z), hi32(v.w)));
#pragma unroll
            for (int off = K / 2; off >= 1; off >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, off));
            const bool rescale = h < kScaleHiThresh;
            if (rescale) { v.x *= 0x1p+256; v.y *= 0x1p+256; v.z *= 0x1p+256; v.w *= 0x1p+256; }
            cur[r] = v;
            cur_sc[r] = lsc[r] + rsc[r] + (rescale ? 1 : 0);
            if (RETAIN && ins.out_slot >= 0) {
              const int64_t p = p0 + pl;
              if (p < a.N) {
                st256(a.node_clv[ins.out_slot] + (p * K + k) * 4, v);
                if (k == 0) a.node_sc[ins.out_slot][p] = cur_sc[r];
              }
            }
          }
        } else {
          // ---- root-edge join (root4_kernel's arithmetic): P applies to the b side only
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int pl = pl0 + r * (NT / K);
            const int64_t p = p0 + pl;
            d4 av = lv[r];
            if (ins.lkind == OPK_TIP) {
              const int m = tb[(size_t)ins.lidx * TILE + pl];
              av = d4{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
            }
            double y[4];
            if (ins.rkind == OPK_TIP) {
              const int m = tb[(size_t)ins.ridx * TILE + pl];
              const d4 bv{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
              apply_inner<K>(psl, k, bv, y);
            } else {
              apply_inner<K>(psl, k, rv[r], y);
            }
            const double lk = (((pi0 * av.x) * y[0] + (pi1 * av.y) * y[1]) + (pi2 * av.z) * y[2]) + (pi3 * av.w) * y[3];
            double l = pk * lk;
#pragma unroll
            for (int off = 1; off < K; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
            if (k == 0 && p < a.N) {
              const int c = lsc[r] + rsc[r];
              double lnl;
              if (a.pinvar >= 0.0) {
                const int m = a.inv[p];
                const double pv = (m & 1 ? pi0 : 0.0) + (m & 2 ? pi1 : 0.0) + (m & 4 ? pi2 : 0.0) + (m & 8 ? pi3 : 0.0);
                lnl = log((1.0 - a.pinvar) * ldexp(l, -kScaleExp * c) + a.pinvar * pv);
              } else {
                lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
              }
              if (a.site_lnl) a.site_lnl[p] = lnl;
              vals[sub * TILE + pl] = (a.weights ? a.weights[p] : 1.0) * lnl;
            }
          }
        }
      }
      __syncthreads();  // tile done: tip buffer `buf` and the ring are free again
    }
    const double rsum = block_fold_1024(vals, wsum);
    if (tid == 0) a.partials[blk] = rsum;
    __syncthreads();
  }
}

}  // namespace phylo
