// Tree-fused 4-state pruning: ONE launch evaluates the whole post-order schedule.
//
// Sites are independent, so a thread can carry its (pattern, rate class) through every node
// of the tree. The running CLV lives in registers; only the first child of a node with two
// interior children is parked on a small shared-memory stack (depth = Strahler number of the
// tree, <= log2 T); tips are 1 byte per pattern. Compared with one streaming kernel per node
// (3 CLV moves per update), HBM sees at most ONE write per update (RETAIN: every interior
// CLV is still produced in its slot for later incremental use) or nothing but the tips.
//
// Execution model (v2, after profiling v1 -- see profiles/README.md):
//   * thread = one (pattern, rate class); CTA = 256 threads = a tile of 256/K patterns;
//     two CTAs per SM (<= 128 registers) so 16 warps hide the fp64 / shared-memory latency;
//   * warps run FREE through the schedule: there is no CTA barrier per step. The only
//     barriers are one per tile (tip buffer hand-over) and one per 1024-pattern block
//     (canonical site-sum fold);
//   * tip masks of a tile ([T rows] x [TILE patterns] bytes) are staged by bulk-TMA
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), double buffered so the
//     next tile's tips land while this tile is computed;
//   * the transition matrices of a step are read through L1 (k-interleaved layout: the K
//     lanes of a pattern read adjacent 16-byte chunks, all patterns the same address) and
//     software-prefetched a few steps ahead (prefetch.global.L1);
//   * retained CLVs leave as 256-bit stores, 1 KB contiguous per warp.
//
// Arithmetic and its order are those of prune4_kernel / root4_kernel (=> same bits); a
// one-hot tip mask takes column j of P directly (0 + P[i][j] == P[i][j]).
#pragma once
#include "common.cuh"
#include "lk_kernels.cuh"  // tip_contrib

namespace phylo {

enum : int { OPK_TIP = 0, OPK_CUR = 1, OPK_POP = 2, OPK_STORED = 3 };

// One step of the compiled schedule, 64 bytes, everything the kernel needs without a
// dependent table lookup.
struct __align__(16) TreeInstr {
  int lkind, rkind;      // OPK_*
  int lidx, ridx;        // tip row (OPK_TIP) or node slot (OPK_STORED)
  int push_first;        // park the running CLV on the stack before this step
  int out_slot;          // node slot that receives the result, or -1
  int pad0, pad1;
  double *out_clv;       // RETAIN: where the result goes (NULL = nowhere)
  int32_t *out_sc;
  const double *l_clv;   // OPK_STORED operands: resident CLV / scale arrays
  const double *r_clv;
};
static_assert(sizeof(TreeInstr) == 64, "TreeInstr is four 16-byte words");

constexpr int kTreeThreads = 256;
constexpr int kTreePrefetch = 4;  // steps of look-ahead for the transition matrices

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
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Transition matrices are stored k-interleaved by pt_build_kernel(interleave=1): element
// e (= i*4+j) of rate class k sits at [e/2][k][e%2]. `pk` points at this thread's k.
// x[i] = sum_j P[i][j] v[j]
template <int K>
__device__ __forceinline__ void apply_inner(const double *__restrict__ pk, const d4 &v, double (&x)[4]) {
  double pm[16];
#pragma unroll
  for (int ep = 0; ep < 8; ++ep) {
    const double2 t = __ldg(reinterpret_cast<const double2 *>(pk + ep * K * 2));
    pm[2 * ep] = t.x;
    pm[2 * ep + 1] = t.y;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    x[i] = ((pm[i * 4 + 0] * v.x + pm[i * 4 + 1] * v.y) + pm[i * 4 + 2] * v.z) + pm[i * 4 + 3] * v.w;
}

// x[i] = sum_{j in mask} P[i][j] (ascending j); one-hot masks take column j directly
template <int K>
__device__ __forceinline__ void apply_tip(const double *__restrict__ pk, int m, double (&x)[4]) {
  m &= 15;
  if (__popc(m) == 1) {
    const int j = __ffs(m) - 1;
    const double *c = pk + (j >> 1) * K * 2 + (j & 1);  // element (i, j) is at c[i * 4K]
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __ldg(c + i * 4 * K);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 p01 = __ldg(reinterpret_cast<const double2 *>(pk + (2 * i) * K * 2));
      const double2 p23 = __ldg(reinterpret_cast<const double2 *>(pk + (2 * i + 1) * K * 2));
      double acc = (m & 1) ? p01.x : 0.0;
      acc += (m & 2) ? p01.y : 0.0;
      acc += (m & 4) ? p23.x : 0.0;
      acc += (m & 8) ? p23.y : 0.0;
      x[i] = acc;
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
  int32_t *const *node_sc; // per slot: scale counters of OPK_STORED operands
  const double *pi, *probs;
  double pinvar;
  const uint8_t *inv;
  const double *weights;
  double *site_lnl;
  double *partials;
  int stack_depth;
};

// operand vector of one step: tips become 0/1 vectors (so x_i = sum_j P[i][j]*L_j reproduces
// the ascending-j sum over the mask bit-for-bit), CUR is the running CLV, POP comes off the
// shared-memory stack, STORED streams from HBM.
__device__ __forceinline__ d4 mask_vec(int m) {
  return d4{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
}

template <int K, bool RETAIN>
__global__ void __launch_bounds__(kTreeThreads, 2) lk_tree4_kernel(const TreeArgs a) {
  constexpr int NT = kTreeThreads;
  constexpr int TILE = NT / K;              // patterns per tile
  constexpr int TILES = kLnlBlock / TILE;   // tiles per 1024-pattern reduction block
  constexpr int PM = 16 * K;                // doubles per transition matrix set (all k)
  static_assert(TILE >= 16 && kLnlBlock % TILE == 0, "tile must divide the reduction block");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *vals = reinterpret_cast<double *>(smem_raw);                       // [1024]
  double *wsum = vals + kLnlBlock;                                           // [32]
  uint64_t *tipbar = reinterpret_cast<uint64_t *>(wsum + 32);                // [2] (+2 pad)
  d4 *stack = reinterpret_cast<d4 *>(tipbar + 4);                            // [depth][NT]
  int *stack_sc = reinterpret_cast<int *>(stack + (size_t)a.stack_depth * NT);  // [depth][NT]
  int4 *sprog = reinterpret_cast<int4 *>(stack_sc + (size_t)a.stack_depth * NT);  // [n_steps][4]
  uint8_t *tipbuf = reinterpret_cast<uint8_t *>(sprog + 4 * (size_t)(a.n_instr + 1));
  tipbuf = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tipbuf) + 127) & ~(uintptr_t)127);
  const size_t tipbuf_bytes = (size_t)a.T * TILE;                            // per buffer

  const int tid = threadIdx.x, k = tid % K, pl = tid / K, lane = tid & 31;
  const int n_steps = a.n_instr + 1;  // + root step
  const int64_t nblocks = (a.N + kLnlBlock - 1) / kLnlBlock;

  if (tid == 0) {
    mbar_init(&tipbar[0], 1);
    mbar_init(&tipbar[1], 1);
    fence_mbar_init();
  }
  // the compiled schedule lives in shared memory for the whole kernel
  for (int i = tid; i < 4 * n_steps; i += NT) sprog[i] = __ldg(reinterpret_cast<const int4 *>(a.prog) + i);
  __syncthreads();

  const double pi0 = a.pi[0], pi1 = a.pi[1], pi2 = a.pi[2], pi3 = a.pi[3], pk_prob = a.probs[k];
  const double *Pk = a.P + k * 2;  // this thread's rate class inside every matrix set
  d4 *mystack = stack + tid;
  int *mystack_sc = stack_sc + tid;

  // warp 0 stages the tip rows of one tile
  auto issue_tips = [&](int64_t p0, int buf) {
    if (tid < 32) {
      if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(&tipbar[buf], (uint32_t)tipbuf_bytes);
      }
      __syncwarp();
      for (int t = tid; t < a.T; t += 32)
        bulk_g2s(tipbuf + (size_t)buf * tipbuf_bytes + (size_t)t * TILE, a.tips + (size_t)t * a.tip_stride + p0,
                 TILE, &tipbar[buf]);
    }
  };

  uint32_t tile_seq = 0;  // tiles processed by this CTA (tip double-buffer phase)
  if ((int64_t)blockIdx.x < nblocks) issue_tips((int64_t)blockIdx.x * kLnlBlock, 0);

  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    for (int i = tid; i < kLnlBlock; i += NT) vals[i] = 0.0;
    const int64_t blk_p0 = blk * kLnlBlock;
    const int tiles_here = (int)min((int64_t)TILES, (a.N - blk_p0 + TILE - 1) / TILE);
    for (int sub = 0; sub < tiles_here; ++sub, ++tile_seq) {
      const int64_t p0 = blk_p0 + (int64_t)sub * TILE;
      const int64_t p = p0 + pl;
      const bool active = p < a.N;
      const int64_t item_off = (p * K + k) * 4;  // this thread's doubles inside any CLV array
      const int buf = tile_seq & 1;
      __syncthreads();  // every warp has finished the previous tile: buffer buf^1 is free
      {
        int64_t np0 = -1;
        if (sub + 1 < tiles_here) np0 = p0 + TILE;
        else if (blk + gridDim.x < nblocks) np0 = (blk + gridDim.x) * kLnlBlock;
        if (np0 >= 0) issue_tips(np0, buf ^ 1);
      }
      // warm L1 with the matrices of the first steps
      if (lane < 2 * K) {
        for (int s = 0; s < kTreePrefetch && s < n_steps; ++s)
          prefetch_l1(a.P + (size_t)(2 * s) * PM + lane * 16);
      }
      mbar_wait(&tipbar[buf], (tile_seq >> 1) & 1);
      const uint8_t *tb = tipbuf + (size_t)buf * tipbuf_bytes + pl;

      d4 cur{0, 0, 0, 0};
      int cur_sc = 0, sp = 0;  // sp counts stack entries in units of NT elements

      // software pipeline: while step s computes, the words of step s+1 are read from shared
      // memory and, once they have landed, its tip masks -- a step starts with everything but
      // its matrices in registers
      int4 iw = sprog[0];    // lkind, rkind, lidx, ridx   (idx: tip row | node slot)
      int4 iw2 = sprog[1];   // push_first, out_slot, -, -
      int ml = tb[(size_t)((iw.x == OPK_TIP) ? iw.z : 0) * TILE];
      int mr = tb[(size_t)((iw.y == OPK_TIP) ? iw.w : 0) * TILE];
      const double *pm_ = Pk;                                       // matrices of this step
      const double *pf_ = a.P + (size_t)(2 * kTreePrefetch) * PM + lane * 16;  // L1 prefetch cursor

      for (int step = 0; step < a.n_instr; ++step, pm_ += 2 * PM, pf_ += 2 * PM) {
        // ---- matrix loads first: they depend on nothing but the step and the (already
        // present) tip masks. A tip operand whose mask is one-hot in EVERY lane of the warp
        // (warp-uniform vote: no divergence) only needs column j of its matrix.
        const int lkind = iw.x, rkind = iw.y;
        const int mlc = ml & 15, mrc = mr & 15;
        const bool lcol = (lkind == OPK_TIP) && __all_sync(0xffffffffu, __popc(mlc) == 1);
        const bool rcol = (rkind == OPK_TIP) && __all_sync(0xffffffffu, __popc(mrc) == 1);
        double pml[16], pmr[16];
        if (lcol) {
          const int j = __ffs(mlc) - 1;
          const double *c = pm_ + (j >> 1) * K * 2 + (j & 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) pml[i] = __ldg(c + i * 4 * K);
        } else {
#pragma unroll
          for (int ep = 0; ep < 8; ++ep) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(pm_ + ep * K * 2));
            pml[2 * ep] = t.x; pml[2 * ep + 1] = t.y;
          }
        }
        if (rcol) {
          const int j = __ffs(mrc) - 1;
          const double *c = pm_ + PM + (j >> 1) * K * 2 + (j & 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) pmr[i] = __ldg(c + i * 4 * K);
        } else {
#pragma unroll
          for (int ep = 0; ep < 8; ++ep) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(pm_ + PM + ep * K * 2));
            pmr[2 * ep] = t.x; pmr[2 * ep + 1] = t.y;
          }
        }
        if (lane < 2 * K && step + kTreePrefetch < n_steps) prefetch_l1(pf_);
        const int lidx = iw.z, ridx = iw.w, push = iw2.x;
        const int4 ow = RETAIN ? sprog[4 * step + 2] : make_int4(0, 0, 0, 0);  // out_clv, out_sc
        const int4 nw = sprog[4 * step + 4], nw2 = sprog[4 * step + 5];       // next step's words
        if (push) {
          mystack[sp] = cur;
          mystack_sc[sp] = cur_sc;
          sp += NT;
        }
        // ---- operand vectors
        d4 lv, rv;
        int sc = 0;
        if (lkind == OPK_TIP) lv = mask_vec(ml);
        else if (lkind == OPK_CUR) { lv = cur; sc = cur_sc; }
        else if (lkind == OPK_POP) { sp -= NT; lv = mystack[sp]; sc = mystack_sc[sp]; }
        else {
          lv = d4{0, 0, 0, 0};
          const int4 sw = sprog[4 * step + 3];
          const double *src = reinterpret_cast<const double *>(((uint64_t)(uint32_t)sw.y << 32) | (uint32_t)sw.x);
          if (active) { lv = ld256_stream(src + item_off); sc = a.node_sc[lidx][p]; }
        }
        if (rkind == OPK_TIP) rv = mask_vec(mr);
        else if (rkind == OPK_CUR) { rv = cur; sc += cur_sc; }
        else if (rkind == OPK_POP) { sp -= NT; rv = mystack[sp]; sc += mystack_sc[sp]; }
        else {
          rv = d4{0, 0, 0, 0};
          const int4 sw = sprog[4 * step + 3];
          const double *src = reinterpret_cast<const double *>(((uint64_t)(uint32_t)sw.w << 32) | (uint32_t)sw.z);
          if (active) { rv = ld256_stream(src + item_off); sc += a.node_sc[ridx][p]; }
        }
        // next step's tip masks (its words have landed by now); consumed one iteration later
        ml = tb[(size_t)((nw.x == OPK_TIP) ? nw.z : 0) * TILE];
        mr = tb[(size_t)((nw.y == OPK_TIP) ? nw.w : 0) * TILE];
        iw = nw;
        iw2 = nw2;
        // ---- the update itself: straight-line fp64 (column fast path: x_i = P[i][j])
        double x[4], y[4];
        if (lcol) {
#pragma unroll
          for (int i = 0; i < 4; ++i) x[i] = pml[i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            x[i] = ((pml[i * 4 + 0] * lv.x + pml[i * 4 + 1] * lv.y) + pml[i * 4 + 2] * lv.z) + pml[i * 4 + 3] * lv.w;
        }
        if (rcol) {
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = pmr[i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            y[i] = ((pmr[i * 4 + 0] * rv.x + pmr[i * 4 + 1] * rv.y) + pmr[i * 4 + 2] * rv.z) + pmr[i * 4 + 3] * rv.w;
        }
        d4 v{x[0] * y[0], x[1] * y[1], x[2] * y[2], x[3] * y[3]};
        int h = max(max(hi32(v.x), hi32(v.y)), max(hi32(v.z), hi32(v.w)));
#pragma unroll
        for (int off = K / 2; off >= 1; off >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, off));
        if (h < kScaleHiThresh) {
          v.x *= 0x1p+256; v.y *= 0x1p+256; v.z *= 0x1p+256; v.w *= 0x1p+256;
          ++sc;
        }
        cur = v;
        cur_sc = sc;
        if (RETAIN) {
          double *oc = reinterpret_cast<double *>(((uint64_t)(uint32_t)ow.y << 32) | (uint32_t)ow.x);
          int32_t *os = reinterpret_cast<int32_t *>(((uint64_t)(uint32_t)ow.w << 32) | (uint32_t)ow.z);
          if (oc != nullptr && active) {
            st256(oc + item_off, v);
            if (k == 0) os[p] = sc;
          }
        }
      }
      // ---- root-edge join (root4_kernel's arithmetic): P applies to the b side only.
      // iw / ml / mr already hold the root step (fetched by the last iteration).
      {
        const int4 sw = sprog[4 * a.n_instr + 3];
        d4 av{0, 0, 0, 0}, bv{0, 0, 0, 0};
        int c = 0;
        // the POP operand (if any) was pushed before the CUR one was computed
        if (iw.x == OPK_TIP) av = mask_vec(ml);
        else if (iw.x == OPK_CUR) { av = cur; c += cur_sc; }
        else if (iw.x == OPK_POP) { av = mystack[sp - NT]; c += mystack_sc[sp - NT]; }
        else if (active) {
          const double *src = reinterpret_cast<const double *>(((uint64_t)(uint32_t)sw.y << 32) | (uint32_t)sw.x);
          av = ld256_stream(src + item_off); c += a.node_sc[iw.z][p];
        }
        if (iw.y == OPK_TIP) bv = mask_vec(mr);
        else if (iw.y == OPK_CUR) { bv = cur; c += cur_sc; }
        else if (iw.y == OPK_POP) { bv = mystack[sp - NT]; c += mystack_sc[sp - NT]; }
        else if (active) {
          const double *src = reinterpret_cast<const double *>(((uint64_t)(uint32_t)sw.w << 32) | (uint32_t)sw.z);
          bv = ld256_stream(src + item_off); c += a.node_sc[iw.w][p];
        }
        double y[4];
        apply_inner<K>(pm_, bv, y);
        const double lk = (((pi0 * av.x) * y[0] + (pi1 * av.y) * y[1]) + (pi2 * av.z) * y[2]) + (pi3 * av.w) * y[3];
        double l = pk_prob * lk;
#pragma unroll
        for (int off = 1; off < K; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
        if (k == 0 && active) {
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
    __syncthreads();
    const double rsum = block_fold_1024(vals, wsum);
    if (tid == 0) a.partials[blk] = rsum;
    __syncthreads();
  }
}

}  // namespace phylo
