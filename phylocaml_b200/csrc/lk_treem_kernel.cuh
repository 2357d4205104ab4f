// Tree-fused pruning for 20 / 61 states on the fp64 tensor cores (BASELINE configs 4 and 5).
//
// One launch evaluates the whole compiled schedule (build_fused_plan, engine.cu). Where the
// per-node path (prune_mma_kernel) streams every child CLV back in from HBM -- 3C + 12 bytes per
// update, 81.5 GB for config 4 -- this kernel keeps the running result of the depth-first walk in
// shared memory and only WRITES each interior CLV once (C + 4 bytes per update, 40.6 GB): the
// value a median needs next is, two times out of three, the one the warp has just produced.
//
// Organisation (unlike the warp-autonomous 4-state kernel the transition matrices are 15-32 KB per
// branch here, so a step's tables must be shared by many pattern groups to be affordable):
//   * one CTA per SM owns a contiguous range of 8-pattern groups and walks it in chunks of NW * R groups
//     (warp w carries groups w, w + NW, ... of the chunk, so a short last chunk still spreads over
//     all warps); every chunk runs the whole program;
//   * the two DMMA A-fragment tables of a step (pt_frag_kernel lays P out as [k][mt][ks][lane]) arrive
//     by bulk-TMA into a double buffer, completion on an mbarrier (SASS UBLKCP + SYNCS). They come out
//     of L2: 3.9 MB for all branches of config 4. There is NO per-step CTA barrier: a warp needs only the
//     step's tables and its own data; the LAST warp to finish with a buffer (shared-memory counter) refills
//     it for the step after next, so a fast warp runs at most one step ahead of the slowest (the bar.sync
//     of the first versions cost 10-20 % in warp-to-warp skew);
//   * the DMMA chains of a group (row tiles x two sides) are independent accumulators advanced
//     together k-step by k-step (non-volatile asm): dependent-issue latency of DMMA is 26 cycles,
//     4 warps x 2 chains saturate the pipe (tools/ubench/dmma_lat.cu);
//   * operand kinds are compile-time inside the step body (the host orders every median so that
//     kind(left) <= kind(right): tip < running value < node slot; x * y is commutative bit for bit):
//     TIP = state masks (one coalesced load per warp and step, prefetched a step ahead). 20 states: the
//     tip side's table is the TRANSPOSE of P plus a row of row sums; an observed state (or a missing
//     cell) is one row, read with one LDS.128 + one LDS.64 per pattern -- no DMMA; a partial ambiguity
//     code (rare, per lane) adds its columns in ascending j, the oracle's own order. 61 states: all-
//     observed groups take column j straight from the A table, others build B fragments from the mask.
//     CUR = the warp's own previous result in shared memory ([k][pattern][PITCH], PITCH = 4 mod 16
//     doubles: conflict-free B-fragment reads). GLB = a retained CLV read back from its node slot
//     (written a few steps earlier by this same warp: an L2 hit). There is no separate stack: every
//     result is retained in its node slot anyway, so "pop" is "read the slot";
//   * results: group-outer loops -- a group's CLV is complete after its K rate classes; 20 states: it
//     leaves at once through ONE cp.async.bulk.tensor.3d (3-D tensor map (S, N, K) per node slot, box
//     (S, 8, K) = the buffer's [k][pattern][S] order; rows >= N clipped), so the drain to HBM overlaps the
//     next group's arithmetic (wait_group.read R - 1 before the buffer is rewritten, wait_group R before
//     a slot written by TMA is read back); 61 states (488-byte rows: no 16-byte strides) store from the
//     C fragments like prune_mma_kernel. Per-site rescaling by the same rule (site maximum over all
//     K * S entries < 2^-256 => times 2^256, counter + 1), scale counters of the running value in registers;
//   * the root-edge join is the program's last step: site lnL (and weight * lnL) go to global
//     memory, the canonical 1024-fold runs over them afterwards (reduce1024_kernel) -- the sum does
//     not depend on how the patterns were cut into groups, CTAs or devices.
// Geometry: 20 states 16 warps x 2 groups (126 registers, A fragments straight from shared memory -- holding
// them in registers with 8 warps x 4 groups was latency-bound at 2 warps per scheduler), 61 states 8 warps x
// 2 groups (the 16 k-steps of B fragments want ~250 registers). profiles/README.md has the measurements
// behind each of these choices (v1 32 ms -> v8 14.2 ms for config 4).
//
// DMMA chains: the k-steps of a row tile are accumulated in ascending order from zero, contraction
// index j = 4 ks + fc (prune_mma_kernel uses a different j <-> slot map for 20 states, so CLVs agree
// with it to rounding, ~1e-16 relative, not bit for bit; tests hold both to the oracle).
#pragma once
#include <type_traits>

#include "lk_kernels.cuh"

namespace phylo {

enum : int { TM_TIP = 0, TM_CUR = 1, TM_GLB = 2 };  // operand modes of this kernel

struct __align__(16) TreeMInstr {
  int kinds;       // lmode | rmode << 2; medians: lmode <= rmode
  int lidx, ridx;  // tip row (TM_TIP) or node slot (TM_GLB)
  int out_slot;    // node slot that receives the result; -1: the root-edge join
};

struct TreeMArgs {
  const TreeMInstr *prog;
  int n_steps;  // medians; step n_steps is the root-edge join
  int K;
  const double *frags;  // [2 n_steps + 1][K][MT][KS][32]: left, right of step 0, 1, ...; last: root edge
  const void *tips;     // MaskT [T][tip_stride]
  int64_t tip_stride, N;
  int64_t g_begin, g_end;  // this launch's 8-pattern groups (a slab of phylo_lk_score_alignment; the whole alignment: 0, ceil(N / 8))
  double *const *node_clv;
  int32_t *const *node_sc;
  const double *pi, *probs, *weights;
  const void *inv;
  double pinvar;
  double *site_lnl, *wsite;
  const char *tmaps;  // 20 states: CUtensorMap per node slot (128 bytes each), dims (S, N, K), box (S, 8, K)
  // measurement only (PHYLO_TREEM_TIMING=1): [variant 0..5][0] cycles of warp 0 inside the step body, [1] cycles
  // it then spent at the step's barrier, [2] steps, [3] cycles of the body spent waiting for the tables; variants: T+T, T+C, T+G, C+G, G+G, root. NULL = off
  unsigned long long *timing;
  int prog_in_smem;  // the launch reserved treem_prog_bytes(n_steps) of shared memory behind the buffers
  int st_swap;       // 20 states: conflict-free order of the result stores (measurement switch PHYLO_TREEM_STSWAP=0)
  int paired;        // 20 states: CLV+CLV steps share every A-fragment read between the warp's two groups (measurement switch PHYLO_TREEM_PAIRED=0)
};

constexpr int treem_pitch(int cols) {
  int p = cols;
  while (p % 16 != 4 && p % 16 != 12) p += 4;
  return p;
}
template <int S>
struct TreeMGeom {
  static constexpr int MT = (S + 7) / 8, KS = (S + 3) / 4, FRAG = MT * KS * 32, PITCH = treem_pitch(KS * 4);
};
template <int S>
__host__ __device__ inline size_t treem_smem_bytes(int K, int R, int NW) {
  using G = TreeMGeom<S>;
  return 128 + sizeof(double) * (4 * (size_t)K * G::FRAG + ((S + 15) & ~15) + (size_t)NW * R * K * 8 * G::PITCH);
}
// + the program itself when it fits behind the above (kernel argument prog_in_smem)
__host__ __device__ inline size_t treem_prog_bytes(int n_steps) { return (size_t)(n_steps + 1) * sizeof(TreeMInstr); }

// P ([branch * K + k][i][j], pt_build_kernel) -> the on-chip tables of the tree kernel, FRAG doubles each.
// DMMA A fragments [mt][ks][lane]: lane (fr = lane / 4, fc = lane % 4) of row tile mt, k-step ks holds
// P[i_of(mt, fr)][4 ks + fc]. For the TIP side of a median with 20 states the table is the transpose
// instead, PT[j][S] = P[0..S-1][j] for j < S, PT[S][.] = the row sums of P (a missing cell), zeros after:
// an observed tip state j contributes column j of P, which a lane then reads as one 128-bit + one 64-bit
// access (rows 2 fr, 2 fr + 1 and 16 + fr; lanes with fr >= 4 read finite values they never store).
template <int S>
__global__ void __launch_bounds__(256) pt_frag_kernel(const double *__restrict__ P, double *__restrict__ frags,
                                                      const TreeMInstr *__restrict__ prog, int n_steps, int K) {
  using G = TreeMGeom<S>;
  const double *src = P + (size_t)blockIdx.x * S * S;
  double *dst = frags + (size_t)blockIdx.x * G::FRAG;
  const int branch = blockIdx.x / K, step = branch >> 1;
  const bool transposed = MmaMap<S>::kVec && step < n_steps && ((prog[step].kinds >> (2 * (branch & 1))) & 3) == TM_TIP;
  for (int idx = threadIdx.x; idx < G::FRAG; idx += blockDim.x) {
    if (transposed) {
      // rows 0 .. S-1: columns of P; row S: the sum of all columns in ascending j (what a missing /
      // fully ambiguous tip contributes -- same additions, same order as the oracle's loop over states)
      const int j = idx / S, i = idx % S;
      double v = 0.0;
      if (j < S) v = src[i * S + j];
      else if (j == S)
        for (int q = 0; q < S; ++q) v += src[i * S + q];
      dst[idx] = v;
    } else {
      const int l = idx & 31, ks = (idx >> 5) % G::KS, mt = (idx >> 5) / G::KS;
      const int i = MmaMap<S>::i_of(mt, l >> 2), j = ks * 4 + (l & 3);
      dst[idx] = (i < S && j < S) ? src[i * S + j] : 0.0;
    }
  }
}

template <typename MaskT>
__device__ __forceinline__ MaskT shfl_mask(MaskT v, int src) {
  if constexpr (sizeof(MaskT) == 8) return (MaskT)__shfl_sync(0xffffffffu, (unsigned long long)v, src);
  else return (MaskT)__shfl_sync(0xffffffffu, (unsigned)v, src);
}
// the same with a 32-bit shared-window address (one cvta per kernel instead of one per access)
__device__ __forceinline__ void sts128_s(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void lds128_s(uint32_t addr, double &a, double &b) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts128(double *p, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(smem_u32(p)), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void lds128(const double *p, double &a, double &b) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void tma_store_3d(const void *tmap, const void *smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// not volatile: a pure function of its operands, so independent chains may be interleaved
__device__ __forceinline__ void dmma_acc(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

template <int S, typename MaskT, int R, int NW, int KT>
__global__ void __launch_bounds__(NW * 32, 1) lk_treem_kernel(const TreeMArgs a) {
  using G = TreeMGeom<S>;
  using Map = MmaMap<S>;
  constexpr int MT = G::MT, KS = G::KS, FRAG = G::FRAG, PITCH = G::PITCH;
  constexpr bool PT = Map::kVec;        // tip sides read the transposed table (20 states)
  constexpr bool TMA = Map::kVec;       // results leave by TMA tensor stores from the running-value buffer (16-byte rows)
  constexpr int MTB = Map::kVec ? MT : (MT % 4 == 0 ? 4 : (MT % 2 == 0 ? 2 : 1));  // row tiles in flight together
  static_assert(!Map::kVec || MT == 3, "the 128-bit store path expects three row tiles");
  static_assert(!Map::kVec || (S + 1) * S + 4 <= FRAG, "the transposed tip table (S columns + the row sums) must fit a table slot");
  constexpr unsigned FULL = 0xffffffffu;
  static_assert(R >= 1 && R <= 4, "a warp's tip masks travel in one register per side: R * 8 <= 32");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = KT ? KT : a.K;
  uint64_t *bar = (uint64_t *)smem_raw;
  unsigned int *done = (unsigned int *)(smem_raw + 16);  // [2]: warps that have finished with table buffer b
  double *pbuf = (double *)(smem_raw + 128);
  const int side = K * FRAG;  // doubles per side
  double *spi = pbuf + 4 * side;
  double *curbase = spi + ((S + 15) & ~15);  // 128-byte aligned: the groups' running CLVs are TMA store sources
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  // st_order: a lane stores its two pattern rows (2 fc, 2 fc + 1 of the dense [k][8][PITCH] TMA source) with two
  // 128-bit stores. In row order the eight lanes of a quarter warp hit 16-byte bank groups (4 fc + fr) mod 8:
  // fc = 0 / 2 and fc = 1 / 3 collide (2-way conflicts: 815 M extra wavefronts in r02_ncu_treem_aa_v9.txt). With the
  // lanes fc >= 2 storing the odd row first the groups are 0, 4, 2, 6 (+ fr), then 2, 6, 0, 4: conflict-free, same
  // bytes to the same addresses.
  const bool st_swap = fc >= 2 && a.st_swap;
  const uint32_t st_first = st_swap ? (uint32_t)PITCH * 8u : 0u, st_second = (uint32_t)PITCH * 8u - st_first;
  const int gsz = K * 8 * PITCH;  // doubles of one group's running CLV
  for (int idx = tid; idx < NW * R * gsz; idx += NW * 32) curbase[idx] = 0.0;  // pad columns stay 0
  for (int i = tid; i < S; i += NW * 32) spi[i] = a.pi[i];
  // the six node-slot pointers a step may need (left / right / out CLV, then their scale arrays): fetched by
  // lanes 0-5 of every warp one step ahead and broadcast by shuffles at the head of the step, so no warp
  // waits on the pointer tables (global memory) there
  auto slot_ptr = [&](const TreeMInstr &in) -> unsigned long long {
    const int which = lane % 3;
    const int idx = which == 0 ? in.lidx : (which == 1 ? in.ridx : in.out_slot);
    const bool glb = which == 0 ? (in.kinds & 3) == TM_GLB : (which == 1 ? ((in.kinds >> 2) & 3) == TM_GLB : in.out_slot >= 0);
    if (!glb) return 0ull;
    return lane < 3 ? (unsigned long long)a.node_clv[idx] : (unsigned long long)a.node_sc[idx];
  };
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    done[0] = done[1] = 0;
    fence_mbar_init();
  }
  // the program is read three times per step (operands, next step's masks, next step's pointers): a copy
  // in shared memory takes two dependent global loads off the head of every step
  const TreeMInstr *prog = a.prog;
  if (a.prog_in_smem) {
    TreeMInstr *sprog = (TreeMInstr *)(curbase + NW * R * gsz);
    for (int i = tid; i <= a.n_steps; i += NW * 32) sprog[i] = a.prog[i];
    prog = sprog;
  }
  unsigned long long cur_ptr = lane < 6 ? slot_ptr(a.prog[0]) : 0ull;
  __syncthreads();

  const int64_t ngroups = a.g_end - a.g_begin;
  const int64_t g0 = a.g_begin + ngroups * blockIdx.x / gridDim.x, g1 = a.g_begin + ngroups * (blockIdx.x + 1) / gridDim.x;
  const int nst = a.n_steps + 1;
  constexpr int CH = NW * R;
  const int64_t nchunks = (g1 - g0 + CH - 1) / CH;
  const uint64_t total_it = (uint64_t)nchunks * nst;
  const uint32_t side_bytes = (uint32_t)side * 8u;
  auto issue = [&](uint64_t it) {  // thread 0: request the tables of global step `it`
    const int s = (int)(it % nst), buf = (int)(it & 1);
    fence_proxy_async();
    const bool root = s == a.n_steps;
    mbar_expect_tx(&bar[buf], root ? side_bytes : 2 * side_bytes);
    const double *src = a.frags + (size_t)(2 * s) * side;
    bulk_g2s(pbuf + buf * 2 * side, src, side_bytes, &bar[buf]);
    if (!root) bulk_g2s(pbuf + buf * 2 * side + side, src + side, side_bytes, &bar[buf]);
  };
  // The warps of a CTA are NOT barrier-synchronised per step: a warp only needs this step's tables (mbarrier)
  // and its own data. A table buffer is refilled (for the step after next) by whichever warp is the LAST to
  // finish with it, so a fast warp runs at most one step ahead of the slowest -- the per-step bar.sync of the
  // first versions cost 10-20 % in warp-to-warp skew (node-slot reads, tensor-pipe contention).
  if (tid == 0) {
    if (total_it > 0) issue(0);
    if (total_it > 1) issue(1);
  }

  const MaskT keep = (S >= 64) ? ~(MaskT)0 : (MaskT)(((uint64_t)1 << S) - 1);
  const MaskT *tips = (const MaskT *)a.tips;
  double *cur = curbase + warp * R * gsz;  // [r][k][8][PITCH]
  const uint32_t cur_s = smem_u32(cur), pbuf_s = smem_u32(pbuf);
  const double kLnScale = kScaleExp * 0.6931471805599453094;
  uint64_t it = 0;
  for (int64_t chunk = 0; chunk < nchunks; ++chunk) {
    const int64_t cbase = g0 + chunk * CH;
    // groups cbase + r * NW + warp, r < nact, are this warp's; only the alignment's last group can be ragged
    int nact = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) nact += (cbase + (int64_t)r * NW + warp < g1) ? 1 : 0;
    // lane l < R * 8 fetches the mask of pattern l % 8 of this warp's group l / 8
    const int64_t mg = cbase + (int64_t)(lane >> 3) * NW + warp;
    const int64_t mp = mg * 8 + (lane & 7);
    const bool mok = lane < R * 8 && mg < g1 && mp < a.N;
    auto load_masks = [&](const TreeMInstr &in, MaskT &ml, MaskT &mr) {  // raw: `& keep` happens at the use
      ml = 0;
      mr = 0;
      if ((in.kinds & 3) == TM_TIP && mok) ml = tips[(size_t)in.lidx * a.tip_stride + mp];
      if (((in.kinds >> 2) & 3) == TM_TIP && mok) mr = tips[(size_t)in.ridx * a.tip_stride + mp];
    };
    TreeMInstr ins = prog[0];
    MaskT ml, mr;
    load_masks(ins, ml, mr);
    ml &= keep;
    mr &= keep;
    int csc0[R], csc1[R];  // scale counters of the running values (patterns 2 fc, 2 fc + 1 of group r)
#pragma unroll
    for (int r = 0; r < R; ++r) csc0[r] = csc1[r] = 0;

#pragma unroll 1
    for (int s = 0; s < nst; ++s, ++it) {
      TreeMInstr nins = ins;
      MaskT nml = 0, nmr = 0;
      unsigned long long next_ptr = 0;
      if (s + 1 < nst) {
        nins = prog[s + 1];
        load_masks(nins, nml, nmr);
        if (lane < 6) next_ptr = slot_ptr(nins);
      } else if (lane < 6 && it + 1 < total_it) {
        next_ptr = slot_ptr(prog[0]);
      }
      if constexpr (TMA) {  // this step's output descriptor: in the descriptor cache by the time the stores are issued
        if (lane == 0 && ins.out_slot >= 0)
          asm volatile("prefetch.tensormap [%0];" ::"l"(a.tmaps + (size_t)ins.out_slot * 128) : "memory");
      }
      const int lk = ins.kinds & 3, rk = (ins.kinds >> 2) & 3;
      const long long t_begin = a.timing ? clock64() : 0;
      const double *lg = (const double *)__shfl_sync(FULL, cur_ptr, 0), *rg = (const double *)__shfl_sync(FULL, cur_ptr, 1);
      double *og = (double *)__shfl_sync(FULL, cur_ptr, 2);
      const int32_t *lgs = (const int32_t *)__shfl_sync(FULL, cur_ptr, 3), *rgs = (const int32_t *)__shfl_sync(FULL, cur_ptr, 4);
      int32_t *ogs = (int32_t *)__shfl_sync(FULL, cur_ptr, 5);
      const unsigned lhotbits = __ballot_sync(FULL, (ml & (ml - 1)) == 0);
      const unsigned rhotbits = __ballot_sync(FULL, (mr & (mr - 1)) == 0);
      const int buf = (int)(it & 1);
      const double *fl = pbuf + buf * 2 * side, *frg = fl + side;
      const long long t_w0 = a.timing ? clock64() : 0;
      mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));
      const long long t_w1 = a.timing ? clock64() : 0;

      // B fragment of an operand for rate class k of group r (lane = pattern fr, columns 4 ks + fc)
      auto fetch_cur = [&](double (&b)[KS], int r, int k) {
        const double *row = cur + r * gsz + (k * 8 + fr) * PITCH + fc;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) b[ks] = row[ks * 4];
      };
      auto fetch_glb = [&](double (&b)[KS], const double *g, int k, int64_t pbase) {
        const bool ok = pbase + fr < a.N;
        const double *row = g + ((size_t)(pbase + fr) * K + k) * S + fc;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) b[ks] = (ok && ks * 4 + fc < S) ? row[ks * 4] : 0.0;
      };
      // column of the A table that holds P[.][j] for this lane's row slot: + mt * KS * 32
      auto hot_col = [&](MaskT m) {
        int j;
        if constexpr (sizeof(MaskT) == 8) j = m ? __ffsll((long long)m) - 1 : 0;
        else j = m ? __ffs((int)m) - 1 : 0;
        return (j >> 2) * 32 + fr * 4 + (j & 3);
      };

      if (s < a.n_steps) {
        // ------------------------------------------------------------------ a median
        auto median = [&](auto LMc, auto RMc) {
          constexpr int LM = decltype(LMc)::value, RM = decltype(RMc)::value;
          if constexpr (TMA) {
            // a node-slot operand was stored two or more steps ago: everything but the previous step's R bulk
            // groups must have landed before it is read back
            if (LM == TM_GLB || RM == TM_GLB) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group %0;" ::"n"(R) : "memory");
              __syncwarp();
            }
          }
          if (PT && R == 2 && LM != TM_TIP && RM != TM_TIP && a.paired) {
            // ---- 20 states, both operands CLVs: rate classes outer, the warp's two groups inner. Every A fragment
            // read from shared memory feeds the DMMAs of both groups: these steps were bound by shared-memory traffic
            // (LSU 78 % busy, 60 of ~85 wavefronts per (rate class, group) were A fragments), not by the tensor
            // pipe. Both CLVs are complete after the last rate class; their tensor stores are read out by the TMA
            // unit while the next step's first rate class is computed (wait_group.read sits just before that step's
            // first store into the running-value buffer). Measured (PHYLO_TREEM_TIMING): CLV+CLV steps 22.7 k ->
            // 19.6 k cycles per chunk; steps with a tip side lose more from the late stores than they gain (tip+tip
            // 7.7 k -> 11.3 k, tip+CLV 12.0 k -> 13.9 k) and keep the group-outer order below.
            constexpr int R2 = 2;
            int64_t pbase2[R2], pa02[R2];
            bool ok0[R2], ok1[R2], act[R2];
            int hh0[R2], hh1[R2], gsa[R2], gsb[R2];
            unsigned mL0[R2], mL1[R2], mR0[R2], mR1[R2];
#pragma unroll
            for (int r = 0; r < R2; ++r) {
              act[r] = r < nact;
              pbase2[r] = (cbase + (int64_t)r * NW + warp) * 8;
              pa02[r] = pbase2[r] + 2 * fc;
              ok0[r] = act[r] && pa02[r] < a.N;
              ok1[r] = act[r] && pa02[r] + 1 < a.N;
              hh0[r] = hh1[r] = (int)0x80000000;
              gsa[r] = gsb[r] = 0;
              mL0[r] = mL1[r] = mR0[r] = mR1[r] = 0;
              if (LM == TM_TIP) {
                mL0[r] = (unsigned)shfl_mask(ml, r * 8 + 2 * fc);
                mL1[r] = (unsigned)shfl_mask(ml, r * 8 + 2 * fc + 1);
              }
              if (RM == TM_TIP) {
                mR0[r] = (unsigned)shfl_mask(mr, r * 8 + 2 * fc);
                mR1[r] = (unsigned)shfl_mask(mr, r * 8 + 2 * fc + 1);
              }
              if (RM == TM_GLB) {
                if (ok0[r]) gsa[r] += rgs[pa02[r]];
                if (ok1[r]) gsb[r] += rgs[pa02[r] + 1];
              }
              if (LM == TM_GLB) {
                if (ok0[r]) gsa[r] += lgs[pa02[r]];
                if (ok1[r]) gsb[r] += lgs[pa02[r] + 1];
              }
            }
            auto tip_row = [&](unsigned mk) {
              return mk == (unsigned)keep ? S : ((mk & (mk - 1)) == 0 ? (mk ? __ffs((int)mk) - 1 : 0) : -1);
            };
            auto tip_side = [&](const double *pt, unsigned mk0, unsigned mk1, double (&c)[MT][2]) {
              const int j0 = tip_row(mk0), j1 = tip_row(mk1);
              if (j0 >= 0 && j1 >= 0) {
                const uint32_t pt_s = pbuf_s + (uint32_t)(pt - pbuf) * 8u;
                lds128_s(pt_s + (uint32_t)(j0 * S + 2 * fr) * 8u, c[0][0], c[1][0]);
                lds128_s(pt_s + (uint32_t)(j1 * S + 2 * fr) * 8u, c[0][1], c[1][1]);
                c[2][0] = pt[j0 * S + 16 + fr];
                c[2][1] = pt[j1 * S + 16 + fr];
              } else {
#pragma unroll 1
                for (int j = 0; j < S; ++j) {
                  const double *col = pt + j * S;
                  if ((mk0 >> j) & 1) { c[0][0] += col[2 * fr]; c[1][0] += col[2 * fr + 1]; c[2][0] += col[16 + fr]; }
                  if ((mk1 >> j) & 1) { c[0][1] += col[2 * fr]; c[1][1] += col[2 * fr + 1]; c[2][1] += col[16 + fr]; }
                }
              }
            };
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
              const double *tabL = fl + k * FRAG, *tabR = frg + k * FRAG;
              double cx[R2][MT][2], cy[R2][MT][2];
#pragma unroll
              for (int r = 0; r < R2; ++r)
#pragma unroll
                for (int m = 0; m < MT; ++m) cx[r][m][0] = cx[r][m][1] = cy[r][m][0] = cy[r][m][1] = 0.0;
              if constexpr (LM == TM_TIP) {
#pragma unroll
                for (int r = 0; r < R2; ++r)
                  if (act[r]) tip_side(tabL, mL0[r], mL1[r], cx[r]);
              }
              if constexpr (RM == TM_TIP) {
#pragma unroll
                for (int r = 0; r < R2; ++r)
                  if (act[r]) tip_side(tabR, mR0[r], mR1[r], cy[r]);
              }
              if constexpr (LM != TM_TIP || RM != TM_TIP) {
                // operand rows: CUR in shared memory ([k][pattern fr][PITCH] + fc), GLB in the node slot
                const double *rowL[R2], *rowR[R2];
                bool gokL[R2], gokR[R2];
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                  gokL[r] = gokR[r] = act[r];
                  rowL[r] = rowR[r] = cur + r * gsz + (k * 8 + fr) * PITCH + fc;
                  if constexpr (LM == TM_GLB) {
                    gokL[r] = act[r] && pbase2[r] + fr < a.N;
                    rowL[r] = lg + ((size_t)(pbase2[r] + fr) * K + k) * S + fc;
                  }
                  if constexpr (RM == TM_GLB) {
                    gokR[r] = act[r] && pbase2[r] + fr < a.N;
                    rowR[r] = rg + ((size_t)(pbase2[r] + fr) * K + k) * S + fc;
                  }
                }
                double bl[R2][KS], br[R2][KS];
#pragma unroll
                for (int r = 0; r < R2; ++r)
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks) {
                    if constexpr (LM == TM_CUR) bl[r][ks] = rowL[r][ks * 4];
                    if constexpr (LM == TM_GLB) bl[r][ks] = (gokL[r] && ks * 4 + fc < S) ? rowL[r][ks * 4] : 0.0;
                    if constexpr (RM == TM_CUR) br[r][ks] = rowR[r][ks * 4];
                    if constexpr (RM == TM_GLB) br[r][ks] = (gokR[r] && ks * 4 + fc < S) ? rowR[r][ks * 4] : 0.0;
                  }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                  for (int m = 0; m < MT; ++m) {
                    if constexpr (LM != TM_TIP) {
                      const double av = tabL[(m * KS + ks) * 32 + lane];
#pragma unroll
                      for (int r = 0; r < R2; ++r) dmma_acc(cx[r][m], av, bl[r][ks]);
                    }
                    if constexpr (RM != TM_TIP) {
                      const double av = tabR[(m * KS + ks) * 32 + lane];
#pragma unroll
                      for (int r = 0; r < R2; ++r) dmma_acc(cy[r][m], av, br[r][ks]);
                    }
                  }
                }
              }
              if (k == 0) {
                // the previous step's two tensor stores must have read these buffers before they are rewritten
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              __syncwarp();  // (and every lane has read this rate class of the running values)
#pragma unroll
              for (int r = 0; r < R2; ++r) {
                if (act[r]) {
                  const double a0 = cx[r][0][0] * cy[r][0][0], a1 = cx[r][0][1] * cy[r][0][1];
                  const double b0 = cx[r][1][0] * cy[r][1][0], b1 = cx[r][1][1] * cy[r][1][1];
                  const double d0 = cx[r][2][0] * cy[r][2][0], d1 = cx[r][2][1] * cy[r][2][1];
                  const uint32_t c0_s = cur_s + (uint32_t)(r * gsz + (k * 8 + 2 * fc) * PITCH + 2 * fr) * 8u;
                  // lanes with fc >= 2 store their odd pattern row first (see st_order below): conflict-free banks
                  sts128_s(c0_s + st_first, st_swap ? a1 : a0, st_swap ? b1 : b0);
                  sts128_s(c0_s + st_second, st_swap ? a0 : a1, st_swap ? b0 : b1);
                  hh0[r] = max(hh0[r], max(hi32(a0), hi32(b0)));
                  hh1[r] = max(hh1[r], max(hi32(a1), hi32(b1)));
                  if (fr < 4) {
                    double *c0 = cur + r * gsz + (k * 8 + 2 * fc) * PITCH;
                    c0[(st_first >> 3) + 16 + fr] = st_swap ? d1 : d0;
                    c0[(st_second >> 3) + 16 + fr] = st_swap ? d0 : d1;
                    hh0[r] = max(hh0[r], hi32(d0));
                    hh1[r] = max(hh1[r], hi32(d1));
                  }
                }
              }
            }
#pragma unroll
            for (int r = 0; r < R2; ++r) {
              if (act[r]) {
                int h0 = hh0[r], h1 = hh1[r];
#pragma unroll
                for (int off = 4; off <= 16; off <<= 1) {
                  h0 = max(h0, __shfl_xor_sync(FULL, h0, off));
                  h1 = max(h1, __shfl_xor_sync(FULL, h1, off));
                }
                const bool r0 = ok0[r] && h0 < kScaleHiThresh, r1 = ok1[r] && h1 < kScaleHiThresh;
                if (r0 || r1) {  // rare: every lane rescales what it stored; rolled loops
#pragma unroll 1
                  for (int k = 0; k < K; ++k) {
                    double *c0 = cur + r * gsz + (k * 8 + 2 * fc) * PITCH, *c1 = c0 + PITCH;
#pragma unroll 1
                    for (int mt = 0; mt < MT; ++mt) {
                      const int i = Map::i_of(mt, fr);
                      if (i < S) {
                        if (r0) c0[i] *= 0x1p+256;
                        if (r1) c1[i] *= 0x1p+256;
                      }
                    }
                  }
                }
                int s0 = gsa[r], s1 = gsb[r];
                if (LM == TM_CUR || RM == TM_CUR) { s0 += csc0[r]; s1 += csc1[r]; }
                s0 += r0 ? 1 : 0;
                s1 += r1 ? 1 : 0;
                csc0[r] = s0;
                csc1[r] = s1;
                if (fr == 0) {
                  if (ok0[r]) ogs[pa02[r]] = s0;
                  if (ok1[r]) ogs[pa02[r] + 1] = s1;
                }
                // the group's finished CLV [k][8][S] -> node slot [pattern][k][S]: one tensor store (rows >= N clipped)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) tma_store_3d(a.tmaps + (size_t)ins.out_slot * 128, cur + r * gsz, 0, (int)pbase2[r], 0);
              }
              if (lane == 0) bulk_commit();  // one bulk group per (step, r), empty for an idle r: the waits count groups
            }
            return;
          }
          // group by group (r outer, rate classes inner): a group's CLV is complete after its K iterations and
          // leaves at once (TMA), so the drain to HBM overlaps the next group's arithmetic
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (r < nact) {
              const int64_t pbase = (cbase + (int64_t)r * NW + warp) * 8, pa0 = pbase + 2 * fc;
              const bool pa0_ok = pa0 < a.N, pa1_ok = pa0 + 1 < a.N;
              int h0 = (int)0x80000000, h1 = (int)0x80000000;
              int colL0 = 0, colL1 = 0, colR0 = 0, colR1 = 0;  // PT: the MASKS of the lane's two result patterns
              MaskT mlb = 0, mrb = 0;
              bool lhot = false, rhot = false;
              int gs0 = 0, gs1 = 0;  // summed scale counters of the node-slot operands (loaded early, used late)
              if (LM == TM_TIP) {
                lhot = ((lhotbits >> (r * 8)) & 0xffu) == 0xffu;
                if constexpr (PT) {
                  colL0 = (int)shfl_mask(ml, r * 8 + 2 * fc);
                  colL1 = (int)shfl_mask(ml, r * 8 + 2 * fc + 1);
                } else if (lhot) {
                  colL0 = hot_col(shfl_mask(ml, r * 8 + 2 * fc));
                  colL1 = hot_col(shfl_mask(ml, r * 8 + 2 * fc + 1));
                } else {
                  mlb = shfl_mask(ml, r * 8 + fr);
                }
              }
              if (RM == TM_TIP) {
                rhot = ((rhotbits >> (r * 8)) & 0xffu) == 0xffu;
                if constexpr (PT) {
                  colR0 = (int)shfl_mask(mr, r * 8 + 2 * fc);
                  colR1 = (int)shfl_mask(mr, r * 8 + 2 * fc + 1);
                } else if (rhot) {
                  colR0 = hot_col(shfl_mask(mr, r * 8 + 2 * fc));
                  colR1 = hot_col(shfl_mask(mr, r * 8 + 2 * fc + 1));
                } else {
                  mrb = shfl_mask(mr, r * 8 + fr);
                }
              }
              if (RM == TM_GLB) {
                if (pa0_ok) gs0 += rgs[pa0];
                if (pa1_ok) gs1 += rgs[pa0 + 1];
              }
              if (LM == TM_GLB) {
                if (pa0_ok) gs0 += lgs[pa0];
                if (pa1_ok) gs1 += lgs[pa0 + 1];
              }
              if constexpr (TMA) {
                // the tensor store that last read this group's buffer (previous step, same r) is R - 1 groups back
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(R - 1) : "memory");
                __syncwarp();
              }
#pragma unroll 1
              for (int k = 0; k < K; ++k) {
                const double *tabL = fl + k * FRAG, *tabR = frg + k * FRAG;
                double bl[LM != TM_TIP ? KS : 1], br[RM != TM_TIP ? KS : 1];
                if constexpr (LM == TM_CUR) fetch_cur(bl, r, k);
                if constexpr (LM == TM_GLB) fetch_glb(bl, lg, k, pbase);
                if constexpr (RM == TM_CUR) fetch_cur(br, r, k);
                if constexpr (RM == TM_GLB) fetch_glb(br, rg, k, pbase);
                if (LM == TM_CUR || RM == TM_CUR) __syncwarp();  // the result overwrites the running value just read
                double *c0 = cur + r * gsz + (k * 8 + 2 * fc) * PITCH, *c1 = c0 + PITCH;
                double *o0 = og + ((size_t)pa0 * K + k) * S, *o1 = o0 + (size_t)K * S;
                // MTB row tiles of both sides advance together: up to 2 MTB independent accumulator chains
#pragma unroll
                for (int m0 = 0; m0 < MT; m0 += MTB) {
                  double cx[MTB][2], cy[MTB][2];
#pragma unroll
                  for (int m = 0; m < MTB; ++m) cx[m][0] = cx[m][1] = cy[m][0] = cy[m][1] = 0.0;
                  if constexpr (PT) {
                    // tip side, 20 states: a row of the transposed table (an observed state's column of P, or the
                    // row sums for a missing cell); a partial ambiguity code (rare, per lane) adds its columns in
                    // ascending j -- the oracle's own order
                    auto tip_row = [&](unsigned mk) {
                      return mk == (unsigned)keep ? S : ((mk & (mk - 1)) == 0 ? (mk ? __ffs((int)mk) - 1 : 0) : -1);
                    };
                    auto tip_side = [&](const double *pt, unsigned mk0, unsigned mk1, double (&c)[MTB][2]) {
                      const int j0 = tip_row(mk0), j1 = tip_row(mk1);
                      if (j0 >= 0 && j1 >= 0) {
                        const uint32_t pt_s = pbuf_s + (uint32_t)(pt - pbuf) * 8u;
                        lds128_s(pt_s + (uint32_t)(j0 * S + 2 * fr) * 8u, c[0][0], c[1][0]);
                        lds128_s(pt_s + (uint32_t)(j1 * S + 2 * fr) * 8u, c[0][1], c[1][1]);
                        c[2][0] = pt[j0 * S + 16 + fr];
                        c[2][1] = pt[j1 * S + 16 + fr];
                      } else {
#pragma unroll 1
                        for (int j = 0; j < S; ++j) {
                          const double *col = pt + j * S;
                          if ((mk0 >> j) & 1) { c[0][0] += col[2 * fr]; c[1][0] += col[2 * fr + 1]; c[2][0] += col[16 + fr]; }
                          if ((mk1 >> j) & 1) { c[0][1] += col[2 * fr]; c[1][1] += col[2 * fr + 1]; c[2][1] += col[16 + fr]; }
                        }
                      }
                    };
                    if constexpr (LM == TM_TIP) tip_side(tabL, (unsigned)colL0, (unsigned)colL1, cx);
                    if constexpr (RM == TM_TIP) tip_side(tabR, (unsigned)colR0, (unsigned)colR1, cy);
                  } else {
#pragma unroll
                    for (int m = 0; m < MTB; ++m) {
                      if (lhot) {
                        cx[m][0] = tabL[(m0 + m) * KS * 32 + colL0];
                        cx[m][1] = tabL[(m0 + m) * KS * 32 + colL1];
                      }
                      if (rhot) {
                        cy[m][0] = tabR[(m0 + m) * KS * 32 + colR0];
                        cy[m][1] = tabR[(m0 + m) * KS * 32 + colR1];
                      }
                    }
                    // a tip with an ambiguous / missing pattern in this group (rare): fragments from the mask bits,
                    // A from the table; a rolled loop, so that it stays a branch around and not predicated code
                    if (LM == TM_TIP && !lhot) {
#pragma unroll 1
                      for (int ks = 0; ks < KS; ++ks) {
                        const double b = ((mlb >> (ks * 4 + fc)) & 1) ? 1.0 : 0.0;
#pragma unroll
                        for (int m = 0; m < MTB; ++m) dmma_acc(cx[m], tabL[((m0 + m) * KS + ks) * 32 + lane], b);
                      }
                    }
                    if (RM == TM_TIP && !rhot) {
#pragma unroll 1
                      for (int ks = 0; ks < KS; ++ks) {
                        const double b = ((mrb >> (ks * 4 + fc)) & 1) ? 1.0 : 0.0;
#pragma unroll
                        for (int m = 0; m < MTB; ++m) dmma_acc(cy[m], tabR[((m0 + m) * KS + ks) * 32 + lane], b);
                      }
                    }
                  }
                  if constexpr (LM != TM_TIP || RM != TM_TIP) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                      if constexpr (LM != TM_TIP) {
#pragma unroll
                        for (int m = 0; m < MTB; ++m) dmma_acc(cx[m], tabL[((m0 + m) * KS + ks) * 32 + lane], bl[ks]);
                      }
                      if constexpr (RM != TM_TIP) {
#pragma unroll
                        for (int m = 0; m < MTB; ++m) dmma_acc(cy[m], tabR[((m0 + m) * KS + ks) * 32 + lane], br[ks]);
                      }
                    }
                  }
                  if constexpr (Map::kVec) {
                    const double a0 = cx[0][0] * cy[0][0], a1 = cx[0][1] * cy[0][1];
                    const double b0 = cx[1][0] * cy[1][0], b1 = cx[1][1] * cy[1][1];
                    const double d0 = cx[2][0] * cy[2][0], d1 = cx[2][1] * cy[2][1];
                    const uint32_t c0_s = cur_s + (uint32_t)(r * gsz + (k * 8 + 2 * fc) * PITCH + 2 * fr) * 8u;
                    sts128_s(c0_s + st_first, st_swap ? a1 : a0, st_swap ? b1 : b0);
                    sts128_s(c0_s + st_second, st_swap ? a0 : a1, st_swap ? b0 : b1);
                    h0 = max(h0, max(hi32(a0), hi32(b0)));
                    h1 = max(h1, max(hi32(a1), hi32(b1)));
                    if (fr < 4) {
                      c0[(st_first >> 3) + 16 + fr] = st_swap ? d1 : d0;
                      c0[(st_second >> 3) + 16 + fr] = st_swap ? d0 : d1;
                      h0 = max(h0, hi32(d0));
                      h1 = max(h1, hi32(d1));
                    }
                    if constexpr (!TMA) {
                      if (pa0_ok) {
                        st128(o0 + 2 * fr, a0, b0);
                        if (fr < 4) o0[16 + fr] = d0;
                      }
                      if (pa1_ok) {
                        st128(o1 + 2 * fr, a1, b1);
                        if (fr < 4) o1[16 + fr] = d1;
                      }
                    }
                  } else {
#pragma unroll
                    for (int m = 0; m < MTB; ++m) {
                      const double v0 = cx[m][0] * cy[m][0], v1 = cx[m][1] * cy[m][1];
                      const int i = Map::i_of(m0 + m, fr);
                      if (i < S) {
                        c0[i] = v0;
                        c1[i] = v1;
                        h0 = max(h0, hi32(v0));
                        h1 = max(h1, hi32(v1));
                        if (pa0_ok) o0[i] = v0;
                        if (pa1_ok) o1[i] = v1;
                      }
                    }
                  }
                }
              }
              // per-site rescaling and scale counters
#pragma unroll
              for (int off = 4; off <= 16; off <<= 1) {
                h0 = max(h0, __shfl_xor_sync(FULL, h0, off));
                h1 = max(h1, __shfl_xor_sync(FULL, h1, off));
              }
              const bool r0 = pa0_ok && h0 < kScaleHiThresh, r1 = pa1_ok && h1 < kScaleHiThresh;
              if (r0 || r1) {  // rare: every lane rescales what it stored (both copies); rolled loops
#pragma unroll 1
                for (int k = 0; k < K; ++k) {
                  double *c0 = cur + r * gsz + (k * 8 + 2 * fc) * PITCH, *c1 = c0 + PITCH;
                  double *o0 = og + ((size_t)pa0 * K + k) * S, *o1 = o0 + (size_t)K * S;
#pragma unroll 1
                  for (int mt = 0; mt < MT; ++mt) {
                    const int i = Map::i_of(mt, fr);
                    if (i < S) {
                      if (r0) { c0[i] *= 0x1p+256; if (!TMA) o0[i] = c0[i]; }
                      if (r1) { c1[i] *= 0x1p+256; if (!TMA) o1[i] = c1[i]; }
                    }
                  }
                }
              }
              int s0 = gs0, s1 = gs1;
              if (LM == TM_CUR || RM == TM_CUR) { s0 += csc0[r]; s1 += csc1[r]; }
              s0 += r0 ? 1 : 0;
              s1 += r1 ? 1 : 0;
              csc0[r] = s0;
              csc1[r] = s1;
              if (fr == 0) {
                if (pa0_ok) ogs[pa0] = s0;
                if (pa1_ok) ogs[pa0 + 1] = s1;
              }
              if constexpr (TMA) {
                // the group's finished CLV [k][8][S] -> node slot [pattern][k][S]: one tensor store (rows >= N clipped)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) tma_store_3d(a.tmaps + (size_t)ins.out_slot * 128, cur + r * gsz, 0, (int)pbase, 0);
              }
            }
            if constexpr (TMA) {
              if (lane == 0) bulk_commit();  // one bulk group per (step, r), empty for an idle r: the waits count groups
            }
          }
        };
        using I0 = std::integral_constant<int, TM_TIP>;
        using I1 = std::integral_constant<int, TM_CUR>;
        using I2 = std::integral_constant<int, TM_GLB>;
        switch (ins.kinds & 15) {  // lmode | rmode << 2, lmode <= rmode (host-normalised)
          case TM_TIP | (TM_TIP << 2): median(I0{}, I0{}); break;
          case TM_TIP | (TM_CUR << 2): median(I0{}, I1{}); break;
          case TM_TIP | (TM_GLB << 2): median(I0{}, I2{}); break;
          case TM_CUR | (TM_GLB << 2): median(I1{}, I2{}); break;
          default: median(I2{}, I2{}); break;
        }
      } else {
        // ------------------------------------------------------ the root-edge join (a = left, b = right)
        const bool ltip = lk == TM_TIP, rtip = rk == TM_TIP, lcur = lk == TM_CUR, rcur = rk == TM_CUR;
        const MaskT *inv = (const MaskT *)a.inv;
        if constexpr (TMA) {
          if (lane == 0) bulk_wait0();
          __syncwarp();
        }
#pragma unroll 1
        for (int r = 0; r < nact; ++r) {
          const int64_t pbase = (cbase + (int64_t)r * NW + warp) * 8, pa0 = pbase + 2 * fc;
          const bool pa0_ok = pa0 < a.N, pa1_ok = pa0 + 1 < a.N;
          const bool bhot = rtip && ((rhotbits >> (r * 8)) & 0xffu) == 0xffu;
          int col_b0 = 0, col_b1 = 0;
          MaskT mbb = 0, ma0 = 0, ma1 = 0;
          if (rtip) {
            if (bhot) {
              col_b0 = hot_col(shfl_mask(mr, r * 8 + 2 * fc));
              col_b1 = hot_col(shfl_mask(mr, r * 8 + 2 * fc + 1));
            } else {
              mbb = shfl_mask(mr, r * 8 + fr);
            }
          }
          if (ltip) {
            ma0 = shfl_mask(ml, r * 8 + 2 * fc);
            ma1 = shfl_mask(ml, r * 8 + 2 * fc + 1);
          }
          const int sca0 = (r == 0 ? csc0[0] : r == 1 ? csc0[R > 1 ? 1 : 0] : r == 2 ? csc0[R > 2 ? 2 : 0] : csc0[R > 3 ? 3 : 0]);
          const int sca1 = (r == 0 ? csc1[0] : r == 1 ? csc1[R > 1 ? 1 : 0] : r == 2 ? csc1[R > 2 ? 2 : 0] : csc1[R > 3 ? 3 : 0]);
          double l0 = 0.0, l1 = 0.0;
#pragma unroll 1
          for (int k = 0; k < K; ++k) {
            const double *fk0 = fl + k * FRAG;
            double bb[KS];
            if (!bhot) {
              if (rtip) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) bb[ks] = ((mbb >> (ks * 4 + fc)) & 1) ? 1.0 : 0.0;
              } else if (rcur) {
                fetch_cur(bb, r, k);
              } else {
                fetch_glb(bb, rg, k, pbase);
              }
            }
            const double *c0 = cur + r * gsz + (k * 8 + 2 * fc) * PITCH, *c1 = c0 + PITCH;
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 2
            for (int mt = 0; mt < MT; ++mt) {
              double cy[2] = {0.0, 0.0};
              if (bhot) {
                cy[0] = fk0[mt * KS * 32 + col_b0];
                cy[1] = fk0[mt * KS * 32 + col_b1];
              } else {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) dmma_acc(cy, fk0[(mt * KS + ks) * 32 + lane], bb[ks]);
              }
              const int i = Map::i_of(mt, fr);
              if (i < S) {
                double a0, a1;
                if (ltip) {
                  a0 = (double)((ma0 >> i) & 1);
                  a1 = (double)((ma1 >> i) & 1);
                } else if (lcur) {
                  a0 = c0[i];
                  a1 = c1[i];
                } else {
                  a0 = pa0_ok ? lg[((size_t)pa0 * K + k) * S + i] : 0.0;
                  a1 = pa1_ok ? lg[((size_t)(pa0 + 1) * K + k) * S + i] : 0.0;
                }
                acc0 += (spi[i] * a0) * cy[0];
                acc1 += (spi[i] * a1) * cy[1];
              }
            }
#pragma unroll
            for (int off = 4; off <= 16; off <<= 1) {
              acc0 += __shfl_xor_sync(FULL, acc0, off);
              acc1 += __shfl_xor_sync(FULL, acc1, off);
            }
            l0 += a.probs[k] * acc0;
            l1 += a.probs[k] * acc1;
          }
          if (fr == 0) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int64_t p = pa0 + u;
              if (p < a.N) {
                const double l = u ? l1 : l0;
                int c = 0;
                if (lcur || rcur) c += u ? sca1 : sca0;
                if (lk == TM_GLB) c += lgs[p];
                if (rk == TM_GLB) c += rgs[p];
                double lnl;
                if (a.pinvar >= 0.0) {
                  const MaskT m = inv[p];
                  double pv = 0.0;
                  for (int i = 0; i < S; ++i)
                    if ((m >> i) & 1) pv += spi[i];
                  lnl = lnl_pinvar(l, c, a.pinvar, pv);
                } else {
                  lnl = log(l) - (double)c * kLnScale;
                }
                if (a.site_lnl) a.site_lnl[p] = lnl;
                a.wsite[p] = (a.weights ? a.weights[p] : 1.0) * lnl;
              }
            }
          }
        }
      }
      ins = nins;
      ml = nml & keep;
      mr = nmr & keep;
      cur_ptr = next_ptr;
      const long long t_body = a.timing ? clock64() : 0;
      // done with this step's tables: the last warp to say so refills the buffer for the step after next
      // (release / acquire at CTA scope around the counter: every warp's table reads precede the refill.
      // compute-sanitizer's racecheck models barriers only and reports this hand-over as a potential WAR.)
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&done[buf], 1u) == (unsigned)NW - 1u) {
          __threadfence_block();
          done[buf] = 0;
          if (it + 2 < total_it) issue(it + 2);
        }
      }
      if (a.timing && tid == 0) {
        const int kk = lk | (rk << 2);
        const int v = s == a.n_steps ? 5 : (kk == 0 ? 0 : kk == (TM_CUR << 2) ? 1 : kk == (TM_GLB << 2) ? 2 : kk == (TM_CUR | (TM_GLB << 2)) ? 3 : 4);
        atomicAdd(a.timing + v * 4, (unsigned long long)(t_body - t_begin));
        atomicAdd(a.timing + v * 4 + 1, (unsigned long long)(clock64() - t_body));
        atomicAdd(a.timing + v * 4 + 2, 1ull);
        atomicAdd(a.timing + v * 4 + 3, (unsigned long long)(t_w1 - t_w0));
      }
    }
  }
  if constexpr (TMA) {
    if (lane == 0) bulk_wait0();
  }
}

}  // namespace phylo
