// Model-parameter derivatives of lnL from retained, 3-directional CLVs (second half;
// the hooks the reference leaves as `failwith "todo"`: gen_subst_opt_func / gen_rates_opt_func /
// gen_prior_opt_func, lib/mlModel.ml:822-829).
//
// The likelihood of a site is multilinear in the branches' transition matrices, so for a parameter theta
//   d l / d theta = sum over branches e of  sum_k p_k (pi o a_e)^T [d P_k,e / d theta] b_e   (+ the prior term),
// where (a_e, b_e) is the directional CLV pair of branch e (down[v], up[v] after phylo_lk_uppass): every
// other branch's matrix is already folded into that pair. One pass over a branch's pair evaluates l and the
// derivative terms of ALL parameters at once (the pair is read once: 2C bytes per pattern and branch);
// d lnL / d theta = sum_s w_s (d site_s / d l_s) / site_s * d l_s  is accumulated per 1024-pattern block and
// parameter in a fixed branch order (deterministic).
//
// mats: [(np + 1)][K][S][S] for this branch: P_k first, then dP_k / d theta_p (built on the host from the
// eigensystem: d exp(X)[D] = V (Phi o (V^-1 D V)) V^-1, Phi_mn = (e^xm - e^xn) / (xm - xn)).
// CTA = one 1024-pattern block, thread = pattern (4 per thread). ST > 0: compile-time S (rows in
// registers); ST == 0: run-time S <= 64 (rows in local memory -- this path is for checking, not speed).
#pragma once
#include "common.cuh"
#include "lk_kernels.cuh"  // EdgeJoin, dmma_8x8x4

namespace phylo {

template <int ST, typename MaskT>
__global__ void __launch_bounds__(256)
param_grad_kernel(const double *__restrict__ mats, int np, int q0, int nq, const double *__restrict__ pi,
                  const double *__restrict__ dpi, const double *__restrict__ probs, double pinvar,
                  const MaskT *__restrict__ inv, const void *__restrict__ asrc, const int32_t *__restrict__ asc, int atip,
                  const void *__restrict__ bsrc, const int32_t *__restrict__ bsc, int btip,
                  const double *__restrict__ weights, double *__restrict__ gpart, int64_t n_part, int64_t N, int S_rt,
                  int K) {
  // parameters q0 .. q0 + nq - 1 of this pass (their matrices + P must fit shared memory)
  constexpr int SMAX = ST > 0 ? ST : 64;
  const int S = ST > 0 ? ST : S_rt;
  extern __shared__ __align__(16) double sm[];
  double *sP = sm;                                   // [K][S][S]
  double *sD = sP + (size_t)K * S * S;               // [nq][K][S][S]
  double *spi = sD + (size_t)nq * K * S * S;         // [S]
  double *sdpi = spi + S;                            // [nq][S] (zeros when dpi == NULL)
  double *vals = sdpi + (size_t)nq * S;              // [1024]
  double *wsum = vals + kLnlBlock;                   // [32]
  const size_t kss = (size_t)K * S * S;
  for (size_t i = threadIdx.x; i < kss; i += blockDim.x) sP[i] = mats[i];
  for (size_t i = threadIdx.x; i < (size_t)nq * kss; i += blockDim.x) sD[i] = mats[(size_t)(1 + q0) * kss + i];
  for (int i = threadIdx.x; i < S; i += blockDim.x) spi[i] = pi[i];
  for (int i = threadIdx.x; i < nq * S; i += blockDim.x) sdpi[i] = dpi ? dpi[(size_t)q0 * S + i] : 0.0;
  __syncthreads();
  const MaskT keep = (S >= 64) ? ~(MaskT)0 : (MaskT)(((uint64_t)1 << S) - 1);
  const int64_t lo = (int64_t)blockIdx.x * kLnlBlock;
  constexpr int NQ = 8;  // parameters per pass (host splits larger sets)
  double g[4][NQ];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int q = 0; q < NQ; ++q) g[u][q] = 0.0;
  for (int u = 0; u < 4; ++u) {
    const int64_t p = lo + threadIdx.x + 256 * u;
    if (p >= N) continue;
    double l = 0.0, dl[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) dl[q] = 0.0;
    MaskT ma = 0, mb = 0;
    if (atip) ma = ((const MaskT *)asrc)[p] & keep;
    if (btip) mb = ((const MaskT *)bsrc)[p] & keep;
    for (int k = 0; k < K; ++k) {
      double av[SMAX], bv[SMAX];
      for (int i = 0; i < S; ++i) {
        av[i] = atip ? (double)((ma >> i) & 1) : ((const double *)asrc)[((size_t)p * K + k) * S + i];
        bv[i] = btip ? (double)((mb >> i) & 1) : ((const double *)bsrc)[((size_t)p * K + k) * S + i];
      }
      const double *Pk = sP + (size_t)k * S * S;
      double lk = 0.0, dlk[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) dlk[q] = 0.0;
      for (int i = 0; i < S; ++i) {
        double y = 0.0;
        for (int j = 0; j < S; ++j) y += Pk[i * S + j] * bv[j];
        lk += (spi[i] * av[i]) * y;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          if (q < nq) {
            const double *Dk = sD + ((size_t)q * K + k) * S * S;
            double yd = 0.0;
            for (int j = 0; j < S; ++j) yd += Dk[i * S + j] * bv[j];
            dlk[q] += (spi[i] * av[i]) * yd + (sdpi[q * S + i] * av[i]) * y;
          }
        }
      }
      l += probs[k] * lk;
#pragma unroll
      for (int q = 0; q < NQ; ++q) dl[q] += probs[k] * dlk[q];
    }
    // d lnL_s / d l_s (the scale factors of l and d l are the same 2^(-256 c): they cancel without pinvar)
    double dscale;
    if (pinvar >= 0.0) {
      const int c = (atip ? 0 : asc[p]) + (btip ? 0 : bsc[p]);
      const MaskT m = inv[p];
      double pv = 0.0;
      for (int i = 0; i < S; ++i)
        if ((m >> i) & 1) pv += spi[i];
      (void)lnl_pinvar(l, c, pinvar, pv, &dscale);
    } else {
      dscale = 1.0 / l;
    }
    const double w = weights ? weights[p] : 1.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) g[u][q] = w * dscale * dl[q];
  }
  // canonical per-block fold, one parameter at a time; accumulated into gpart[q0 + q][block]
  for (int q = 0; q < nq; ++q) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double v = 0.0;
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq)
        if (qq == q) v = g[u][qq];
      vals[threadIdx.x + 256 * u] = v;
    }
    __syncthreads();
    const double r = block_fold_1024(vals, wsum);
    if (threadIdx.x == 0) gpart[(size_t)(q0 + q) * n_part + blockIdx.x] += r;
    __syncthreads();
  }
}

// ------------------------------------------- 4 states: every branch in one launch, fp64 tensor cores ----
// With X_k[i][j] = a_k[i] b_k[j] (the outer product of a branch's directional pair) the site likelihood and
// every derivative term are Frobenius products with host-built matrices:
//   l = sum_k <p_k pi_i P_k[i][j], X_k>,   dl_q = sum_k <p_k (pi_i dP_k/dtheta_q [i][j] + dpi_q,i P_k[i][j]), X_k>
// i.e. one (8 x 16K) x (16K x patterns) contraction: row 0 = l, rows 1..7 = up to seven parameters. It maps
// onto mma.sync.m8n8k4.f64 with a warp owning 8 patterns: lane (p = lane / 4, c = lane % 4) reads a_k (32 bytes,
// shared by the pattern's four lanes) and b_k[c], forms the B fragment a_k[s] * b_k[c] of k-step (k, s) with one
// multiply, and 4K chained DMMAs leave (row lane / 4; patterns 2c, 2c + 1) in two registers -- 16 DMMA + 16 DMUL
// per 8 patterns instead of 5 k FMAs + 3.6 k shared-memory reads, and the A fragments live in registers for a
// whole 1024-pattern block. Work item = (branch, block), all branches of a tree in one launch (descriptor table
// as in root4_batch_kernel); each item writes its block sum (fixed order, see the kernel) to gp[branch][q][block], and
// grad_sum_branches_kernel adds the branches in schedule order, so the result does not depend on the grid.
// afrag: [branch][K][4][32] (host: grad_fill_afrag).
// (non-volatile forms: read-only data and pure arithmetic, so the compiler is free to hoist the loads of a group
// above the DMMA chain of the previous one -- with the volatile helpers of lk_kernels.cuh every load waited
// for the DMMAs issued before it: 2.2 TB/s)
__device__ __forceinline__ d4 grad_ld256(const double *p) {
  d4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void grad_dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int K>
struct GradPair {  // one lane's operands of an 8-pattern group: a_k (all four states) and b_k[c], every class
  d4 a[K];
  double b[K];
};

// (p is clamped to the last pattern by the caller: a group past N loads valid addresses and its results are
// dropped with f = 0, so no load is predicated)
template <int K>
__device__ __forceinline__ void grad_load_pair(GradPair<K> &o, const EdgeJoin &ej, int64_t p, int c) {
  if (ej.atip) {
    const int m = ((const uint8_t *)ej.asrc)[p];
    const d4 v{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
#pragma unroll
    for (int k = 0; k < K; ++k) o.a[k] = v;
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) o.a[k] = grad_ld256((const double *)ej.asrc + ((size_t)p * K + k) * 4);
  }
  if (ej.btip) {
    const int m = ((const uint8_t *)ej.bsrc)[p];
#pragma unroll
    for (int k = 0; k < K; ++k) o.b[k] = (double)((m >> c) & 1);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) o.b[k] = __ldg((const double *)ej.bsrc + ((size_t)p * K + k) * 4 + c);
  }
}

// Occupancy, not prefetch depth, hides the latencies here (ncu of the ping-pong version with the A fragments in
// registers, 128 registers = 4 warps per scheduler: tensor + fp64 pipe 47 % busy, every warp serialised on its own
// DMMA chain -> reciprocal -> shuffle sequence): the A fragments live in shared memory (one conflict-free 8-byte
// read per DMMA), a warp holds one group's operands, and 3-4 CTAs share an SM.
template <int K>
__global__ void __launch_bounds__(256, K <= 4 ? 3 : 2)  // (4 CTAs = 64 registers measured the same: 3.96 vs 3.98 ms)
param_grad4_mma_kernel(const EdgeJoin *__restrict__ edges, int n_edges, const double *__restrict__ afrag, int nq,
                       const double *__restrict__ pi, double pinvar, const uint8_t *__restrict__ inv,
                       const double *__restrict__ weights, double *__restrict__ gp, int64_t N) {
  __shared__ double sA[2][K * 4 * 32];  // [item parity][k][s][lane]
  __shared__ double wsum[2][8][8];      // [item parity][warp][row]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pg = lane >> 2, c = lane & 3;
  constexpr int kGroups = kLnlBlock / 8 / 8;  // 8-pattern groups per warp and block
  const int64_t nblocks = (N + kLnlBlock - 1) / kLnlBlock, items = nblocks * n_edges;
  int parity = 0;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x, parity ^= 1) {
    const int edge = (int)(item / nblocks);
    const int64_t blk = item - (int64_t)edge * nblocks;
    const EdgeJoin ej = edges[edge];
    for (int i = threadIdx.x; i < K * 128; i += 256) sA[parity][i] = __ldg(afrag + (size_t)edge * K * 128 + i);
    const double *A = sA[parity] + lane;
    const int64_t pbase = blk * kLnlBlock + pg;
    GradPair<K> d;
    grad_load_pair<K>(d, ej, min(pbase + warp * 8, N - 1), c);
    __syncthreads();  // A fragments of this item (the other half of sA / wsum belongs to the previous item's readers)
    // this lane's running sums of w (d lnL_s / d l_s) d l_s: row pg, patterns 2c and 2c + 1 of the warp's groups.
    // The block result is built in a fixed order (groups in sequence per lane, the two columns, the row's four
    // lanes by xor 1 and 2, the eight warps in sequence), so it depends on the block alone -- not on the grid
    // or on which device scores the block.
    double g0 = 0.0, g1 = 0.0;
#pragma unroll 1
    for (int it = 0; it < kGroups; ++it) {
      const int pl0 = (warp + 8 * it) * 8;
      // two accumulator chains (even / odd k-steps), added at the end: DMMA latency is 26 cycles
      double acc[2] = {0.0, 0.0}, acc2[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double b = d.b[k];
        grad_dmma(acc, A[(k * 4 + 0) * 32], d.a[k].x * b);
        grad_dmma(acc2, A[(k * 4 + 1) * 32], d.a[k].y * b);
        grad_dmma(acc, A[(k * 4 + 2) * 32], d.a[k].z * b);
        grad_dmma(acc2, A[(k * 4 + 3) * 32], d.a[k].w * b);
      }
      if (it + 1 < kGroups) grad_load_pair<K>(d, ej, min(pbase + pl0 + 64, N - 1), c);  // (in flight during the epilogue)
      acc[0] += acc2[0];
      acc[1] += acc2[1];
      // row 0 (lanes 0-3) holds l of patterns pl0 + 2c, + 1: those lanes turn it into w * d lnL_s / d l_s
      double f0 = 0.0, f1 = 0.0;
      if (pg == 0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int64_t pp = blk * kLnlBlock + pl0 + 2 * c + u;
          if (pp < N) {
            const double l = acc[u];
            double ds;
            if (pinvar >= 0.0) {
              const int cnt = (ej.atip ? 0 : ej.asc[pp]) + (ej.btip ? 0 : ej.bsc[pp]);
              const int m = inv[pp];
              const double pv = (m & 1 ? pi[0] : 0.0) + (m & 2 ? pi[1] : 0.0) + (m & 4 ? pi[2] : 0.0) + (m & 8 ? pi[3] : 0.0);
              (void)lnl_pinvar(l, cnt, pinvar, pv, &ds);
            } else {
              ds = 1.0 / l;
            }
            const double f = (weights ? weights[pp] : 1.0) * ds;
            if (u == 0) f0 = f; else f1 = f;
          }
        }
      }
      f0 = __shfl_sync(0xffffffffu, f0, c);
      f1 = __shfl_sync(0xffffffffu, f1, c);
      if (f0 != 0.0) g0 += f0 * acc[0];  // (patterns past N: f = 0, and acc is the clamped pattern's)
      if (f1 != 0.0) g1 += f1 * acc[1];
    }
    double g = g0 + g1;
    g += __shfl_xor_sync(0xffffffffu, g, 1);
    g += __shfl_xor_sync(0xffffffffu, g, 2);
    if (c == 0) wsum[parity][warp][pg] = g;
    __syncthreads();
    if (threadIdx.x < nq) {
      double r = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) r += wsum[parity][w][1 + threadIdx.x];
      gp[((size_t)edge * nq + threadIdx.x) * nblocks + blk] = r;
    }
  }
}

// out[i] += gp[0][i] + gp[1][i] + ... in branch order (i over nq * nblocks)
__global__ void __launch_bounds__(256)
grad_sum_branches_kernel(const double *__restrict__ gp, int n_edges, int64_t cols, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cols) return;
  double s = out[i];
#pragma unroll 8
  for (int e = 0; e < n_edges; ++e) s += __ldg(gp + (size_t)e * cols + i);  // (loads batched by the unroll, adds in order)
  out[i] = s;
}

}  // namespace phylo
