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

}  // namespace phylo
