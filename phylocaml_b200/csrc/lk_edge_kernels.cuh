// Branch-length loop on one edge (a, b): Likelihood.adjust / distance_1 (lib/nodeData.ml:25-26,
// :32 -- the reference leaves them TODO, lib/likelihood_c.ml:19-24; MlModel exposes the
// eigensystem they need, lib/mlModel.ml:53-63, lib/mlmodel.c:325-342).
//
// With P_k(t) = UL diag(exp(lam_m r_k t)) UR (UL = U, UR = Ui for GTR; UL = U^T, UR = U for the
// symmetric path, lib/mlmodel.c:280-302) the site likelihood on the edge is
//     l(t) = sum_k p_k sum_m c_km exp(lam_m r_k t),
//     c_km = (sum_i pi_i a_ki UL[i][m]) * (sum_j UR[m][j] b_kj),
// so ONE pass over the two CLVs builds the "sum table" c (one CLV-sized array; it is an ordinary
// pruning update with M1[m][i] = pi_i UL[i][m] and M2[m][j] = UR[m][j] as the two "transition
// matrices", so phylo_lk_edge_prepare runs the pruning kernels of lk_kernels.cuh) and every
// further evaluation of lnL(t), dlnL/dt and d2lnL/dt2 -- for several t at once -- streams
// that single array: K*S*8 bytes per pattern instead of two CLVs and an S x S product.
//
// Sums over patterns use the canonical 1024-block fold so shards combine
// reproducibly. Values agree with root*_kernel to rounding (different association), not bits.
#pragma once
#include "common.cuh"

namespace phylo {

struct EdgeLengths {  // the branch lengths of one pass travel in the kernel parameters
  double t[16];
};

// PER > 0: lane `sub` of a group owns the PER consecutive table entries sub*PER .. sub*PER+PER-1
// and keeps their 3*PER coefficients in registers for the whole block (DNA+G4: G = 4, PER = 4:
// one 256-bit load and 12 FMAs per pattern and lane). PER = 0: entries strided by G, coefficients
// read from shared memory (any shape).
template <typename MaskT, int G, int PER>
__global__ void __launch_bounds__(256)
edge_eval_kernel(const double *__restrict__ sum, const int32_t *__restrict__ sum_sc,
                 const double *__restrict__ lam, const double *__restrict__ rates,
                 const double *__restrict__ probs, const double *__restrict__ pi, double pinvar,
                 const MaskT *__restrict__ inv, const double *__restrict__ weights,
                 const EdgeLengths tl, int n_t, int sym, int S, int K, int64_t N,
                 double *__restrict__ part) {
  extern __shared__ __align__(16) double esm[];
  __shared__ double vals[3][kLnlBlock];
  __shared__ double wsum[32];
  double *coef = esm;  // [n_t][3][K*S]
  const int KS = K * S;
  for (int i = threadIdx.x; i < n_t * KS; i += blockDim.x) {
    const int ti = i / KS, km = i - ti * KS, k = km / S, m = km - k * S;
    double tau = tl.t[ti] * rates[k];
    if (sym) tau = (double)(float)tau;  // compose_sym's `const float t` (lib/mlmodel.c:280)
    const double g = lam[m] * rates[k];
    const double e0 = probs[k] * (tau >= 1e-10 ? exp(lam[m] * tau) : 1.0);  // t < 1e-10 -> identity (:339-341)
    coef[(ti * 3 + 0) * KS + km] = e0;
    coef[(ti * 3 + 1) * KS + km] = e0 * g;
    coef[(ti * 3 + 2) * KS + km] = e0 * g * g;
  }
  __syncthreads();
  constexpr int PPB = 256 / G;  // patterns per pass of the CTA
  const int sub_lane = threadIdx.x % G, grp = threadIdx.x / G;
  const int64_t nblocks = (N + kLnlBlock - 1) / kLnlBlock;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    for (int ti = 0; ti < n_t; ++ti) {
      const double *c0 = coef + (ti * 3) * KS, *c1 = c0 + KS, *c2 = c1 + KS;
      // phase 1: the three sums of every pattern of the block, G lanes per pattern
      double cc[3][PER > 0 ? PER : 1];
      if (PER > 0) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int km = sub_lane * PER + j;
          cc[0][j] = km < KS ? c0[km] : 0.0;
          cc[1][j] = km < KS ? c1[km] : 0.0;
          cc[2][j] = km < KS ? c2[km] : 0.0;
        }
      }
#pragma unroll 4
      for (int sub = 0; sub < kLnlBlock / PPB; ++sub) {
        const int64_t p = blk * kLnlBlock + sub * PPB + grp;
        double l0 = 0.0, l1 = 0.0, l2 = 0.0;
        if (p < N) {
          const double *c = sum + p * KS;
          if (PER == 4) {  // 32-byte aligned when KS % 4 == 0 (the host only picks PER = 4 then)
            const d4 x = ld256_stream(c + 4 * sub_lane);
            l0 = ((x.x * cc[0][0] + x.y * cc[0][1]) + x.z * cc[0][2]) + x.w * cc[0][3];
            l1 = ((x.x * cc[1][0] + x.y * cc[1][1]) + x.z * cc[1][2]) + x.w * cc[1][3];
            l2 = ((x.x * cc[2][0] + x.y * cc[2][1]) + x.z * cc[2][2]) + x.w * cc[2][3];
          } else if (PER > 0) {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
              const int km = sub_lane * PER + j;
              const double x = km < KS ? __ldg(c + km) : 0.0;
              l0 += x * cc[0][j];
              l1 += x * cc[1][j];
              l2 += x * cc[2][j];
            }
          } else {
            for (int km = sub_lane; km < KS; km += G) {
              const double x = __ldg(c + km);
              l0 += x * c0[km];
              l1 += x * c1[km];
              l2 += x * c2[km];
            }
          }
        }
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
          l0 += __shfl_xor_sync(0xffffffffu, l0, off);
          l1 += __shfl_xor_sync(0xffffffffu, l1, off);
          l2 += __shfl_xor_sync(0xffffffffu, l2, off);
        }
        if (sub_lane == 0) {
          vals[0][sub * PPB + grp] = l0;
          vals[1][sub * PPB + grp] = l1;
          vals[2][sub * PPB + grp] = l2;
        }
      }
      __syncthreads();
      // phase 2: one thread per pattern turns the sums into weighted ln / derivative terms
      for (int i = threadIdx.x; i < kLnlBlock; i += 256) {
        const int64_t p = blk * kLnlBlock + i;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (p < N) {
          const double l0 = vals[0][i], l1 = vals[1][i], l2 = vals[2][i];
          const int sc = sum_sc[p];
          const double w = weights ? weights[p] : 1.0;
          if (pinvar >= 0.0) {
            const uint64_t mk = (uint64_t)inv[p];
            double pv = 0.0;
            for (int s = 0; s < S; ++s)
              if ((mk >> s) & 1) pv += pi[s];
            double ds;
            const double lnl = lnl_pinvar(l0, sc, pinvar, pv, &ds);
            const double r1 = ds * l1, r2 = ds * l2;
            v0 = w * lnl;
            v1 = w * r1;
            v2 = w * (r2 - r1 * r1);
          } else {
            const double r1 = l1 / l0, r2 = l2 / l0;
            v0 = w * (log(l0) - (double)sc * (kScaleExp * 0.6931471805599453094));
            v1 = w * r1;
            v2 = w * (r2 - r1 * r1);
          }
        }
        vals[0][i] = v0;
        vals[1][i] = v1;
        vals[2][i] = v2;
      }
      __syncthreads();
      for (int d = 0; d < 3; ++d) {
        const double r = block_fold_1024(vals[d], wsum);
        if (threadIdx.x == 0) part[(size_t)(ti * 3 + d) * nblocks + blk] = r;
        __syncthreads();
      }
    }
  }
}

// one CTA per row: canonical fold of `n` block partials (levels of 1024) into out[row]
__global__ void __launch_bounds__(256)
fold_rows_kernel(const double *__restrict__ part, int64_t n, double *__restrict__ out) {
  __shared__ double vals[kLnlBlock];
  __shared__ double lvl[kLnlBlock];
  __shared__ double wsum[32];
  const double *row = part + (size_t)blockIdx.x * n;
  // n <= 1024*1024 partials (2^30 patterns): two levels
  const int64_t nb = (n + kLnlBlock - 1) / kLnlBlock;
  for (int64_t b = 0; b < nb; ++b) {
    for (int i = threadIdx.x; i < kLnlBlock; i += blockDim.x) {
      const int64_t j = b * kLnlBlock + i;
      vals[i] = j < n ? row[j] : 0.0;
    }
    __syncthreads();
    const double r = block_fold_1024(vals, wsum);
    if (threadIdx.x == 0) lvl[b] = r;
    __syncthreads();
  }
  if (nb == 1) {
    if (threadIdx.x == 0) out[blockIdx.x] = lvl[0];
    return;
  }
  for (int i = threadIdx.x; i < kLnlBlock; i += blockDim.x) vals[i] = i < nb ? lvl[i] : 0.0;
  __syncthreads();
  const double r = block_fold_1024(vals, wsum);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

}  // namespace phylo
