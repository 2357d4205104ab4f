// Likelihood kernels: P(t) construction, Felsenstein pruning update with per-site rescaling,
// root-edge site log-likelihood with the canonical site-sum reduction.
//
// Data layout in HBM: CLV[node][pattern][k][i] fp64 -- pattern-major, the K*S
// entries of one pattern contiguous (DNA+G4: 128 B = one cache line); scale[node][pattern]
// int32; tips stay as state masks (1 B per pattern for S<=8).
//
// The arithmetic follows SURVEY.md Appendix C.2 / oracle/phylo_oracle.c; the reference
// itself has no pruning code (lib/likelihood_c.ml:1-33 is all TODO) -- only P(t)
// (lib/mlmodel.c:280-342), which pt_build_kernel reproduces including its special cases.
#pragma once
#include "common.cuh"

namespace phylo {

// ------------------------------------------------------------------------- pt_build ----
// P[(b*K+k)][i][j] = sum_m U[i][m] * (exp(lam[m]*tau) * Ui[m][j]),  tau = tlen[b]*rates[k]
// (gtr, lib/mlmodel.c:325-342) or with U^T in place of U and Ui (sym, :280-302, where tau is
// first rounded to float as the reference's `const float t` does). tau == -1 -> Q,
// tau < 1e-10 -> identity. One CTA per (branch, rate class); latency-bound, tiny.
__global__ void pt_build_kernel(const double *__restrict__ U, const double *__restrict__ lam,
                                const double *__restrict__ Ui, const double *__restrict__ rates,
                                const double *__restrict__ tlen, int S, int K,
                                double *__restrict__ P, int interleave) {
  extern __shared__ double sh_e[];
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  const bool sym = (Ui == nullptr);
  double tau = tlen[b] * rates[k];
  if (sym) tau = (double)(float)tau;
  const int mode = (tau == -1.0) ? 2 : (tau >= 1e-10 ? 1 : 0);
  for (int m = threadIdx.x; m < S; m += blockDim.x)
    sh_e[m] = (mode == 2) ? lam[m] : exp(lam[m] * tau);
  __syncthreads();
  // interleave (4-state tree-fused kernel): per branch [e/2][k][e%2] so the K lanes of a
  // pattern read adjacent 16-byte chunks of shared memory; otherwise [k][i][j]
  double *out = interleave ? P + (size_t)b * K * S * S : P + (size_t)blockIdx.x * S * S;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx % S;
    double acc;
    if (mode == 0) {
      acc = (i == j) ? 1.0 : 0.0;
    } else {
      acc = 0.0;
      if (sym) {
        for (int m = 0; m < S; ++m) acc += U[m * S + i] * (sh_e[m] * U[m * S + j]);
      } else {
        for (int m = 0; m < S; ++m) acc += U[i * S + m] * (sh_e[m] * Ui[m * S + j]);
      }
    }
    out[interleave ? ((idx >> 1) * K + k) * 2 + (idx & 1) : idx] = acc;
  }
}

// ------------------------------------------------------------- 4-state pruning update ----
// Thread q owns item q = pattern*K + k: the 4 states of one rate class of one pattern, i.e.
// 32 contiguous bytes, so every warp request is one fully used 1 KB run (256-bit accesses).
// k = q % K is loop-invariant (the grid stride is a multiple of K), so the two 4x4 P
// matrices of that rate class stay in registers for the whole kernel. A tip child costs one
// byte per pattern; its contribution x_i = sum_{j in mask} P[i][j] is formed from the
// register-resident matrix with predicated adds in ascending j -- the same values in the
// same order as the oracle's sum_j P[i][j]*L[j] with L in {0,1}, so it is bit-identical,
// and it needs no shared-memory table (an earlier table version was bank-conflict bound).
// Per-site rescale: the K lanes of a pattern agree by xor-shuffle on max(high word); values
// are >= 0 so comparing high words as integers is the same test as max < 2^-256.
__device__ __forceinline__ void tip_contrib(const double (&pm)[16], int m, double (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double a = (m & 1) ? pm[i * 4 + 0] : 0.0;
    a += (m & 2) ? pm[i * 4 + 1] : 0.0;
    a += (m & 4) ? pm[i * 4 + 2] : 0.0;
    a += (m & 8) ? pm[i * 4 + 3] : 0.0;
    x[i] = a;
  }
}

// (body shared by prune4_kernel and prune4_level_kernel: bx / gx = this CTA's index and the number of CTAs
// that share the node)
template <int K, bool LTIP, bool RTIP, int U>
__device__ __forceinline__ void
prune4_body(const double *__restrict__ Pl, const double *__restrict__ Pr,
            const void *__restrict__ lsrc, const int32_t *__restrict__ lsc,
            const void *__restrict__ rsrc, const int32_t *__restrict__ rsc,
            double *__restrict__ out, int32_t *__restrict__ osc, int64_t N, int bx, int gx) {
  static_assert(K == 1 || K == 2 || K == 4 || K == 8 || K == 16, "K must divide the warp");
  const int k = threadIdx.x % K;
  double pl[16], pr[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) pl[e] = Pl[k * 16 + e];
#pragma unroll
  for (int e = 0; e < 16; ++e) pr[e] = Pr[k * 16 + e];

  const int64_t total = N * K;
  const int64_t stride = (int64_t)gx * blockDim.x * U;
  const double *lclv = (const double *)lsrc, *rclv = (const double *)rsrc;
  const uint8_t *ltip = (const uint8_t *)lsrc, *rtip = (const uint8_t *)rsrc;

  for (int64_t base = (int64_t)bx * blockDim.x * U; base < total; base += stride) {
    d4 l[U], r[U];
    int ml[U], mr[U], sc[U];
    bool act[U];
    // ---- phase 1: issue every global load of this iteration back to back (tip bytes, CLV
    // vectors, scale counters) so their latencies overlap; nothing here depends on a load
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
      act[u] = q < total;
      const int64_t p = q / K;
      sc[u] = 0;
      ml[u] = mr[u] = 0;
      l[u] = d4{0, 0, 0, 0};
      r[u] = d4{0, 0, 0, 0};
      if (act[u]) {
        if (LTIP) {
          ml[u] = ltip[p];
        } else {
          l[u] = ld256_stream(lclv + q * 4);
          if (k == 0) sc[u] = lsc[p];
        }
        if (RTIP) {
          mr[u] = rtip[p];
        } else {
          r[u] = ld256_stream(rclv + q * 4);
          if (k == 0) sc[u] += rsc[p];
        }
      }
    }
    // ---- phase 2: contraction, rescale, store
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
      double x[4], y[4];
      if (LTIP) {
        tip_contrib(pl, ml[u], x);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          x[i] = ((pl[i * 4 + 0] * l[u].x + pl[i * 4 + 1] * l[u].y) + pl[i * 4 + 2] * l[u].z) +
                 pl[i * 4 + 3] * l[u].w;
      }
      if (RTIP) {
        tip_contrib(pr, mr[u], y);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          y[i] = ((pr[i * 4 + 0] * r[u].x + pr[i * 4 + 1] * r[u].y) + pr[i * 4 + 2] * r[u].z) +
                 pr[i * 4 + 3] * r[u].w;
      }
      d4 v{x[0] * y[0], x[1] * y[1], x[2] * y[2], x[3] * y[3]};
      int h = max(max(hi32(v.x), hi32(v.y)), max(hi32(v.z), hi32(v.w)));
#pragma unroll
      for (int off = K / 2; off >= 1; off >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, off));
      const bool rescale = h < kScaleHiThresh;
      if (rescale) {
        const double f = 0x1p+256;
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
      }
      if (act[u]) {
        st256(out + q * 4, v);
        if (k == 0) osc[q / K] = sc[u] + (rescale ? 1 : 0);
      }
    }
  }
}

template <int K, bool LTIP, bool RTIP, int U>
__global__ void __launch_bounds__(256)
prune4_kernel(const double *__restrict__ Pl, const double *__restrict__ Pr,
              const void *__restrict__ lsrc, const int32_t *__restrict__ lsc,
              const void *__restrict__ rsrc, const int32_t *__restrict__ rsc,
              double *__restrict__ out, int32_t *__restrict__ osc, int64_t N) {
  prune4_body<K, LTIP, RTIP, U>(Pl, Pr, lsrc, lsc, rsrc, rsc, out, osc, N, blockIdx.x, gridDim.x);
}

// Many independent pruning updates in one launch (the pre-order pass: every update of one tree level):
// blockIdx.y picks the update from a descriptor table, the CTAs along x share it. Same body, so the values
// are those of the per-node launches bit for bit. Tip + tip updates keep their own kernel (prune4_tt_kernel).
struct PruneItem {
  const double *Pl, *Pr;
  const void *lsrc, *rsrc;
  const int32_t *lsc, *rsc;
  double *out;
  int32_t *osc;
  int ltip, rtip;
};
template <int K>
__global__ void __launch_bounds__(256)
prune4_level_kernel(const PruneItem *__restrict__ items, int64_t N) {
  const PruneItem it = items[blockIdx.y];
  if (it.ltip) prune4_body<K, true, false, 4>(it.Pl, it.Pr, it.lsrc, it.lsc, it.rsrc, it.rsc, it.out, it.osc, N, blockIdx.x, gridDim.x);
  else if (it.rtip) prune4_body<K, false, true, 4>(it.Pl, it.Pr, it.lsrc, it.lsc, it.rsrc, it.rsc, it.out, it.osc, N, blockIdx.x, gridDim.x);
  else prune4_body<K, false, false, 2>(it.Pl, it.Pr, it.lsrc, it.lsc, it.rsrc, it.rsc, it.out, it.osc, N, blockIdx.x, gridDim.x);
}

// ------------------------------------------------- 4-state update, both children tips ----
// With two tip children the parent CLV of a pattern depends only on the two 4-bit masks:
// 256 possible (maskL, maskR) pairs. Each CTA builds the 256 finished output rows once
// (K*4 doubles each, already rescaled, plus the rescale flag) in shared memory, laid out
// [pair][k][i] so the K lanes of a pattern read one contiguous row (no intra-pattern bank
// conflict); the kernel body is then 2 byte loads, one 32-byte table read and one 256-bit
// store per item -- a pure write stream. Values are built with the same operations in the
// same order as the general path (tip_contrib, product, rescale) => bit-identical.
template <int K, int U>
__global__ void __launch_bounds__(256, 4)
prune4_tt_kernel(const double *__restrict__ Pl, const double *__restrict__ Pr,
                 const uint8_t *__restrict__ ltip, const uint8_t *__restrict__ rtip,
                 double *__restrict__ out, int32_t *__restrict__ osc, int64_t N) {
  extern __shared__ __align__(32) unsigned char tt_smem[];
  d4 *tab = reinterpret_cast<d4 *>(tt_smem);                 // [256][K]
  int *flag = reinterpret_cast<int *>(tab + 256 * K);        // [256]
  for (int pair = threadIdx.x; pair < 256; pair += blockDim.x) {
    const int ml = pair >> 4, mr = pair & 15;
    int h = (int)0x80000000;
#pragma unroll 1
    for (int kk = 0; kk < K; ++kk) {
      double pl[16], pr[16], x[4], y[4];
#pragma unroll
      for (int e = 0; e < 16; ++e) { pl[e] = Pl[kk * 16 + e]; pr[e] = Pr[kk * 16 + e]; }
      tip_contrib(pl, ml, x);
      tip_contrib(pr, mr, y);
      const d4 v{x[0] * y[0], x[1] * y[1], x[2] * y[2], x[3] * y[3]};
      tab[pair * K + kk] = v;
      h = max(h, max(max(hi32(v.x), hi32(v.y)), max(hi32(v.z), hi32(v.w))));
    }
    const bool rescale = h < kScaleHiThresh;
    if (rescale) {
#pragma unroll 1
      for (int kk = 0; kk < K; ++kk) {
        d4 v = tab[pair * K + kk];
        v.x *= 0x1p+256; v.y *= 0x1p+256; v.z *= 0x1p+256; v.w *= 0x1p+256;
        tab[pair * K + kk] = v;
      }
    }
    flag[pair] = rescale ? 1 : 0;
  }
  __syncthreads();
  const int k = threadIdx.x % K;
  const int64_t total = N * K;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U; base < total; base += stride) {
    int pair[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
      pair[u] = 0;
      if (q < total) pair[u] = ((ltip[q / K] & 15) << 4) | (rtip[q / K] & 15);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
      if (q < total) {
        st256(out + q * 4, tab[pair[u] * K + k]);
        if (k == 0) osc[q / K] = flag[pair[u]];
      }
    }
  }
}

// ------------------------------------------------------------- 4-state root-edge lnL ----
// site likelihood l_s = sum_k probs[k] sum_i pi_i La[k,i] (sum_j P_k[i,j] Lb[k,j]); then
// ln l_s - c_s*256 ln2 (or the +pinvar form), times the pattern weight, folded per block of
// 1024 patterns in the canonical shape. One CTA per 1024-pattern block (grid-strided).
template <int K, bool ATIP, bool BTIP>
__global__ void __launch_bounds__(256, 4)
root4_kernel(const double *__restrict__ Proot, const double *__restrict__ pi,
             const double *__restrict__ probs, double pinvar, const uint8_t *__restrict__ inv,
             const void *__restrict__ asrc, const int32_t *__restrict__ asc,
             const void *__restrict__ bsrc, const int32_t *__restrict__ bsc,
             const double *__restrict__ weights, double *__restrict__ site_lnl,
             double *__restrict__ partials, int64_t N) {
  __shared__ double vals[kLnlBlock];
  __shared__ double wsum[32];
  const int k = threadIdx.x % K;
  double pm[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) pm[e] = Proot[k * 16 + e];
  const double pi0 = pi[0], pi1 = pi[1], pi2 = pi[2], pi3 = pi[3], pk = probs[k];
  const int64_t nblocks = (N + kLnlBlock - 1) / kLnlBlock;
  const double *aclv = (const double *)asrc, *bclv = (const double *)bsrc;
  const uint8_t *atip = (const uint8_t *)asrc, *btip = (const uint8_t *)bsrc;

  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    for (int it = 0; it < kLnlBlock * K; it += blockDim.x) {
      const int ql = it + threadIdx.x;
      const int plocal = ql / K;
      const int64_t p = blk * kLnlBlock + plocal;
      const bool act = p < N;
      d4 a{0, 0, 0, 0}, b{0, 0, 0, 0};
      int c = 0;
      if (act) {
        const int64_t q = p * K + k;
        if (ATIP) {
          const int m = atip[p];
          a = d4{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
        } else {
          a = ld256_stream(aclv + q * 4);
          if (k == 0) c += asc[p];
        }
        if (BTIP) {
          const int m = btip[p];
          b = d4{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
        } else {
          b = ld256_stream(bclv + q * 4);
          if (k == 0) c += bsc[p];
        }
      }
      double y0 = ((pm[0] * b.x + pm[1] * b.y) + pm[2] * b.z) + pm[3] * b.w;
      double y1 = ((pm[4] * b.x + pm[5] * b.y) + pm[6] * b.z) + pm[7] * b.w;
      double y2 = ((pm[8] * b.x + pm[9] * b.y) + pm[10] * b.z) + pm[11] * b.w;
      double y3 = ((pm[12] * b.x + pm[13] * b.y) + pm[14] * b.z) + pm[15] * b.w;
      double lk = (((pi0 * a.x) * y0 + (pi1 * a.y) * y1) + (pi2 * a.z) * y2) + (pi3 * a.w) * y3;
      double l = pk * lk;
#pragma unroll
      for (int off = 1; off < K; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
      if (k == 0) {
        double wl = 0.0;
        if (act) {
          double lnl;
          if (pinvar >= 0.0) {
            const int m = inv[p];
            const double pv = (m & 1 ? pi0 : 0.0) + (m & 2 ? pi1 : 0.0) + (m & 4 ? pi2 : 0.0) +
                              (m & 8 ? pi3 : 0.0);
            lnl = lnl_pinvar(l, c, pinvar, pv);
          } else {
            lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
          }
          if (site_lnl) site_lnl[p] = lnl;
          wl = (weights ? weights[p] : 1.0) * lnl;
        }
        vals[plocal] = wl;
      }
    }
    __syncthreads();
    const double r = block_fold_1024(vals, wsum);
    if (threadIdx.x == 0) partials[blk] = r;
    __syncthreads();
  }
}

// One more level of the canonical reduction: out[b] = fold(in[b*1024 .. b*1024+1023]).
__global__ void __launch_bounds__(256)
reduce1024_kernel(const double *__restrict__ in, int64_t n, double *__restrict__ out) {
  __shared__ double vals[kLnlBlock];
  __shared__ double wsum[32];
  const int64_t lo = (int64_t)blockIdx.x * kLnlBlock;
  for (int i = threadIdx.x; i < kLnlBlock; i += blockDim.x) vals[i] = (lo + i < n) ? in[lo + i] : 0.0;
  __syncthreads();
  const double r = block_fold_1024(vals, wsum);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

// ------------------------------------------------ many root-edge joins in one launch ----
// phylo_lk_edge_lnl_batch for 4 states: one work item = (edge, block of 1024 patterns). The edge's
// two operands (CLV + scale counters, or tip masks) and its K transition matrices come from a
// descriptor table; the arithmetic and the fold are root4_kernel's, so an edge's lnL is
// bit-identical to a single phylo_lk_edge_lnl call. partials[edge * nblocks + blk].
struct EdgeJoin {
  const void *asrc, *bsrc;
  const int32_t *asc, *bsc;
  int atip, btip;
};
template <int K>
__global__ void __launch_bounds__(256, 4)
root4_batch_kernel(const EdgeJoin *__restrict__ edges, int n_edges, const double *__restrict__ Pall,
                   const double *__restrict__ pi, const double *__restrict__ probs, double pinvar,
                   const uint8_t *__restrict__ inv, const double *__restrict__ weights,
                   double *__restrict__ partials, int64_t N) {
  __shared__ double vals[kLnlBlock];
  __shared__ double wsum[32];
  const int k = threadIdx.x % K;
  const double pi0 = pi[0], pi1 = pi[1], pi2 = pi[2], pi3 = pi[3], pk = probs[k];
  const int64_t nblocks = (N + kLnlBlock - 1) / kLnlBlock, items = nblocks * n_edges;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int edge = (int)(item / nblocks);
    const int64_t blk = item - (int64_t)edge * nblocks;
    const EdgeJoin ej = edges[edge];
    double pm[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) pm[e] = __ldg(Pall + ((size_t)edge * K + k) * 16 + e);
    const double *aclv = (const double *)ej.asrc, *bclv = (const double *)ej.bsrc;
    const uint8_t *atip = (const uint8_t *)ej.asrc, *btip = (const uint8_t *)ej.bsrc;
    for (int it = 0; it < kLnlBlock * K; it += blockDim.x) {
      const int ql = it + threadIdx.x;
      const int plocal = ql / K;
      const int64_t p = blk * kLnlBlock + plocal;
      const bool act = p < N;
      d4 a{0, 0, 0, 0}, b{0, 0, 0, 0};
      int c = 0;
      if (act) {
        const int64_t q = p * K + k;
        if (ej.atip) {
          const int m = atip[p];
          a = d4{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
        } else {
          a = ld256_stream(aclv + q * 4);
          if (k == 0) c += ej.asc[p];
        }
        if (ej.btip) {
          const int m = btip[p];
          b = d4{(double)(m & 1), (double)((m >> 1) & 1), (double)((m >> 2) & 1), (double)((m >> 3) & 1)};
        } else {
          b = ld256_stream(bclv + q * 4);
          if (k == 0) c += ej.bsc[p];
        }
      }
      double y0 = ((pm[0] * b.x + pm[1] * b.y) + pm[2] * b.z) + pm[3] * b.w;
      double y1 = ((pm[4] * b.x + pm[5] * b.y) + pm[6] * b.z) + pm[7] * b.w;
      double y2 = ((pm[8] * b.x + pm[9] * b.y) + pm[10] * b.z) + pm[11] * b.w;
      double y3 = ((pm[12] * b.x + pm[13] * b.y) + pm[14] * b.z) + pm[15] * b.w;
      double lk = (((pi0 * a.x) * y0 + (pi1 * a.y) * y1) + (pi2 * a.z) * y2) + (pi3 * a.w) * y3;
      double l = pk * lk;
#pragma unroll
      for (int off = 1; off < K; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
      if (k == 0) {
        double wl = 0.0;
        if (act) {
          double lnl;
          if (pinvar >= 0.0) {
            const int m = inv[p];
            const double pv = (m & 1 ? pi0 : 0.0) + (m & 2 ? pi1 : 0.0) + (m & 4 ? pi2 : 0.0) +
                              (m & 8 ? pi3 : 0.0);
            lnl = lnl_pinvar(l, c, pinvar, pv);
          } else {
            lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
          }
          wl = (weights ? weights[p] : 1.0) * lnl;
        }
        vals[plocal] = wl;
      }
    }
    __syncthreads();
    const double r = block_fold_1024(vals, wsum);
    if (threadIdx.x == 0) partials[item] = r;
    __syncthreads();
  }
}

// One more level of the canonical reduction for many rows at once: row = blockIdx.y,
// out[row * nb + b] = fold(in[row * n + b*1024 .. ]) with nb = gridDim.x.
__global__ void __launch_bounds__(256)
reduce1024_rows_kernel(const double *__restrict__ in, int64_t n, double *__restrict__ out) {
  __shared__ double vals[kLnlBlock];
  __shared__ double wsum[32];
  const double *row = in + (size_t)blockIdx.y * n;
  const int64_t lo = (int64_t)blockIdx.x * kLnlBlock;
  for (int i = threadIdx.x; i < kLnlBlock; i += blockDim.x) vals[i] = (lo + i < n) ? row[lo + i] : 0.0;
  __syncthreads();
  const double r = block_fold_1024(vals, wsum);
  if (threadIdx.x == 0) out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = r;
}

// -------------------------------------------------- any-S pruning update (20, 61, ...) ----
// One CTA = a tile of TP patterns (one per thread), looping over rate classes. Per class the
// two transition matrices (stored transposed, rows padded to a multiple of 4, so a thread
// fetches P[i..i+3][j] with one broadcast 32-byte read) and the two child tiles are staged in
// shared memory with coalesced loads; each thread then pulls its own child row into
// registers and runs 4 independent fp64 FMA chains per matrix row block. Rows in shared
// memory are padded to an odd number of doubles (conflict-free 64-bit column access).
// ST > 0: compile-time S (child row in registers); ST == 0: run-time S (row read from smem).
template <typename MaskT>
__device__ __forceinline__ void stage_tile(double *tile, int SP, const void *src, bool tip, int S,
                                           int K, int k, int64_t p0, int np) {
  if (tip) {
    const MaskT *m = (const MaskT *)src;
    for (int idx = threadIdx.x; idx < np * S; idx += blockDim.x) {
      const int r = idx / S, j = idx - r * S;
      tile[r * SP + j] = (double)((m[p0 + r] >> j) & 1);
    }
  } else {
    const double *clv = (const double *)src;
    for (int idx = threadIdx.x; idx < np * S; idx += blockDim.x) {
      const int r = idx / S, j = idx - r * S;
      tile[r * SP + j] = clv[((p0 + r) * K + k) * S + j];
    }
  }
}

__device__ __forceinline__ void stage_pt(double *pt, const double *P, int S, int S4) {
  // pt[j*S4 + i] = P[i][j], zero padded in i
  for (int idx = threadIdx.x; idx < S * S4; idx += blockDim.x) {
    const int j = idx / S4, i = idx - j * S4;
    pt[idx] = (i < S) ? P[i * S + j] : 0.0;
  }
}

// y[i0..i0+3] = sum_j P[i][j] * row[j]. With a compile-time S the j loop is fully unrolled so
// the child row really lives in registers (a partially unrolled loop would index it
// dynamically and push it to local memory).
template <int ST>
__device__ __forceinline__ void matvec4(const double *pt, int S, int S4, int i0, const double *rowreg,
                                        const double *rowsm, double acc[4]) {
  acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
  if constexpr (ST > 0) {
    constexpr int S4c = (ST + 3) & ~3;
    const double *pp = pt + i0;
#pragma unroll
    for (int j = 0; j < ST; ++j) {
      const double a = rowreg[j];
      const double2 p01 = *(const double2 *)(pp + j * S4c);
      const double2 p23 = *(const double2 *)(pp + j * S4c + 2);
      acc[0] += p01.x * a;
      acc[1] += p01.y * a;
      acc[2] += p23.x * a;
      acc[3] += p23.y * a;
    }
  } else {
#pragma unroll 4
    for (int j = 0; j < S; ++j) {
      const double a = rowsm[j];
      const double2 p01 = *(const double2 *)(pt + j * S4 + i0);
      const double2 p23 = *(const double2 *)(pt + j * S4 + i0 + 2);
      acc[0] += p01.x * a;
      acc[1] += p01.y * a;
      acc[2] += p23.x * a;
      acc[3] += p23.y * a;
    }
  }
}

template <int ST, typename MaskT, int TP>
__global__ void __launch_bounds__(TP)
prune_any_kernel(const double *__restrict__ Pl, const double *__restrict__ Pr,
                 const void *__restrict__ lsrc, const int32_t *__restrict__ lsc, bool ltip,
                 const void *__restrict__ rsrc, const int32_t *__restrict__ rsc, bool rtip,
                 double *__restrict__ out, int32_t *__restrict__ osc, int64_t N, int S, int K) {
  extern __shared__ __align__(16) double smem[];
  const int SS = ST > 0 ? ST : S;
  const int S4 = (SS + 3) & ~3, SP = SS | 1;
  double *ptl = smem, *ptr_ = ptl + SS * S4, *tl = ptr_ + SS * S4, *tr = tl + TP * SP;
  const int64_t ntiles = (N + TP - 1) / TP;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * TP;
    const int np = (int)min((int64_t)TP, N - p0);
    const bool act = threadIdx.x < np;
    int hmax = (int)0x80000000;
    for (int k = 0; k < K; ++k) {
      __syncthreads();  // previous class / tile fully consumed
      stage_pt(ptl, Pl + (size_t)k * SS * SS, SS, S4);
      stage_pt(ptr_, Pr + (size_t)k * SS * SS, SS, S4);
      stage_tile<MaskT>(tl, SP, lsrc, ltip, SS, K, k, p0, np);
      stage_tile<MaskT>(tr, SP, rsrc, rtip, SS, K, k, p0, np);
      __syncthreads();
      double *myl = tl + threadIdx.x * SP, *myr = tr + threadIdx.x * SP;
      double row[ST > 0 ? ST : 1];
      if (act) {
        if (ST > 0) {
#pragma unroll
          for (int j = 0; j < ST; ++j) row[j] = myl[j];
        }
        // x_i overwrite this thread's own left row in shared memory (it is in registers
        // by now); the products x_i*y_i then overwrite its right row, which becomes the
        // output tile for the coalesced store below.
        if (ST > 0) {
          for (int i0 = 0; i0 < SS; i0 += 4) {
            double acc[4];
            matvec4<ST>(ptl, SS, S4, i0, row, myl, acc);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (i0 + u < SS) myl[i0 + u] = acc[u];
          }
#pragma unroll
          for (int j = 0; j < ST; ++j) row[j] = myr[j];
          for (int i0 = 0; i0 < SS; i0 += 4) {
            double acc[4];
            matvec4<ST>(ptr_, SS, S4, i0, row, myr, acc);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (i0 + u < SS) {
                const double v = myl[i0 + u] * acc[u];
                myr[i0 + u] = v;
                hmax = max(hmax, hi32(v));
              }
          }
        }
      }
      if (ST == 0) {
        // run-time S: results cannot overwrite the rows they are computed from, so they go
        // to global directly (each thread its own pattern row; L2 absorbs the stride)
        if (act) {
          double *o = out + ((p0 + threadIdx.x) * K + k) * SS;
          for (int i0 = 0; i0 < SS; i0 += 4) {
            double ax[4], ay[4];
            matvec4<0>(ptl, SS, S4, i0, nullptr, myl, ax);
            matvec4<0>(ptr_, SS, S4, i0, nullptr, myr, ay);
            for (int u = 0; u < 4; ++u)
              if (i0 + u < SS) {
                const double v = ax[u] * ay[u];
                o[i0 + u] = v;
                hmax = max(hmax, hi32(v));
              }
          }
        }
      } else {
        __syncthreads();
        // coalesced store of the output tile (held in the right-child tile)
        for (int idx = threadIdx.x; idx < np * SS; idx += blockDim.x) {
          const int r = idx / SS, j = idx - r * SS;
          out[((p0 + r) * K + k) * SS + j] = tr[r * SP + j];
        }
      }
    }
    __syncthreads();  // this CTA's global writes visible to its own threads
    if (act) {
      const int64_t p = p0 + threadIdx.x;
      int c = (ltip ? 0 : lsc[p]) + (rtip ? 0 : rsc[p]);
      if (hmax < kScaleHiThresh) {  // rare: rescale the K*S entries in place
        double *o = out + p * K * SS;
        for (int e = 0; e < K * SS; ++e) o[e] *= 0x1p+256;
        c += 1;
      }
      osc[p] = c;
    }
  }
}

template <int ST, typename MaskT, int TP>
__global__ void __launch_bounds__(TP)
root_any_kernel(const double *__restrict__ Proot, const double *__restrict__ pi,
                const double *__restrict__ probs, double pinvar, const MaskT *__restrict__ inv,
                const void *__restrict__ asrc, const int32_t *__restrict__ asc, bool atip,
                const void *__restrict__ bsrc, const int32_t *__restrict__ bsc, bool btip,
                const double *__restrict__ weights, double *__restrict__ site_lnl,
                double *__restrict__ partials, int64_t N, int S, int K) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double vals[kLnlBlock];
  __shared__ double wsum[32];
  const int SS = ST > 0 ? ST : S;
  const int S4 = (SS + 3) & ~3, SP = SS | 1;
  double *pt = smem, *ta = pt + SS * S4, *tb = ta + TP * SP, *spi = tb + TP * SP;
  for (int i = threadIdx.x; i < SS; i += blockDim.x) spi[i] = pi[i];
  const int64_t nblocks = (N + kLnlBlock - 1) / kLnlBlock;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    for (int sub = 0; sub < kLnlBlock / TP; ++sub) {
      const int64_t p0 = blk * kLnlBlock + (int64_t)sub * TP;
      const int np = (int)max((int64_t)0, min((int64_t)TP, N - p0));
      const bool act = threadIdx.x < np;
      double l = 0.0;
      for (int k = 0; k < K; ++k) {
        __syncthreads();
        stage_pt(pt, Proot + (size_t)k * SS * SS, SS, S4);
        stage_tile<MaskT>(ta, SP, asrc, atip, SS, K, k, p0, np);
        stage_tile<MaskT>(tb, SP, bsrc, btip, SS, K, k, p0, np);
        __syncthreads();
        if (act) {
          const double *mya = ta + threadIdx.x * SP, *myb = tb + threadIdx.x * SP;
          double row[ST > 0 ? ST : 1];
          if (ST > 0) {
#pragma unroll
            for (int j = 0; j < ST; ++j) row[j] = myb[j];
          }
          double lk = 0.0;
          for (int i0 = 0; i0 < SS; i0 += 4) {
            double acc[4];
            matvec4<ST>(pt, SS, S4, i0, row, myb, acc);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (i0 + u < SS) lk += (spi[i0 + u] * mya[i0 + u]) * acc[u];
          }
          l += probs[k] * lk;
        }
      }
      double wl = 0.0;
      if (act) {
        const int64_t p = p0 + threadIdx.x;
        const int c = (atip ? 0 : asc[p]) + (btip ? 0 : bsc[p]);
        double lnl;
        if (pinvar >= 0.0) {
          const MaskT m = inv[p];
          double pv = 0.0;
          for (int i = 0; i < SS; ++i)
            if ((m >> i) & 1) pv += spi[i];
          lnl = lnl_pinvar(l, c, pinvar, pv);
        } else {
          lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
        }
        if (site_lnl) site_lnl[p] = lnl;
        wl = (weights ? weights[p] : 1.0) * lnl;
      }
      vals[sub * TP + threadIdx.x] = wl;
    }
    __syncthreads();
    const double r = block_fold_1024(vals, wsum);
    if (threadIdx.x == 0) partials[blk] = r;
    __syncthreads();
  }
}

// ------------------------------------- 20/61-state pruning update on fp64 tensor cores ----
// For S = 20 (amino acids) and S = 61 (codons) the update is a genuine dense contraction:
// X[i][p] = sum_j P[i][j] L[p][j] for every pattern p. It maps onto mma.sync.m8n8k4.f64
// (DMMA; measured 37 TFLOP/s on B200 vs 34 TFLOP/s for FMA, and ~6x fewer issued
// instructions per flop): M = 8 rows of P, N = 8 patterns, K = 4 columns.
//   * A fragments (P, zero padded to MT*8 x KS*4) are laid out once per CTA in shared memory
//     as [k][mt][ks][lane]: one conflict-free 8-byte read per DMMA;
//   * B fragments are the CLV rows themselves: lane (p = lane/4, c = lane%4) reads
//     L[p][ks*4 + c] straight from HBM (4 lanes = one 32-byte sector), tips from mask bits;
//   * the accumulators of X and Y share one layout (row i = mt*8 + lane/4, patterns
//     2*(lane%4), +1), so product, maximum and store need no data exchange; per-site
//     rescaling combines the row groups with three xor-shuffles.
// A warp owns 8 patterns and loops over the K rate classes; results are stored as they are
// produced and (rarely) rescaled in place afterwards by the lanes that wrote them.
// Summation order inside a DMMA differs from the oracle's ascending-j FMA chain: results agree
// to ~1e-16 relative (tests: CLV <= 1e-12, lnL <= 1e-9).
__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// Which contraction index j sits in k-slot (ks, fc) of a fragment, and which output state i in
// row slot (mt, fr), is free as long as the A table and the B / C fragments agree. The identity
// (j = 4 ks + fc, i = 8 mt + fr) makes every lane touch 8 bytes per access: a warp-wide load is
// 8 rows x 32 bytes, a store 4 rows x 64 bytes, and ncu shows the LSU 87 % busy with those
// (profiles/r01_ncu_prune_mma_*). For 20 states (160-byte rows, 32-byte aligned) the map below
// lets a lane load j = 4 fc .. 4 fc + 3 with one 256-bit access (k-steps 0-3; k-step 4 takes
// j = 16 + fc) and store i = 2 fr, 2 fr + 1 with one 128-bit access (row tiles 0, 1; tile 2 takes
// i = 16 + fr). 61-state rows are only 8-byte aligned (488 bytes), so codons keep the identity.
template <int S>
struct MmaMap {
  static constexpr bool kVec = false;
  __device__ static __forceinline__ int j_of(int ks, int fc) { return ks * 4 + fc; }
  __device__ static __forceinline__ int i_of(int mt, int fr) { return mt * 8 + fr; }
  // offset of column j inside one row tile of the A table ([ks][lane]), without the fr * 4 term
  __device__ static __forceinline__ int col_slot(int j) { return (j >> 2) * 32 + (j & 3); }
};
template <>
struct MmaMap<20> {
  static constexpr bool kVec = true;
  __device__ static __forceinline__ int j_of(int ks, int fc) { return ks < 4 ? 4 * fc + ks : 16 + fc; }
  __device__ static __forceinline__ int i_of(int mt, int fr) { return mt < 2 ? 2 * fr + mt : 16 + fr; }
  __device__ static __forceinline__ int col_slot(int j) { return j < 16 ? (j & 3) * 32 + (j >> 2) : 4 * 32 + (j - 16); }
};

__device__ __forceinline__ void st128(double *p, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// LT / RT: the side is a tip (compile-time, so the inner+inner instance carries none of the
// tip code: with run-time flags its main loop lost 20 % to register pressure).
// Measured and not kept (tools/mma_ab.sh, profiles/README.md): several pattern groups per warp
// sharing each A-fragment read (R = 2: +0.5 %, R = 4: 55 % slower for 20 states; R = 2: 19 % slower
// for 61 states); prefetching the next group's rows into L2 (inner+inner 193 -> 250 us); loading
// the B fragments of rate class k + 1 before the DMMAs of class k (80 registers, 193 -> 203 us).
template <int S, typename MaskT, bool LT, bool RT>
__global__ void __launch_bounds__(256)
prune_mma_kernel(const double *__restrict__ Pl, const double *__restrict__ Pr,
                 const void *__restrict__ lsrc, const int32_t *__restrict__ lsc,
                 const void *__restrict__ rsrc, const int32_t *__restrict__ rsc,
                 double *__restrict__ out, int32_t *__restrict__ osc, int64_t N, int K) {
  using Map = MmaMap<S>;
  constexpr bool ltip = LT, rtip = RT;
  constexpr int MT = (S + 7) / 8, KS = (S + 3) / 4;
  extern __shared__ __align__(16) double frag[];  // [2][K][MT][KS][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const size_t side = (size_t)K * MT * KS * 32;
  for (size_t idx = threadIdx.x; idx < 2 * side; idx += blockDim.x) {
    const int l = (int)(idx & 31);
    size_t rest = idx >> 5;
    const int ks = (int)(rest % KS); rest /= KS;
    const int mt = (int)(rest % MT); rest /= MT;
    const int k = (int)(rest % K);
    const int which = (int)(rest / K);
    const int i = Map::i_of(mt, l >> 2), j = Map::j_of(ks, l & 3);
    const double *P = which ? Pr : Pl;
    frag[idx] = (i < S && j < S) ? P[((size_t)k * S + i) * S + j] : 0.0;
  }
  __syncthreads();
  const double *fl = frag, *frg = frag + side;
  const double *lclv = (const double *)lsrc, *rclv = (const double *)rsrc;
  const MaskT *lmask = (const MaskT *)lsrc, *rmask = (const MaskT *)rsrc;
  const MaskT keep = (S >= 64) ? ~(MaskT)0 : (MaskT)(((uint64_t)1 << S) - 1);

  const int64_t ngroups = (N + 7) / 8;
  for (int64_t g = (int64_t)blockIdx.x * nwarps + warp; g < ngroups; g += (int64_t)gridDim.x * nwarps) {
    const int64_t pb = g * 8 + fr;            // pattern whose CLV row this lane loads (B fragment)
    const int64_t pa0 = g * 8 + 2 * fc;       // patterns whose results this lane holds (C fragment)
    const bool pb_ok = pb < N, pa0_ok = pa0 < N, pa1_ok = pa0 + 1 < N;
    MaskT ml = 0, mr = 0;
    if (ltip && pb_ok) ml = lmask[pb];
    if (rtip && pb_ok) mr = rmask[pb];
    // One-hot tips (the usual case: an observed state): P L is column j of P, taken straight
    // from the A-fragment table -- no DMMA for that side. 0/1 products and additions of +0 are
    // exact, so the values are those of the DMMA path bit for bit. Decided per warp.
    const bool lhot = ltip && __all_sync(0xffffffffu, !pb_ok || ((ml & keep) & ((ml & keep) - 1)) == 0);
    const bool rhot = rtip && __all_sync(0xffffffffu, !pb_ok || ((mr & keep) & ((mr & keep) - 1)) == 0);
    int jl0 = 0, jl1 = 0, jr0 = 0, jr1 = 0;  // state of the tip for this lane's two result patterns
    if (lhot) {
      jl0 = pa0_ok ? __ffsll((long long)(lmask[pa0] & keep)) - 1 : 0;
      jl1 = pa1_ok ? __ffsll((long long)(lmask[pa0 + 1] & keep)) - 1 : 0;
    }
    if (rhot) {
      jr0 = pa0_ok ? __ffsll((long long)(rmask[pa0] & keep)) - 1 : 0;
      jr1 = pa1_ok ? __ffsll((long long)(rmask[pa0 + 1] & keep)) - 1 : 0;
    }
    // element (i, j) of a fragment table: [row tile of i][k-step of j][lane = row slot * 4 + k slot]
    const int col_l0 = Map::col_slot(jl0) + fr * 4, col_l1 = Map::col_slot(jl1) + fr * 4;
    const int col_r0 = Map::col_slot(jr0) + fr * 4, col_r1 = Map::col_slot(jr1) + fr * 4;
    int h0 = (int)0x80000000, h1 = (int)0x80000000;
    for (int k = 0; k < K; ++k) {
      double bl[KS], br[KS];
      if constexpr (Map::kVec) {
        // 20 states: j = 4 fc .. 4 fc + 3 in one 256-bit load (k-steps 0-3), j = 16 + fc for k-step 4
        if (ltip) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) bl[ks] = (pb_ok && ((ml >> Map::j_of(ks, fc)) & 1)) ? 1.0 : 0.0;
        } else if (pb_ok) {
          const double *row = lclv + ((size_t)pb * K + k) * S;
          const d4 v = ld256_stream(row + 4 * fc);
          bl[0] = v.x; bl[1] = v.y; bl[2] = v.z; bl[3] = v.w; bl[4] = row[16 + fc];
        } else {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) bl[ks] = 0.0;
        }
        if (rtip) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) br[ks] = (pb_ok && ((mr >> Map::j_of(ks, fc)) & 1)) ? 1.0 : 0.0;
        } else if (pb_ok) {
          const double *row = rclv + ((size_t)pb * K + k) * S;
          const d4 v = ld256_stream(row + 4 * fc);
          br[0] = v.x; br[1] = v.y; br[2] = v.z; br[3] = v.w; br[4] = row[16 + fc];
        } else {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) br[ks] = 0.0;
        }
      } else {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int j = Map::j_of(ks, fc);
          const bool ok = pb_ok && j < S;
          if (ltip) bl[ks] = (ok && ((ml >> j) & 1)) ? 1.0 : 0.0;
          else bl[ks] = ok ? lclv[((size_t)pb * K + k) * S + j] : 0.0;
          if (rtip) br[ks] = (ok && ((mr >> j) & 1)) ? 1.0 : 0.0;
          else br[ks] = ok ? rclv[((size_t)pb * K + k) * S + j] : 0.0;
        }
      }
      const double *flk0 = fl + (size_t)k * MT * KS * 32, *frk0 = frg + (size_t)k * MT * KS * 32;
      const double *flk = flk0 + lane, *frk = frk0 + lane;
      auto tile = [&](int mt, double &v0, double &v1) {  // both patterns' results for row slot (mt, fr)
        double cx[2] = {0.0, 0.0}, cy[2] = {0.0, 0.0};
        if (lhot) {
          cx[0] = flk0[mt * KS * 32 + col_l0];
          cx[1] = flk0[mt * KS * 32 + col_l1];
        } else {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) dmma_8x8x4(cx, flk[(mt * KS + ks) * 32], bl[ks]);
        }
        if (rhot) {
          cy[0] = frk0[mt * KS * 32 + col_r0];
          cy[1] = frk0[mt * KS * 32 + col_r1];
        } else {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) dmma_8x8x4(cy, frk[(mt * KS + ks) * 32], br[ks]);
        }
        v0 = cx[0] * cy[0];
        v1 = cx[1] * cy[1];
      };
      if constexpr (Map::kVec) {
        // row tiles 0 and 1 hold i = 2 fr and 2 fr + 1: one 128-bit store per pattern; tile 2: i = 16 + fr
        double a0, a1, b0, b1, c0, c1;
        tile(0, a0, a1);
        tile(1, b0, b1);
        tile(2, c0, c1);
        double *o0 = out + ((size_t)pa0 * K + k) * S, *o1 = o0 + (size_t)K * S;
        if (pa0_ok) {
          st128(o0 + 2 * fr, a0, b0);
          h0 = max(h0, max(hi32(a0), hi32(b0)));
          if (fr < 4) { o0[16 + fr] = c0; h0 = max(h0, hi32(c0)); }
        }
        if (pa1_ok) {
          st128(o1 + 2 * fr, a1, b1);
          h1 = max(h1, max(hi32(a1), hi32(b1)));
          if (fr < 4) { o1[16 + fr] = c1; h1 = max(h1, hi32(c1)); }
        }
      } else {
#pragma unroll 2
        for (int mt = 0; mt < MT; ++mt) {
          double v0, v1;
          tile(mt, v0, v1);
          const int i = Map::i_of(mt, fr);
          if (i < S) {
            if (pa0_ok) { out[((size_t)pa0 * K + k) * S + i] = v0; h0 = max(h0, hi32(v0)); }
            if (pa1_ok) { out[((size_t)(pa0 + 1) * K + k) * S + i] = v1; h1 = max(h1, hi32(v1)); }
          }
        }
      }
    }
    // site maximum over all rows: combine the 8 row groups (lanes with equal lane%4)
#pragma unroll
    for (int off = 4; off <= 16; off <<= 1) {
      h0 = max(h0, __shfl_xor_sync(0xffffffffu, h0, off));
      h1 = max(h1, __shfl_xor_sync(0xffffffffu, h1, off));
    }
    const bool r0 = pa0_ok && h0 < kScaleHiThresh, r1 = pa1_ok && h1 < kScaleHiThresh;
    if (r0 || r1) {  // rare: rescale in place what this lane stored
      for (int k = 0; k < K; ++k)
        for (int mt = 0; mt < MT; ++mt) {
          const int i = Map::i_of(mt, fr);
          if (i < S) {
            if (r0) out[((size_t)pa0 * K + k) * S + i] *= 0x1p+256;
            if (r1) out[((size_t)(pa0 + 1) * K + k) * S + i] *= 0x1p+256;
          }
        }
    }
    if (fr == 0) {
      if (pa0_ok) osc[pa0] = (ltip ? 0 : lsc[pa0]) + (rtip ? 0 : rsc[pa0]) + (r0 ? 1 : 0);
      if (pa1_ok) osc[pa0 + 1] = (ltip ? 0 : lsc[pa0 + 1]) + (rtip ? 0 : rsc[pa0 + 1]) + (r1 ? 1 : 0);
    }
  }
}

// ---------------------------------------------------- 20-state tip+tip update as a table copy ----
// Both children are tips. For observed states (one-hot masks, the usual case) the result is
// out[k][i] = Pl[k][i][jl] * Pr[k][i][jr]: one of S*S finished vectors. tt_table_kernel builds all
// of them (256 KB for 20 states x 4 classes, 1.8 MB for 61 states: L2-resident), already rescaled
// where the rule asks for it, and prune_tt_copy_kernel turns the update into a table-row copy per
// pattern -- a pure write stream instead of 12 scattered shared-memory lookups per DMMA row tile
// (tip+tip was at 50 % of HBM for 20 states; 61 % with this path. For 61 states it measured slower
// than the DMMA kernel and is not dispatched). A pattern with an ambiguous tip (more
// than one state bit) is computed in place by the warp that meets it. The products are the ones the
// DMMA kernel's one-hot path forms (same two operands), so for observed tips the CLVs are
// bit-identical to it; an ambiguous tip is summed in ascending state order here, in fragment order
// there (agreement to rounding).
template <int S>
__global__ void __launch_bounds__(128)
tt_table_kernel(const double *__restrict__ Pl, const double *__restrict__ Pr, int K, double *__restrict__ tab,
                int32_t *__restrict__ tabsc) {
  __shared__ int wmax[4];
  const int jl = blockIdx.x / S, jr = blockIdx.x % S, KS = K * S;
  double *row = tab + (size_t)blockIdx.x * KS;
  int h = (int)0x80000000;
  for (int e = threadIdx.x; e < KS; e += blockDim.x) {  // e = k * S + i
    const double v = Pl[(size_t)e * S + jl] * Pr[(size_t)e * S + jr];
    row[e] = v;
    h = max(h, hi32(v));
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, off));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = h;
  __syncthreads();
  h = max(max(wmax[0], wmax[1]), max(wmax[2], wmax[3]));
  const bool rescale = h < kScaleHiThresh;
  if (rescale)
    for (int e = threadIdx.x; e < KS; e += blockDim.x) row[e] *= 0x1p+256;  // each thread rescales what it wrote
  if (threadIdx.x == 0) tabsc[blockIdx.x] = rescale ? 1 : 0;
}

template <int S, typename MaskT>
__global__ void __launch_bounds__(256)
prune_tt_copy_kernel(const double *__restrict__ Pl, const double *__restrict__ Pr, const double *__restrict__ tab,
                     const int32_t *__restrict__ tabsc, const MaskT *__restrict__ lmask,
                     const MaskT *__restrict__ rmask, double *__restrict__ out, int32_t *__restrict__ osc,
                     int64_t N, int K) {
  const int lane = threadIdx.x & 31, KS = K * S;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const MaskT keep = (S >= 64) ? ~(MaskT)0 : (MaskT)(((uint64_t)1 << S) - 1);
  const bool vec = (KS % 4) == 0;  // rows are whole 32-byte units (20 states): 256-bit copies
  const int64_t ngroups = (N + 7) / 8;
  for (int64_t g = warp; g < ngroups; g += nwarps) {
    const int64_t pmine = g * 8 + (lane & 7);
    MaskT ml = keep, mr = keep;
    if (pmine < N) { ml = lmask[pmine] & keep; mr = rmask[pmine] & keep; }
#pragma unroll 2
    for (int r = 0; r < 8; ++r) {
      const int64_t p = g * 8 + r;
      if (p >= N) break;  // warp-uniform
      const MaskT a = __shfl_sync(0xffffffffu, ml, r), b = __shfl_sync(0xffffffffu, mr, r);
      double *o = out + (size_t)p * KS;
      if ((a & (a - 1)) == 0 && (b & (b - 1)) == 0) {  // both observed: copy the finished vector
        const int combo = (__ffsll((long long)a) - 1) * S + (__ffsll((long long)b) - 1);
        const double *row = tab + (size_t)combo * KS;
        if (vec) {
          for (int q = lane; q < KS / 4; q += 32) st256(o + 4 * q, ld256_stream(row + 4 * q));
        } else {
          for (int e = lane; e < KS; e += 32) o[e] = row[e];
        }
        if (lane == 0) osc[p] = tabsc[combo];
      } else {  // ambiguous tip(s): sum the allowed columns (ascending j), rescale by the same rule
        int h = (int)0x80000000;
        for (int e = lane; e < KS; e += 32) {
          const double *pl = Pl + (size_t)e * S, *pr = Pr + (size_t)e * S;
          double sl = 0.0, sr = 0.0;
          for (int j = 0; j < S; ++j) {
            if ((a >> j) & 1) sl += pl[j];
            if ((b >> j) & 1) sr += pr[j];
          }
          const double v = sl * sr;
          o[e] = v;
          h = max(h, hi32(v));
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, off));
        const bool rescale = h < kScaleHiThresh;
        if (rescale)
          for (int e = lane; e < KS; e += 32) o[e] *= 0x1p+256;
        if (lane == 0) osc[p] = rescale ? 1 : 0;
      }
    }
  }
}

// ------------------------------------------ 20/61-state root-edge join on fp64 tensor cores ----
// Site values of the root edge with the fragment machinery of prune_mma_kernel: Y = P b by
// DMMA (or a column lookup for a one-hot tip), l = sum_k p_k sum_i (pi_i a_ki) Y_ki with the
// row sum folded across the 8 row groups by xor-shuffles. A warp owns 8 patterns, so the
// grid is as wide as the alignment (root_any_kernel ties one CTA to each 1024-pattern block:
// 196 CTAs of 128 threads for 200 k codon patterns). Weighted site ln-likelihoods go to
// `wsite`; the canonical fold runs over that array afterwards (reduce1024_kernel).
template <int S, typename MaskT, bool AT, bool BT>
__global__ void __launch_bounds__(256)
root_mma_kernel(const double *__restrict__ Proot, const double *__restrict__ pi,
                const double *__restrict__ probs, double pinvar, const MaskT *__restrict__ inv,
                const void *__restrict__ asrc, const int32_t *__restrict__ asc,
                const void *__restrict__ bsrc, const int32_t *__restrict__ bsc,
                const double *__restrict__ weights, double *__restrict__ site_lnl,
                double *__restrict__ wsite, int64_t N, int K) {
  constexpr int MT = (S + 7) / 8, KS = (S + 3) / 4;
  extern __shared__ __align__(16) double frag[];  // [K][MT][KS][32] + pi[S]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const size_t side = (size_t)K * MT * KS * 32;
  double *spi = frag + side;
  for (size_t idx = threadIdx.x; idx < side; idx += blockDim.x) {
    const int l = (int)(idx & 31);
    size_t rest = idx >> 5;
    const int ks = (int)(rest % KS); rest /= KS;
    const int mt = (int)(rest % MT); rest /= MT;
    const int k = (int)rest;
    const int i = mt * 8 + (l >> 2), j = ks * 4 + (l & 3);
    frag[idx] = (i < S && j < S) ? Proot[((size_t)k * S + i) * S + j] : 0.0;
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) spi[i] = pi[i];
  __syncthreads();
  const double *aclv = (const double *)asrc, *bclv = (const double *)bsrc;
  const MaskT *amask = (const MaskT *)asrc, *bmask = (const MaskT *)bsrc;
  const MaskT keep = (S >= 64) ? ~(MaskT)0 : (MaskT)(((uint64_t)1 << S) - 1);

  const int64_t ngroups = (N + 7) / 8;
  for (int64_t g = (int64_t)blockIdx.x * nwarps + warp; g < ngroups; g += (int64_t)gridDim.x * nwarps) {
    const int64_t pb = g * 8 + fr, pa0 = g * 8 + 2 * fc;
    const bool pb_ok = pb < N, pa0_ok = pa0 < N, pa1_ok = pa0 + 1 < N;
    MaskT mb = 0, ma0 = 0, ma1 = 0;
    if (BT && pb_ok) mb = bmask[pb];
    if (AT) {
      if (pa0_ok) ma0 = amask[pa0];
      if (pa1_ok) ma1 = amask[pa0 + 1];
    }
    const bool bhot = BT && __all_sync(0xffffffffu, !pb_ok || ((mb & keep) & ((mb & keep) - 1)) == 0);
    int jb0 = 0, jb1 = 0;
    if (bhot) {
      jb0 = pa0_ok ? __ffsll((long long)(bmask[pa0] & keep)) - 1 : 0;
      jb1 = pa1_ok ? __ffsll((long long)(bmask[pa0 + 1] & keep)) - 1 : 0;
    }
    const int col_b0 = (jb0 >> 2) * 32 + fr * 4 + (jb0 & 3), col_b1 = (jb1 >> 2) * 32 + fr * 4 + (jb1 & 3);
    double l0 = 0.0, l1 = 0.0;
    for (int k = 0; k < K; ++k) {
      double bb[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int j = ks * 4 + fc;
        const bool ok = pb_ok && j < S;
        if (BT) bb[ks] = (ok && ((mb >> j) & 1)) ? 1.0 : 0.0;
        else bb[ks] = ok ? bclv[((size_t)pb * K + k) * S + j] : 0.0;
      }
      const double *fk0 = frag + (size_t)k * MT * KS * 32, *fk = fk0 + lane;
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 2
      for (int mt = 0; mt < MT; ++mt) {
        double cy[2] = {0.0, 0.0};
        if (bhot) {
          cy[0] = fk0[mt * KS * 32 + col_b0];
          cy[1] = fk0[mt * KS * 32 + col_b1];
        } else {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) dmma_8x8x4(cy, fk[(mt * KS + ks) * 32], bb[ks]);
        }
        const int i = mt * 8 + fr;
        if (i < S) {
          double a0, a1;
          if (AT) {
            a0 = (double)((ma0 >> i) & 1);
            a1 = (double)((ma1 >> i) & 1);
          } else {
            a0 = pa0_ok ? aclv[((size_t)pa0 * K + k) * S + i] : 0.0;
            a1 = pa1_ok ? aclv[((size_t)(pa0 + 1) * K + k) * S + i] : 0.0;
          }
          acc0 += (spi[i] * a0) * cy[0];
          acc1 += (spi[i] * a1) * cy[1];
        }
      }
#pragma unroll
      for (int off = 4; off <= 16; off <<= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
      }
      l0 += probs[k] * acc0;
      l1 += probs[k] * acc1;
    }
    if (fr == 0) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t p = pa0 + u;
        if (p < N) {
          const double l = u ? l1 : l0;
          const int c = (AT ? 0 : asc[p]) + (BT ? 0 : bsc[p]);
          double lnl;
          if (pinvar >= 0.0) {
            const MaskT m = inv[p];
            double pv = 0.0;
            for (int i = 0; i < S; ++i)
              if ((m >> i) & 1) pv += spi[i];
            lnl = lnl_pinvar(l, c, pinvar, pv);
          } else {
            lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
          }
          if (site_lnl) site_lnl[p] = lnl;
          wsite[p] = (weights ? weights[p] : 1.0) * lnl;
        }
      }
    }
  }
}

// ----------------------------------------------------------------- tip preparation ----
// Converts raw tip masks of any element width to the device width (rows padded to
// out_stride elements so bulk-TMA tile loads stay aligned and in bounds), masks off bits >= S,
// counts invalid (empty) masks and builds the per-pattern AND over all tips (used by the
// invariant-sites term).
template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
tips_prepare_kernel(const InT *__restrict__ in, OutT *__restrict__ out, int64_t out_stride,
                    OutT *__restrict__ inv, int T, int64_t N, int S,
                    unsigned long long *__restrict__ n_bad, int64_t p_lo, int64_t p_hi,
                    const uint64_t *__restrict__ lut) {
  // lut != NULL (1-byte input only): the cells are alphabet symbols, lut[byte] is the state mask
  // (phylo_engine_set_symbol_table); a symbol the table does not know maps to 0 and is counted bad
  // patterns [p_lo, p_hi): a slab of the alignment (the whole of it for a plain set_tips)
  const uint64_t keep = (S >= 64) ? ~0ull : ((1ull << S) - 1);
  unsigned long long bad = 0;
  for (int64_t p = p_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p_hi;
       p += (int64_t)gridDim.x * blockDim.x) {
    uint64_t all = keep;
    for (int t = 0; t < T; ++t) {
      const uint64_t raw = (uint64_t)in[(int64_t)t * N + p];
      const uint64_t m = (lut ? lut[raw & 0xff] : raw) & keep;
      out[(int64_t)t * out_stride + p] = (OutT)m;
      if (m == 0) ++bad;
      all &= m;
    }
    inv[p] = (OutT)all;
  }
  if (bad) atomicAdd(n_bad, bad);
}

// Two 4-bit masks per byte (even pattern in the low nibble) for the tree-fused kernel's tip
// tiles; cells beyond N read as "all states". Bytes [b_lo, b_hi) of every row. Output is
// group-major: [group of 32 patterns][T][16 bytes], so the tips a warp/tile needs are one
// contiguous run (a single bulk-TMA copy).
__global__ void __launch_bounds__(256)
tips_pack4_kernel(const uint8_t *__restrict__ tips, uint8_t *__restrict__ tips4, int T, int64_t N,
                  int64_t stride, int64_t b_lo, int64_t b_hi) {
  // b_lo, b_hi are multiples of 16 (whole groups); thread i writes output byte i of the slab
  const int64_t g_lo = b_lo >> 4, total = (b_hi - b_lo) * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gt = i >> 4, g = g_lo + gt / T, t = gt % T, p = 2 * ((g << 4) + (i & 15));
    const int lo = p < N ? (tips[t * stride + p] & 15) : 15;
    const int hi = p + 1 < N ? (tips[t * stride + p + 1] & 15) : 15;
    tips4[(g_lo << 4) * T + i] = (uint8_t)(lo | (hi << 4));
  }
}

// ---- packed 4-state input (phylo_lk_set_tips with mask_bytes == 0): the host alignment is already
// two 4-bit masks per byte, tip-major rows of `in_pitch` bytes. No staging copy at device width, no
// re-packing: the upload is half the bytes of the one-byte-per-cell form and this kernel only turns
// the tip-major rows into the group-major tiles the tree-fused kernel bulk-copies
// ([group of 32 patterns][T][16 bytes]). A CTA moves a tile of 32 taxa x 32 groups through shared
// memory: rows are read as 512-byte runs (one taxon per warp, 16 bytes per lane), tiles are written
// as T*16-byte runs per group. Patterns >= N (last byte / trailing groups up to the 1024-padded
// stride) are forced to "all states". Groups [g_lo, g_hi).
__global__ void __launch_bounds__(256)
tips_nibbles_to_groups_kernel(const uint8_t *__restrict__ in, uint64_t in_pitch, uint8_t *__restrict__ tips4, int T,
                              int64_t N, int64_t g_lo, int64_t g_hi) {
  __shared__ uint4 tile[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t g0 = g_lo + (int64_t)blockIdx.x * 32;
  const int t0 = blockIdx.y * 32;
  const int64_t n_bytes = (N + 1) >> 1;  // valid bytes per input row
  for (int r = warp; r < 32; r += 8) {   // taxon t0 + r: lane reads the 16 bytes of group g0 + lane
    const int t = t0 + r;
    const int64_t g = g0 + lane;
    uint4 v = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (t < T && g < g_hi) {
      const int64_t b = g << 4;
      const uint8_t *src = in + (uint64_t)t * in_pitch + b;
      if (b + 16 <= n_bytes && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        v = *reinterpret_cast<const uint4 *>(src);
      } else {
        uint8_t tmp[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) tmp[i] = (b + i < n_bytes) ? src[i] : 0xff;
        v = *reinterpret_cast<uint4 *>(tmp);
      }
      if (2 * (b + 16) > N) {  // the group holds patterns >= N: all states there
        uint8_t *q = reinterpret_cast<uint8_t *>(&v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int64_t pat = 2 * (b + i);
          if (pat >= N) q[i] = 0xff;
          else if (pat + 1 >= N) q[i] |= 0xf0;
        }
      }
    }
    tile[r][lane] = v;
  }
  __syncthreads();
  for (int gg = warp; gg < 32; gg += 8) {  // group g0 + gg: lane writes taxon t0 + lane's 16 bytes
    const int64_t g = g0 + gg;
    const int t = t0 + lane;
    if (g < g_hi && t < T) *reinterpret_cast<uint4 *>(tips4 + ((uint64_t)g * T + t) * 16) = tile[lane][gg];
  }
}

// Validation of the group-major nibble tiles (a nibble with no state bit is an invalid cell) and, when
// the model has an invariant-sites class, the AND over all tips per pattern (`inv`, one byte per
// pattern). Thread = one byte (two patterns) of a group; the 16 threads of a group walk its T rows.
__global__ void __launch_bounds__(256)
tips_groups_check_kernel(const uint8_t *__restrict__ tips4, int T, int64_t N, int64_t g_lo, int64_t g_hi,
                         uint8_t *__restrict__ inv, unsigned long long *__restrict__ n_bad) {
  unsigned long long bad = 0;
  const int64_t total = (g_hi - g_lo) * 16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = g_lo + (i >> 4);
    const int b = (int)(i & 15);
    const int64_t p = (g << 5) + 2 * b;
    if (p >= N) continue;
    const uint8_t *col = tips4 + (uint64_t)g * T * 16 + b;
    unsigned all = 0xff;
    for (int t = 0; t < T; ++t) {
      const unsigned v = col[(size_t)t * 16];
      bad += ((v & 15) == 0) + ((p + 1 < N) && ((v >> 4) == 0));
      all &= v;
    }
    if (inv) {
      inv[p] = (uint8_t)(all & 15);
      if (p + 1 < N) inv[p + 1] = (uint8_t)(all >> 4);
    }
  }
  if (bad) atomicAdd(n_bad, bad);
}

// One byte per cell rows (what the per-node kernels and the edge functions read for a tip operand)
// rebuilt from the group-major nibble tiles: done lazily, only when such a kernel first needs them
// after a packed upload. Thread = one output byte pair.
__global__ void __launch_bounds__(256)
tips_groups_to_bytes_kernel(const uint8_t *__restrict__ tips4, uint8_t *__restrict__ tips, int T, int64_t stride) {
  const int64_t pairs = stride >> 1, total = pairs * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / pairs, b = i - t * pairs, g = b >> 4;
    const unsigned v = tips4[((uint64_t)g * T + t) * 16 + (b & 15)];
    *reinterpret_cast<uchar2 *>(tips + t * stride + 2 * b) = make_uchar2((unsigned char)(v & 15), (unsigned char)(v >> 4));
  }
}

}  // namespace phylo
