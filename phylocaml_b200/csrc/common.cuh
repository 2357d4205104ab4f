// Shared device helpers for the phylocaml B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phylo {

constexpr int kSMs = 148;           // B200: 148 SMs; grids are sized in multiples of this
constexpr int kScaleExp = 256;      // rescale factor 2^256 (spec C.2.4)
constexpr int kScaleHiThresh = (1023 - kScaleExp) << 20;  // high word of 2^-256
constexpr int kLnlBlock = 1024;     // patterns per level-1 reduction block

struct __align__(32) d4 {
  double x, y, z, w;
};

// 256-bit global accesses (LDG.E.256 / STG.E.256 on sm_100a): one request moves a thread's
// whole 4-state fp64 vector. CLVs are streamed: read once, written once per evaluation.
__device__ __forceinline__ d4 ld256_stream(const double *p) {
  d4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(double *p, const d4 &v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z),
               "d"(v.w)
               : "memory");
}

__device__ __forceinline__ int hi32(double v) { return __double2hiint(v); }

// x[j] += x[j+off] for off = 16..1: lane 0 ends with the canonical 32-group fold
// (oracle/phylo_oracle.c reduce_block).
__device__ __forceinline__ double warp_fold(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

// Canonical fold of 1024 values held in shared memory `vals` (zero padded) into one double,
// returned on thread 0. `wsum` is 32 doubles of scratch. Any block size that is a multiple
// of 32 works; the result does not depend on it.
__device__ __forceinline__ double block_fold_1024(const double *vals, double *wsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int g = warp; g < 32; g += nwarps) {
    double v = warp_fold(vals[g * 32 + lane]);
    if (lane == 0) wsum[g] = v;
  }
  __syncthreads();
  double r = 0.0;
  if (warp == 0) r = warp_fold(wsum[lane]);
  return r;
}

// Site log-likelihood with an invariant-sites class (spec C.2.5):
//   ln( (1 - v) * l * 2^(-256 c) + v * pv ),  l = scaled site likelihood, c = summed scale counter,
// evaluated in the SCALED domain: ldexp(l, -256 c) underflows to 0 once c reaches 4-5 (site
// likelihood below 2^-1074, ordinary for deep trees), which would turn a variable site (pv == 0)
// into -inf. ln( (1-v) l + v pv 2^(256 c) ) - c 256 ln 2 is the same number without that underflow;
// when v pv 2^(256 c) would overflow it exceeds (1-v) l <= 1 by more than 2^64 and the sum is v pv
// to rounding. *dscale (may be NULL) = d(site)/d(l) / site, the factor that turns dl/dt, d2l/dt2
// into the derivative terms of the branch-length loop. oracle/phylo_oracle.c lnl_pinvar is the
// same statement.
__device__ __forceinline__ double lnl_pinvar(double l, int c, double pinvar, double pv, double *dscale = nullptr) {
  const double kLnScale = kScaleExp * 0.6931471805599453094;
  const double inv_term = pinvar * pv, var = (1.0 - pinvar) * l;
  if (inv_term == 0.0) {
    if (dscale) *dscale = 1.0 / l;
    return log(var) - (double)c * kLnScale;
  }
  if (c >= 4) {
    int ex;
    (void)frexp(inv_term, &ex);
    if (ex + kScaleExp * c > 64) {
      if (dscale) *dscale = 0.0;
      return log(inv_term);
    }
  }
  const double site = var + ldexp(inv_term, kScaleExp * c);
  if (dscale) *dscale = (1.0 - pinvar) / site;
  return log(site) - (double)c * kLnScale;
}

// operand kinds of a compiled (tree-fused) schedule step
enum : int { OPK_TIP = 0, OPK_CUR = 1, OPK_POP = 2, OPK_STORED = 3 };

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
// cp.async (LDGSTS): 16-byte global->shared copy that bypasses registers and L1
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

}  // namespace phylo
