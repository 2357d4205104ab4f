// Microbenchmark prepared for round 2 (NOT yet run on a GPU): does landing the CLV rows of the
// 20-state DMMA pruning update in shared memory (cp.async, double buffered per warp) instead of
// registers lift the inner+inner kernel off its register-bound data-in-flight limit?
// (profiles/README.md: tensor pipe 58 %, DRAM 60 %, L1 77 %, 32 warps/SM at 64 registers.)
//
// Two kernels compute the same update out[p][k][i] = (sum_j Pl[k][i][j] L[p][k][j]) * (sum_j Pr[k][i][j] R[p][k][j])
// for S = 20, K = 4 (no tips, no rescaling -- the parts that do not matter for the question):
//   A  "reg":  the product kernel's scheme -- each lane loads its B fragments straight from HBM
//              (one LDG.256 + one LDG.64 per side and class), 8 warps x 4 CTAs per SM;
//   B  "smem": a warp copies the next group's two contiguous 5 KB blocks (8 patterns x 640 B per
//              side) into its own padded staging buffer with cp.async while it computes the current
//              group out of the other buffer; row pitch 656 B makes the fragment reads conflict-free;
//              8 warps x 1 CTA per SM (2 x 10.5 KB per warp + 30 KB A-fragment table).
// Prints ms, GB/s (3 x 640 B per pattern) and useful TFLOP/s for both and checks that the outputs
// are bit-identical (same DMMA order). Build: make -C tools/ubench mma_stage; run under gpurun.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int S = 20, K = 4, MT = 3, KS = 5, ROW = K * S;  // ROW doubles per pattern
constexpr int PITCH = 82;                                   // staging row pitch in doubles (656 B)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ int j_of(int ks, int fc) { return ks < 4 ? 4 * fc + ks : 16 + fc; }
__device__ __forceinline__ int i_of(int mt, int fr) { return mt < 2 ? 2 * fr + mt : 16 + fr; }
__device__ __forceinline__ void cp16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void build_table(double *frag, const double *Pl, const double *Pr) {
  const int side = K * MT * KS * 32;
  for (int idx = threadIdx.x; idx < 2 * side; idx += blockDim.x) {
    const int l = idx & 31;
    int rest = idx >> 5;
    const int ks = rest % KS; rest /= KS;
    const int mt = rest % MT; rest /= MT;
    const int k = rest % K, which = rest / K;
    const int i = i_of(mt, l >> 2), j = j_of(ks, l & 3);
    const double *P = which ? Pr : Pl;
    frag[idx] = (i < S && j < S) ? P[(k * S + i) * S + j] : 0.0;
  }
  __syncthreads();
}

// the DMMAs and stores of one rate class, B fragments given
__device__ __forceinline__ void compute_class(const double *fl, const double *fr_, int lane, int k, const double (&bl)[KS],
                                              const double (&br)[KS], double *o0, double *o1, bool ok0, bool ok1) {
  const int fr = lane >> 2;
  double v0[MT], v1[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    double cx[2] = {0, 0}, cy[2] = {0, 0};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) dmma(cx, fl[((k * MT + mt) * KS + ks) * 32 + lane], bl[ks]);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) dmma(cy, fr_[((k * MT + mt) * KS + ks) * 32 + lane], br[ks]);
    v0[mt] = cx[0] * cy[0];
    v1[mt] = cx[1] * cy[1];
  }
  if (ok0) {
    *reinterpret_cast<double2 *>(o0 + k * S + 2 * fr) = make_double2(v0[0], v0[1]);
    if (fr < 4) o0[k * S + 16 + fr] = v0[2];
  }
  if (ok1) {
    *reinterpret_cast<double2 *>(o1 + k * S + 2 * fr) = make_double2(v1[0], v1[1]);
    if (fr < 4) o1[k * S + 16 + fr] = v1[2];
  }
}

__global__ void __launch_bounds__(256) kern_reg(const double *Pl, const double *Pr, const double *L, const double *R, double *out,
                                                long N) {
  extern __shared__ __align__(16) double frag[];
  build_table(frag, Pl, Pr);
  const double *fl = frag, *fr_ = frag + K * MT * KS * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, fr = lane >> 2, fc = lane & 3;
  const long ngroups = (N + 7) / 8;
  for (long g = (long)blockIdx.x * nw + warp; g < ngroups; g += (long)gridDim.x * nw) {
    const long pb = g * 8 + fr, pa0 = g * 8 + 2 * fc;
    for (int k = 0; k < K; ++k) {
      double bl[KS], br[KS];
      if (pb < N) {
        const double4 a = *reinterpret_cast<const double4 *>(L + pb * ROW + k * S + 4 * fc);
        const double4 b = *reinterpret_cast<const double4 *>(R + pb * ROW + k * S + 4 * fc);
        bl[0] = a.x; bl[1] = a.y; bl[2] = a.z; bl[3] = a.w; bl[4] = L[pb * ROW + k * S + 16 + fc];
        br[0] = b.x; br[1] = b.y; br[2] = b.z; br[3] = b.w; br[4] = R[pb * ROW + k * S + 16 + fc];
      } else {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) bl[ks] = br[ks] = 0.0;
      }
      compute_class(fl, fr_, lane, k, bl, br, out + pa0 * ROW, out + (pa0 + 1) * ROW, pa0 < N, pa0 + 1 < N);
    }
  }
}

__global__ void __launch_bounds__(256) kern_smem(const double *Pl, const double *Pr, const double *L, const double *R, double *out,
                                                 long N) {
  extern __shared__ __align__(16) double smem[];
  double *frag = smem;
  build_table(frag, Pl, Pr);
  const double *fl = frag, *fr_ = frag + K * MT * KS * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, fr = lane >> 2, fc = lane & 3;
  double *stage = smem + 2 * K * MT * KS * 32 + (size_t)warp * (2 * 2 * 8 * PITCH);  // [buf][side][8 rows][PITCH]
  const long ngroups = (N + 7) / 8, stride = (long)gridDim.x * nw;
  auto issue = [&](long g, int buf) {  // 8 rows x 640 B per side = 320 chunks of 16 B per side
    const long rows = (N - g * 8 < 8) ? (N - g * 8) : 8;
    double *dst = stage + buf * (2 * 8 * PITCH);
    for (int c = lane; c < 320; c += 32) {
      const int row = c / 40, col = c % 40;
      if (row < rows) {
        cp16(dst + row * PITCH + 2 * col, L + (g * 8 + row) * ROW + 2 * col);
        cp16(dst + 8 * PITCH + row * PITCH + 2 * col, R + (g * 8 + row) * ROW + 2 * col);
      }
    }
    cp_commit();
  };
  long g = (long)blockIdx.x * nw + warp;
  int buf = 0;
  if (g < ngroups) issue(g, 0);
  for (; g < ngroups; g += stride, buf ^= 1) {
    cp_wait0();
    __syncwarp();
    if (g + stride < ngroups) issue(g + stride, buf ^ 1);  // lands while this group is computed
    const double *sl = stage + buf * (2 * 8 * PITCH) + fr * PITCH, *sr = sl + 8 * PITCH;
    const long pb = g * 8 + fr, pa0 = g * 8 + 2 * fc;
    for (int k = 0; k < K; ++k) {
      double bl[KS], br[KS];
      if (pb < N) {
        const double2 a0 = *reinterpret_cast<const double2 *>(sl + k * S + 4 * fc), a1 = *reinterpret_cast<const double2 *>(sl + k * S + 4 * fc + 2);
        const double2 b0 = *reinterpret_cast<const double2 *>(sr + k * S + 4 * fc), b1 = *reinterpret_cast<const double2 *>(sr + k * S + 4 * fc + 2);
        bl[0] = a0.x; bl[1] = a0.y; bl[2] = a1.x; bl[3] = a1.y; bl[4] = sl[k * S + 16 + fc];
        br[0] = b0.x; br[1] = b0.y; br[2] = b1.x; br[3] = b1.y; br[4] = sr[k * S + 16 + fc];
      } else {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) bl[ks] = br[ks] = 0.0;
      }
      compute_class(fl, fr_, lane, k, bl, br, out + pa0 * ROW, out + (pa0 + 1) * ROW, pa0 < N, pa0 + 1 < N);
    }
    __syncwarp();  // all lanes are done reading `buf` before it is refilled two iterations later
  }
}

int main(int argc, char **argv) {
  const long N = argc > 1 ? atol(argv[1]) : 500000;
  const size_t bytes = (size_t)N * ROW * sizeof(double);
  double *L, *R, *oa, *ob, *Pl, *Pr;
  cudaMalloc(&L, bytes); cudaMalloc(&R, bytes); cudaMalloc(&oa, bytes); cudaMalloc(&ob, bytes);
  cudaMalloc(&Pl, K * S * S * 8); cudaMalloc(&Pr, K * S * S * 8);
  std::vector<double> h((size_t)N * ROW), hp(K * S * S);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0.01 + (double)((i * 2654435761u) % 1000) / 1000.0;
  cudaMemcpy(L, h.data(), bytes, cudaMemcpyHostToDevice);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0.02 + (double)((i * 40503u) % 997) / 997.0;
  cudaMemcpy(R, h.data(), bytes, cudaMemcpyHostToDevice);
  for (size_t i = 0; i < hp.size(); ++i) hp[i] = (double)((i * 7919u) % 101) / 2020.0;
  cudaMemcpy(Pl, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice);
  for (size_t i = 0; i < hp.size(); ++i) hp[i] = (double)((i * 104729u) % 103) / 2060.0;
  cudaMemcpy(Pr, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smA = sizeof(double) * 2 * K * MT * KS * 32, smB = smA + sizeof(double) * 8 * (2 * 2 * 8 * PITCH);
  cudaFuncSetAttribute(kern_reg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
  cudaFuncSetAttribute(kern_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char *name, auto launch) {
    for (int w = 0; w < 3; ++w) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 10; ++r) {
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("{\"kernel\": \"%s\", \"patterns\": %ld, \"ms\": %.4f, \"gbs\": %.1f, \"useful_tflops\": %.2f, \"err\": \"%s\"}\n", name, N, best,
           3.0 * bytes / best / 1e6, (double)N * K * S * (2.0 * (2 * S - 1) + 1) / best / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  run("reg (product scheme)", [&] { kern_reg<<<sms * 4, 256, smA>>>(Pl, Pr, L, R, oa, N); });
  run("smem (cp.async staged, double buffered)", [&] { kern_smem<<<sms, 256, smB>>>(Pl, Pr, L, R, ob, N); });
  std::vector<double> ha((size_t)N * ROW), hb((size_t)N * ROW);
  cudaMemcpy(ha.data(), oa, bytes, cudaMemcpyDeviceToHost);
  cudaMemcpy(hb.data(), ob, bytes, cudaMemcpyDeviceToHost);
  printf("{\"bit_identical\": %s}\n", memcmp(ha.data(), hb.data(), bytes) == 0 ? "true" : "false");
  return 0;
}
