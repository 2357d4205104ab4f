// Microbenchmark: per-step 2 sides x 4 rate classes x (4x4 matrix) x (4-vector) in fp64, thread = pattern,
// matrices warp-uniform. Variant C: matrices in __constant__ (LDCU -> uniform registers, DFMA Rx, URy);
// variant S: matrices in shared memory (broadcast LDS.128). Reports cycles per warp-step.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cP[8000];
template <bool CONST>
__global__ void __launch_bounds__(256, 1) bench(const double *__restrict__ gP, double *out, int nsteps, int reps) {
  extern __shared__ double sP[];
  if (!CONST) {
    for (int i = threadIdx.x; i < nsteps * 128; i += blockDim.x) sP[i] = gP[i];
    __syncthreads();
  }
  double v[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) v[k][i] = 0.25 + 1e-3 * (threadIdx.x + k + i);
  for (int r = 0; r < reps; ++r)
    for (int s = 0; s < nsteps; ++s) {
      const double *p = (CONST ? cP : sP) + s * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double x[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[i] = ((p[k * 16 + i * 4] * v[k][0] + p[k * 16 + i * 4 + 1] * v[k][1]) + p[k * 16 + i * 4 + 2] * v[k][2]) + p[k * 16 + i * 4 + 3] * v[k][3];
          y[i] = ((p[64 + k * 16 + i * 4] * v[k][3] + p[64 + k * 16 + i * 4 + 1] * v[k][2]) + p[64 + k * 16 + i * 4 + 2] * v[k][1]) + p[64 + k * 16 + i * 4 + 3] * v[k][0];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) v[k][i] = x[i] * y[i];
      }
    }
  double acc = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += v[k][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  const int nsteps = 60, reps = 200;
  double *h = new double[nsteps * 128];
  for (int i = 0; i < nsteps * 128; ++i) h[i] = 0.7 + 0.001 * (i % 17);
  double *gP, *out;
  cudaMalloc(&gP, sizeof(double) * nsteps * 128);
  cudaMalloc(&out, sizeof(double) * 148 * 512);
  cudaMemcpy(gP, h, sizeof(double) * nsteps * 128, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(cP, h, sizeof(double) * nsteps * 128);
  cudaFuncSetAttribute(bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, nsteps * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps = 4; warps <= 16; warps += 4) {
    for (int variant = 0; variant < 2; ++variant) {
      float best = 1e9;
      for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        if (variant == 0) bench<true><<<148, warps * 32, 0>>>(gP, out, nsteps, reps);
        else bench<false><<<148, warps * 32, nsteps * 1024>>>(gP, out, nsteps, reps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double warp_steps_per_sm = (double)warps * nsteps * reps;
      printf("%s warps/SM=%2d: %.3f ms, %.1f clk(nominal %d kHz) per warp-step per SM, %.2f G pattern-updates/s, err=%s\n",
             variant == 0 ? "const/LDCU" : "smem/LDS  ", warps, best, best * 1e-3 * clk * 1e3 / warp_steps_per_sm, clk,
             148.0 * warp_steps_per_sm * 32 / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
