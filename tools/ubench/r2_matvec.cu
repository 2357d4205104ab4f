// Microbenchmark: does sharing the (warp-uniform, shared-memory) matrix reads between R patterns per thread pay?
// Per step: 2 sides x 4 rate classes x 4x4 matvec + product, thread = R patterns. Reports pattern-updates/s.
#include <cstdio>
#include <cuda_runtime.h>
template <int R>
__global__ void __launch_bounds__(384, 1) bench(const double *__restrict__ gP, double *out, int nsteps, int reps) {
  extern __shared__ double sP[];
  for (int i = threadIdx.x; i < nsteps * 128; i += blockDim.x) sP[i] = gP[i];
  __syncthreads();
  double v[R][4][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) v[r][k][i] = 0.25 + 1e-3 * (threadIdx.x + k + i + r);
  for (int rep = 0; rep < reps; ++rep)
    for (int s = 0; s < nsteps; ++s) {
      const double2 *p = reinterpret_cast<const double2 *>(sP + s * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double x[R][4], y[R][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 a = p[k * 8 + 2 * i], b = p[k * 8 + 2 * i + 1], c = p[32 + k * 8 + 2 * i], d = p[32 + k * 8 + 2 * i + 1];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            x[r][i] = ((a.x * v[r][k][0] + a.y * v[r][k][1]) + b.x * v[r][k][2]) + b.y * v[r][k][3];
            y[r][i] = ((c.x * v[r][k][3] + c.y * v[r][k][2]) + d.x * v[r][k][1]) + d.y * v[r][k][0];
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) v[r][k][i] = x[r][i] * y[r][i];
      }
    }
  double acc = 0;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += v[r][k][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int R>
void run(const double *gP, double *out, int nsteps, int reps, int warps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(bench<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, nsteps * 1024);
  float best = 1e9;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    bench<R><<<148, warps * 32, nsteps * 1024>>>(gP, out, nsteps, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  printf("R=%d warps/SM=%2d: %.3f ms, %.1f G pattern-updates/s (%s)\n", R, warps, best,
         148.0 * warps * 32 * R * nsteps * reps / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int nsteps = 60, reps = 200;
  double *h = new double[nsteps * 128];
  for (int i = 0; i < nsteps * 128; ++i) h[i] = 0.7 + 0.001 * (i % 17);
  double *gP, *out;
  cudaMalloc(&gP, sizeof(double) * nsteps * 128);
  cudaMalloc(&out, sizeof(double) * 148 * 512);
  cudaMemcpy(gP, h, sizeof(double) * nsteps * 128, cudaMemcpyHostToDevice);
  for (int w : {4, 6, 8, 12}) run<1>(gP, out, nsteps, reps, w);
  for (int w : {3, 4, 6, 8}) run<2>(gP, out, nsteps, reps, w);
  for (int w : {2, 3, 4}) run<4>(gP, out, nsteps, reps, w);
  return 0;
}
