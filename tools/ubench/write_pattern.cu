// Microbenchmark: HBM write bandwidth for the tree-fused kernel's store pattern -- every warp
// writes one CHUNK-byte piece to each of A arrays in turn (array stride 512 MB), then moves to
// its next piece -- against a plain linear stream of the same volume.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st256(void *p, const double4 &v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__global__ void __launch_bounds__(384, 1) pattern(double *base, size_t array_bytes, int arrays, size_t chunk, size_t pieces_per_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t gw = (size_t)blockIdx.x * nw + warp;
  const double4 v = make_double4(1.0, 2.0, 3.0, 4.0);
  for (size_t pc = 0; pc < pieces_per_warp; ++pc) {
    const size_t off = (gw * pieces_per_warp + pc) * chunk;
    for (int a = 0; a < arrays; ++a) {
      char *dst = (char *)base + (size_t)a * array_bytes + off;
      for (size_t b = lane * 32; b < chunk; b += 32 * 32) st256(dst + b, v);
    }
  }
}
__global__ void linear(double4 *base, size_t n) {
  const double4 v = make_double4(1.0, 2.0, 3.0, 4.0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st256(base + i, v);
}
int main() {
  const int arrays = 254;
  const size_t array_bytes = (size_t)512 << 20;  // 4 M patterns x 128 B
  double *buf;
  if (cudaMalloc(&buf, array_bytes * arrays) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int warps = 10, ctas = 148;
  for (size_t chunk : {(size_t)4096, (size_t)8192, (size_t)16384, (size_t)65536}) {
    const size_t pieces_total = array_bytes / chunk, ppw = pieces_total / ((size_t)ctas * warps);
    float best = 1e9;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      pattern<<<ctas, warps * 32>>>(buf, array_bytes, arrays, chunk, ppw);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double bytes = (double)ppw * ctas * warps * chunk * arrays;
    printf("pattern chunk %6zu B: %.2f ms, %.2f TB/s (%s)\n", chunk, best, bytes / best / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  for (size_t span : {(size_t)4 << 30, (size_t)32 << 30, array_bytes * arrays}) {
    const size_t n = span / 32;
    float best = 1e9;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      linear<<<148 * 16, 512>>>((double4 *)buf, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("linear stream over %.0f GB: %.2f ms, %.2f TB/s\n", span / 1e9, best, (double)n * 32 / best / 1e9);
  }
  return 0;
}
