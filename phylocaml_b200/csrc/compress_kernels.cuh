// Site-pattern compression: the step BEFORE the scoring path.
// Identical alignment columns are merged into one site pattern whose weight is the number (or
// weight sum) of the columns it stands for -- the `weights` the reference carries in
// NonAdditive_c.t (lib/nonAdditive_c.ml:3) and that both phylo_lk_set_tips and
// phylo_fitch_set_tips accept. The reference has no code for it.
//
// Device algorithm (all HBM-streaming integer work, no sort):
//   1. transpose the tip-major alignment [T][N] into site-major records [N][TP] (TP = T bytes
//      per element row, padded to 16) so that a column is one contiguous record;
//   2. hash every record (64 bit) and insert it into an open-addressing table of >= 2N slots:
//      the slot's key is claimed with atomicCAS, its representative is the SMALLEST site
//      index among the equal columns (atomicMin) -- so the result does not depend on the
//      order in which threads arrive;
//   3. verify: every site compares its full record with its representative's (a 64-bit hash
//      collision of two different columns makes the call retry with another seed);
//   4. representatives are numbered in site order by an exclusive prefix sum (patterns come
//      out ordered by first occurrence), weights are summed with fp64 atomics (exact for the
//      integer-valued weights the reference uses), the representatives' columns are gathered
//      back into a tip-major [T][P] array.
#pragma once
#include "common.cuh"

namespace phylo {

// [T][N] elements of EB bytes  ->  [N][TP] bytes (row t of a record at byte t*EB), 32x32 tiles
template <int EB>
__global__ void __launch_bounds__(256)
cmp_transpose_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ rec, int T, int64_t N, int TP) {
  __shared__ uint8_t tile[32][32 * EB + 4];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int t0 = blockIdx.y * 32;
  for (int r = ty; r < 32; r += 8) {  // rows = taxa, columns = sites (coalesced over sites)
    const int t = t0 + r;
    const int64_t s = s0 + tx;
#pragma unroll
    for (int b = 0; b < EB; ++b) tile[r][tx * EB + b] = (t < T && s < N) ? in[((int64_t)t * N + s) * EB + b] : 0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {  // rows = sites, columns = taxa (coalesced over taxa)
    const int64_t s = s0 + r;
    const int t = t0 + tx;
    if (s < N && t < T) {
#pragma unroll
      for (int b = 0; b < EB; ++b) rec[s * TP + (int64_t)t * EB + b] = tile[tx][r * EB + b];
    }
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

// one thread per site: 64-bit hash of its TP-byte record (TP multiple of 16), never 0
__global__ void __launch_bounds__(256)
cmp_hash_kernel(const uint8_t *__restrict__ rec, int64_t N, int TP, uint64_t seed, uint64_t *__restrict__ key) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
    const uint4 *p = reinterpret_cast<const uint4 *>(rec + s * TP);
    uint64_t h = seed;
    for (int i = 0; i < TP / 16; ++i) {
      const uint4 v = __ldg(p + i);
      h = mix64(h ^ (((uint64_t)v.y << 32) | v.x)) + 0x9e3779b97f4a7c15ull;
      h = mix64(h ^ (((uint64_t)v.w << 32) | v.z)) + 0x9e3779b97f4a7c15ull;
    }
    key[s] = h ? h : 1;
  }
}

// open addressing, linear probing; slot_of[s] = the slot that holds s's key
__global__ void __launch_bounds__(256)
cmp_insert_kernel(const uint64_t *__restrict__ key, int64_t N, unsigned long long *__restrict__ tkey,
                  int *__restrict__ trep, uint64_t mask, uint32_t *__restrict__ slot_of) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long h = key[s];
    uint64_t slot = mix64(h) & mask;
    for (;;) {
      const unsigned long long old = atomicCAS(&tkey[slot], 0ull, h);
      if (old == 0ull || old == h) break;
      slot = (slot + 1) & mask;
    }
    atomicMin(&trep[slot], (int)s);
    slot_of[s] = (uint32_t)slot;
  }
}

// every site against its representative; counts differing records (hash collisions)
__global__ void __launch_bounds__(256)
cmp_verify_kernel(const uint8_t *__restrict__ rec, int64_t N, int TP, const int *__restrict__ trep,
                  const uint32_t *__restrict__ slot_of, int *__restrict__ is_rep,
                  unsigned long long *__restrict__ n_collide) {
  unsigned long long bad = 0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
    const int r = trep[slot_of[s]];
    is_rep[s] = (r == (int)s);
    if (r != (int)s) {
      const uint4 *a = reinterpret_cast<const uint4 *>(rec + s * TP), *b = reinterpret_cast<const uint4 *>(rec + (int64_t)r * TP);
      bool same = true;
      for (int i = 0; i < TP / 16; ++i) {
        const uint4 x = __ldg(a + i), y = __ldg(b + i);
        same = same && x.x == y.x && x.y == y.y && x.z == y.z && x.w == y.w;
      }
      if (!same) ++bad;
    }
  }
  if (bad) atomicAdd(n_collide, bad);
}

// ---- exclusive prefix sum of int flags, three phases over blocks of 1024 elements
__global__ void __launch_bounds__(256)
cmp_scan_block_sums(const int *__restrict__ flag, int64_t N, int *__restrict__ bsum) {
  __shared__ int ws[8];
  const int64_t base = (int64_t)blockIdx.x * 1024;
  int v = 0;
  for (int i = threadIdx.x; i < 1024; i += 256) v += (base + i < N) ? flag[base + i] : 0;
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    bsum[blockIdx.x] = t;
  }
}
// one CTA: exclusive scan of the block sums in place, total to *total
__global__ void __launch_bounds__(1024)
cmp_scan_sums(int *__restrict__ bsum, int64_t nb, long long *__restrict__ total) {
  __shared__ long long carry;
  __shared__ int ws[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < nb ? bsum[i] : 0;
    int inc = v;  // inclusive scan within the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, off);
      if ((threadIdx.x & 31) >= off) inc += n;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = ws[threadIdx.x], winc = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, winc, off);
        if (threadIdx.x >= off) winc += n;
      }
      ws[threadIdx.x] = winc - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const long long c = carry;
    if (i < nb) bsum[i] = (int)(c + ws[threadIdx.x >> 5] + inc - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + ws[31] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
// pid[s] = exclusive prefix of flag (pattern number of a representative); rep_site[pid] = s
__global__ void __launch_bounds__(256)
cmp_scan_apply(const int *__restrict__ flag, int64_t N, const int *__restrict__ bsum, int *__restrict__ pid,
               int *__restrict__ rep_site) {
  __shared__ int ws[8];
  __shared__ int run;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  if (threadIdx.x == 0) run = bsum[blockIdx.x];
  __syncthreads();
  for (int chunk = 0; chunk < 4; ++chunk) {
    const int64_t s = base + chunk * 256 + threadIdx.x;
    const int v = s < N ? flag[s] : 0;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, off);
      if ((threadIdx.x & 31) >= off) inc += n;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += ws[w];
    const int ex = run + woff + inc - v;
    if (s < N) {
      pid[s] = ex;
      if (v) rep_site[ex] = (int)s;
    }
    __syncthreads();
    if (threadIdx.x == 255) run = ex + v;
    __syncthreads();
  }
}

// site -> pattern, weights
__global__ void __launch_bounds__(256)
cmp_assign_kernel(int64_t N, const int *__restrict__ trep, const uint32_t *__restrict__ slot_of,
                  const int *__restrict__ pid, const double *__restrict__ w_in, int *__restrict__ site_pat,
                  double *__restrict__ w_out) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
    const int p = pid[trep[slot_of[s]]];
    site_pat[s] = p;
    atomicAdd(&w_out[p], w_in ? w_in[s] : 1.0);
  }
}

// out[t][p] = in[t][rep_site[p]]   (tip-major both sides, EB bytes per element)
template <int EB>
__global__ void __launch_bounds__(256)
cmp_gather_kernel(const uint8_t *__restrict__ in, int T, int64_t N, const int *__restrict__ rep_site, int64_t P,
                  uint8_t *__restrict__ out) {
  const int64_t total = (int64_t)T * P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / P, p = i - t * P;
    const int64_t src = (t * N + rep_site[p]) * EB;
#pragma unroll
    for (int b = 0; b < EB; ++b) out[i * EB + b] = in[src + b];
  }
}

}  // namespace phylo
