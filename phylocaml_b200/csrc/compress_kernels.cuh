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
// In-place symbol -> state-mask translation of a 1-byte alignment (phylo_engine_set_symbol_table;
// every table entry fits a byte here). 16 bytes per thread; unknown symbols (entry 0) are counted.
__global__ void __launch_bounds__(256)
cmp_symbols_kernel(uint8_t *__restrict__ buf, size_t n, const uint64_t *__restrict__ lut,
                   unsigned long long *__restrict__ n_bad) {
  __shared__ uint8_t tab[256];
  tab[threadIdx.x] = (uint8_t)lut[threadIdx.x];
  __syncthreads();
  unsigned long long bad = 0;
  const size_t nvec = n / 16;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    uint4 w = reinterpret_cast<uint4 *>(buf)[v];
    uint32_t *x = reinterpret_cast<uint32_t *>(&w);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t m = tab[(x[q] >> (8 * b)) & 0xff];
        bad += (m == 0);
        o |= m << (8 * b);
      }
      x[q] = o;
    }
    reinterpret_cast<uint4 *>(buf)[v] = w;
  }
  if (blockIdx.x == 0)  // ragged tail
    for (size_t i = nvec * 16 + threadIdx.x; i < n; i += blockDim.x) {
      const uint8_t m = tab[buf[i]];
      bad += (m == 0);
      buf[i] = m;
    }
  if (bad) atomicAdd(n_bad, bad);
}

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
// Record hash = finalise(seed ^ sum over 32-bit words w of term(w, position)): the sum is
// commutative, so lanes can each take some words and combine by shuffles.
// A term is a 32-bit avalanche (murmur3 finaliser) of the position-salted word, widened by a
// position-dependent odd multiplier: ~10 integer instructions per 4 bytes (a splitmix64 per
// word cost three times that and made the transpose ALU-bound; a plain multiply-add term
// collided persistently on real alignments -- the verification caught it).
__device__ __forceinline__ uint64_t hash_term(uint32_t word, int pos) {
  const uint32_t salt = (uint32_t)(pos + 1) * 0x9e3779b1u;
  uint32_t x = word ^ salt;
  x ^= x >> 16; x *= 0x85ebca6bu;
  x ^= x >> 13; x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return (uint64_t)x * (salt | 1u) + ((uint64_t)x << 32);
}
__device__ __forceinline__ uint64_t hash_final(uint64_t seed, uint64_t sum) {
  const uint64_t h = mix64(seed ^ sum);
  return h ? h : 1;
}

// Fast path for 1-byte elements and N % 4 == 0: a thread turns a 4-taxa x 4-sites micro-tile
// (four aligned 32-bit loads, lanes along the sites: 128 contiguous bytes per row and warp)
// into four words "4 taxa of one site" with byte permutes, parks them in a site-major
// shared-memory tile, and the CTA copies whole records out (256 contiguous bytes per site).
constexpr int kCmpTileSites = 128, kCmpTileTaxa = 256, kCmpPitch = kCmpTileTaxa + 4;
__global__ void __launch_bounds__(256)
cmp_transpose4_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ rec, int T, int64_t N, int TP,
                      uint64_t seed, uint64_t *__restrict__ key) {  // key != NULL (one taxa tile): hash on the way out
  extern __shared__ __align__(16) uint8_t ttile[];  // [128 sites][kCmpPitch]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s0 = (int64_t)blockIdx.x * kCmpTileSites;
  const int t0 = blockIdx.y * kCmpTileTaxa;
  const int tt = min(kCmpTileTaxa, T - t0);              // taxa in this tile
  const int tq_n = (tt + 3) / 4;
  const int64_t s = s0 + 4 * lane;                       // this lane's four sites
  constexpr int U = 4;  // micro-tiles in flight per thread: 16 independent 128-byte row requests per warp
  for (int tq0 = warp; tq0 < tq_n; tq0 += 8 * U) {
    uint32_t r[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = t0 + 4 * (tq0 + 8 * u) + j;
        r[u][j] = (tq0 + 8 * u < tq_n && t < T && s < N) ? __ldg(reinterpret_cast<const uint32_t *>(in + (int64_t)t * N + s)) : 0u;
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int tq = tq0 + 8 * u;
      if (tq < tq_n) {
        const uint32_t a = __byte_perm(r[u][0], r[u][1], 0x5140), b = __byte_perm(r[u][2], r[u][3], 0x5140);
        const uint32_t c = __byte_perm(r[u][0], r[u][1], 0x7362), d = __byte_perm(r[u][2], r[u][3], 0x7362);
        uint32_t *row = reinterpret_cast<uint32_t *>(ttile + (size_t)(4 * lane) * kCmpPitch + 4 * tq);
        row[0] = __byte_perm(a, b, 0x5410);
        row[kCmpPitch / 4] = __byte_perm(a, b, 0x7632);
        row[2 * (kCmpPitch / 4)] = __byte_perm(c, d, 0x5410);
        row[3 * (kCmpPitch / 4)] = __byte_perm(c, d, 0x7632);
      }
    }
  }
  __syncthreads();
  const int words = tq_n;  // 32-bit words per record piece (padding taxa are zero)
  for (int i = warp; i < kCmpTileSites; i += 8) {
    const int64_t si = s0 + i;
    if (si >= N) break;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(ttile + (size_t)i * kCmpPitch);
    uint32_t *dst = reinterpret_cast<uint32_t *>(rec + si * TP + t0);
    uint64_t part = 0;
    for (int wd = lane; wd < words; wd += 32) {
      const uint32_t v = src[wd];
      dst[wd] = v;
      part += hash_term(v, wd);
    }
    if (key) {
      // the zero padding words of the record (positions words .. TP/4-1) belong to the hash too
      for (int wd = words + lane; wd < TP / 4; wd += 32) part += hash_term(0u, wd);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (lane == 0) key[si] = hash_final(seed, part);
    }
  }
}

// G lanes per site (G = power of two <= 32, about one lane per 16-byte chunk): 64-bit hash of the
// TP-byte record (TP multiple of 16), never 0. Lanes of a group read consecutive chunks.
template <int G>
__global__ void __launch_bounds__(256)
cmp_hash_kernel(const uint8_t *__restrict__ rec, int64_t N, int TP, uint64_t seed, uint64_t *__restrict__ key) {
  const int sub = threadIdx.x % G;
  const int64_t stride = (int64_t)gridDim.x * (256 / G);
  for (int64_t s0 = (int64_t)blockIdx.x * (256 / G); s0 < N; s0 += stride) {  // all lanes of a warp stay in the loop
    const int64_t s = s0 + threadIdx.x / G;
    uint64_t part = 0;
    if (s < N) {
      const uint4 *p = reinterpret_cast<const uint4 *>(rec + s * TP);
      for (int i = sub; i < TP / 16; i += G) {
        const uint4 v = __ldg(p + i);
        part += hash_term(v.x, 4 * i) + hash_term(v.y, 4 * i + 1) + hash_term(v.z, 4 * i + 2) + hash_term(v.w, 4 * i + 3);
      }
    }
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (s < N && sub == 0) key[s] = hash_final(seed, part);
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

// every site against its representative, G lanes per site (consecutive 16-byte chunks); counts
// differing records (hash collisions)
template <int G>
__global__ void __launch_bounds__(256)
cmp_verify_kernel(const uint8_t *__restrict__ rec, int64_t N, int TP, const int *__restrict__ trep,
                  const uint32_t *__restrict__ slot_of, int *__restrict__ is_rep,
                  unsigned long long *__restrict__ n_collide) {
  const int sub = threadIdx.x % G;
  const int64_t stride = (int64_t)gridDim.x * (256 / G);
  unsigned long long bad = 0;
  for (int64_t s0 = (int64_t)blockIdx.x * (256 / G); s0 < N; s0 += stride) {
    const int64_t s = s0 + threadIdx.x / G;
    int diff = 0;
    if (s < N) {
      const int r = trep[slot_of[s]];
      if (sub == 0) is_rep[s] = (r == (int)s);
      if (r != (int)s) {
        const uint4 *a = reinterpret_cast<const uint4 *>(rec + s * TP), *b = reinterpret_cast<const uint4 *>(rec + (int64_t)r * TP);
        for (int i = sub; i < TP / 16; i += G) {
          const uint4 x = __ldg(a + i), y = __ldg(b + i);
          diff |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
        }
      }
    }
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, off);
    if (s < N && sub == 0 && diff) ++bad;
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
