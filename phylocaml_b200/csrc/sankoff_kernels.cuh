// Cost-vector (Sankoff) parsimony under a general transformation cost matrix -- SURVEY 8(f) rank 4: the
// weighted-state generalisation of the Fitch path (lib/costMatrix.ml:68-124 works on state SETS; the
// cost-vector passes are only named in the reference: the commented-out bv_CAML_sankoff_median2_downpass /
// _uppass of lib/bitvector/bv.h:97-98). Per character and node a vector c[s] = cheapest cost of the subtree
// below the node given state s there; a median is a min-plus contraction of each child's vector with M:
//   c_p[s] = min_i (M[s][i] + c_l[i]) + min_j (M[s][j] + c_r[j]),
// a tip contributes 0 for the states in its mask and "infinity" elsewhere, and the tree length is
//   sum_chars w * min_{i,j} (c_a[i] + M[i][j] + c_b[j])   across the root edge (a, b).
// With the 0/1 matrix this is the Fitch length (the reference's own test property for its set-based
// medians, test/costMatrixTest.ml:110-125).
//
// Layout: node vectors as state planes [s][N] int32 (thread = character: every plane access coalesced), tips as
// one uint32 state mask per character, M [ST][ST] in shared memory (padded states carry "infinity").
// sankoff_tree_kernel evaluates a whole compiled schedule (build_fused_plan: TIP / CUR / POP / STORED operands)
// in ONE launch with the running vector in registers and parked vectors on a per-thread stack, reading only the
// tip masks (T * 4 bytes per character) -- the per-node kernel moves 3 * S * 4 bytes per character and node.
#pragma once
#include "common.cuh"

namespace phylo {

constexpr int kSkInf = 1 << 28;
constexpr int kSkMaxDepth = 12;

struct SkInstr {
  int kinds;  // lkind | rkind << 2 | push_first << 4 (OPK_*)
  int lidx, ridx;
  int out_slot;  // node slot whose vector is written (retain) or -1
};

template <int ST>
__device__ __forceinline__ void sk_tip(uint32_t mask, int (&c)[ST]) {
#pragma unroll
  for (int s = 0; s < ST; ++s) c[s] = ((mask >> s) & 1u) ? 0 : kSkInf;
}
template <int ST>
__device__ __forceinline__ void sk_load(const int *__restrict__ planes, int64_t N, int64_t n, int S, int (&c)[ST]) {
#pragma unroll
  for (int s = 0; s < ST; ++s) c[s] = s < S ? planes[(size_t)s * N + n] : kSkInf;
}
// r[s] = min_i (M[s][i] + c[i])
template <int ST>
__device__ __forceinline__ void sk_relax(const int *M, const int (&c)[ST], int (&r)[ST]) {
#pragma unroll
  for (int s = 0; s < ST; ++s) {
    int best = kSkInf;
#pragma unroll
    for (int i = 0; i < ST; ++i) best = min(best, M[s * ST + i] + c[i]);
    r[s] = best;
  }
}
template <int ST>
__device__ __forceinline__ void sk_median(const int *M, const int (&l)[ST], const int (&r)[ST], int (&out)[ST]) {
  int a[ST], b[ST];
  sk_relax<ST>(M, l, a);
  sk_relax<ST>(M, r, b);
#pragma unroll
  for (int s = 0; s < ST; ++s) out[s] = min(a[s] + b[s], kSkInf);
}
template <int ST>
__device__ __forceinline__ int sk_join(const int *M, const int (&a)[ST], const int (&b)[ST]) {
  int rb[ST];
  sk_relax<ST>(M, b, rb);
  int best = kSkInf;
#pragma unroll
  for (int s = 0; s < ST; ++s) best = min(best, a[s] + rb[s]);
  return best;
}
__device__ __forceinline__ void sk_accumulate(unsigned long long v, unsigned long long *total) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  __shared__ unsigned long long wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? wsum[lane] : 0ull;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0 && v) atomicAdd(total, v);
  }
}
template <int ST>
__device__ __forceinline__ void sk_stage_matrix(int *sM, const int *__restrict__ M, int S) {
  for (int idx = threadIdx.x; idx < ST * ST; idx += blockDim.x) {
    const int s = idx / ST, i = idx % ST;
    sM[idx] = (s < S && i < S) ? M[s * S + i] : kSkInf;
  }
  __syncthreads();
}

// one median (parent >= 0: vector stored, *total += sum_chars w * min_s c_p[s] when total != NULL) or the
// root-edge join (out == NULL: *total += sum_chars w * join)
template <int ST>
__global__ void __launch_bounds__(128)
sankoff_node_kernel(const int *__restrict__ M, int S, int64_t N, const uint32_t *__restrict__ ltip, const int *__restrict__ lvec,
                    const uint32_t *__restrict__ rtip, const int *__restrict__ rvec, int *__restrict__ out,
                    const uint32_t *__restrict__ w, unsigned long long *total) {
  __shared__ int sM[ST * ST];
  sk_stage_matrix<ST>(sM, M, S);
  unsigned long long acc = 0;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    int l[ST], r[ST];
    if (ltip) sk_tip<ST>(ltip[n], l); else sk_load<ST>(lvec, N, n, S, l);
    if (rtip) sk_tip<ST>(rtip[n], r); else sk_load<ST>(rvec, N, n, S, r);
    if (out) {
      int c[ST];
      sk_median<ST>(sM, l, r, c);
      int best = kSkInf;
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        if (s < S) out[(size_t)s * N + n] = c[s];
        best = min(best, c[s]);
      }
      acc += (unsigned long long)best * (w ? w[n] : 1u);
    } else {
      acc += (unsigned long long)sk_join<ST>(sM, l, r) * (w ? w[n] : 1u);
    }
  }
  if (total) sk_accumulate(acc, total);
}

// whole schedule in one launch; prog has n_steps medians + the root join
template <int ST, bool RETAIN>
__global__ void __launch_bounds__(128)
sankoff_tree_kernel(const int *__restrict__ M, int S, int64_t N, const uint32_t *__restrict__ tips, int64_t tip_stride,
                    const SkInstr *__restrict__ prog, int n_steps, int *const *__restrict__ node_vec,
                    const uint32_t *__restrict__ w, unsigned long long *total) {
  __shared__ int sM[ST * ST];
  sk_stage_matrix<ST>(sM, M, S);
  unsigned long long acc = 0;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    int stk[kSkMaxDepth][ST];
    int cur[ST];
    int sp = 0;
#pragma unroll
    for (int s = 0; s < ST; ++s) cur[s] = kSkInf;
    for (int step = 0; step <= n_steps; ++step) {
      const SkInstr in = prog[step];
      const int lk = in.kinds & 3, rk = (in.kinds >> 2) & 3;
      if (in.kinds & 16) {
#pragma unroll
        for (int s = 0; s < ST; ++s) stk[sp][s] = cur[s];
        ++sp;
      }
      int l[ST], r[ST];
      auto fetch = [&](int kind, int idx, int (&v)[ST]) {
        if (kind == OPK_TIP) sk_tip<ST>(tips[(size_t)idx * tip_stride + n], v);
        else if (kind == OPK_STORED) sk_load<ST>(node_vec[idx], N, n, S, v);
        else if (kind == OPK_CUR) {
#pragma unroll
          for (int s = 0; s < ST; ++s) v[s] = cur[s];
        } else {
          --sp;
#pragma unroll
          for (int s = 0; s < ST; ++s) v[s] = stk[sp][s];
        }
      };
      // a POP operand is the value parked most recently: fetch the CUR side first so the order does not matter
      fetch(lk, in.lidx, l);
      fetch(rk, in.ridx, r);
      if (step < n_steps) {
        sk_median<ST>(sM, l, r, cur);
        if (RETAIN && in.out_slot >= 0) {
          int *o = node_vec[in.out_slot];
#pragma unroll
          for (int s = 0; s < ST; ++s)
            if (s < S) o[(size_t)s * N + n] = cur[s];
        }
      } else {
        acc += (unsigned long long)sk_join<ST>(sM, l, r) * (w ? w[n] : 1u);
      }
    }
  }
  sk_accumulate(acc, total);
}

}  // namespace phylo
