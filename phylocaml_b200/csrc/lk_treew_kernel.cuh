// Tree-fused 4-state pruning, warp-autonomous version (v8+): thread = R PATTERNS (R = 1 or 2),
// all K rate classes each; warp = R groups of 32 consecutive patterns carried through the whole
// schedule. R = 2 shares every transition-matrix read between two patterns: the LSU register
// write-back of those (warp-uniform) reads is what bounds R = 1 (tools/ubench/r2_matvec.cu:
// 71 G pattern-updates/s at R = 1, 100-108 G at R = 2 with 4 or 8 warps per SM).
//
// Why (ncu on v5..v7, profiles/README.md): with thread = (pattern, rate class) every lane
// re-read its own two 4x4 matrices each step -- 4 distinct 16-byte chunks per LDS.128, i.e.
// 4 shared-memory wavefronts per instruction, 64 per warp-step for 16 pattern-updates; the
// L1/LSU pipe was 77-85 % busy while the fp64 pipe idled at 35 %. Here all 32 lanes of a
// warp need the SAME matrix element at the same time: every matrix read is a warp-uniform
// LDS.128 (one broadcast wavefront), 64 per warp-step for 32 pattern-updates -- half the
// LSU work per update, no shuffles (the K rate classes of a pattern sit in one thread, so
// the rescale test and the root mixture are plain register code).
//
//   * a warp is a self-contained worker: own tip buffer (bulk-TMA, one contiguous T*16-byte
//     copy per group thanks to the group-major tip layout, requested as soon as the previous
//     group's last tip byte is in registers), own mbarrier, own matrix ring (cp.async, two
//     steps ahead), own CLV stack, own store staging. There is NO CTA-wide barrier after the prologue;
//   * retained CLVs leave through TMA tensor stores: each lane writes its pattern's K*32
//     bytes into the warp's staging tile with the 128/64/32-byte swizzle (conflict-free
//     STS.128), one lane issues cp.async.bulk.tensor.2d (un-swizzles, clips rows >= N). The
//     LSU sees 1 wavefront per 128 bytes stored; an ordinary st.global at a 128-byte lane
//     stride would cost one per 32 bytes and split every line over 4 instructions;
//   * CLV stack in shared memory as 16-byte chunks [level][chunk][lane] (conflict-free); only
//     the first two levels -- deeper ones are rare (a tree needs level l about 4^-l as often)
//     and live in a small L2-resident global scratch, which buys a ninth warp per SM;
//
// Arithmetic and its order are those of prune4_kernel / root4_kernel / lk_tree4_kernel
// (same expressions => same bits): x_i = ((P_i0 v_0 + P_i1 v_1) + P_i2 v_2) + P_i3 v_3, the
// rate-class mixture summed as a butterfly ((l0+l1)+(l2+l3)).
#pragma once
#include "lk_tree_kernel.cuh"

namespace phylo {

constexpr int kTreeWMaxWarps = 12;  // 12 warps x 168 registers fill the register file (R = 1)
constexpr int kTreeWSlots = 3;      // matrix ring: the copy runs two steps ahead
constexpr int kTreeWSmemLevels = 2; // CLV stack levels kept in shared memory; deeper (rare) ones go to an L2-resident scratch

// One step of the compiled schedule for this kernel: 16 bytes.
struct __align__(16) TreeWInstr {
  int kinds;     // lkind | rkind << 2 | push_first << 4
  int lidx, ridx;  // tip row (OPK_TIP) or node slot (OPK_STORED)
  int out_slot;  // RETAIN: node slot that receives the result (tensor map + scale array), or -1
};

// Shared memory per warp: the store staging tiles (swizzled, need 1024-byte alignment: all
// warps' tiles form one block right after the program) and the rest (128-byte aligned).
__host__ __device__ inline size_t treew_stage_bytes(int K, bool retain, int R) {
  return retain ? (size_t)R * 32 * 32 * K : 0;
}
__host__ __device__ inline size_t treew_warp_bytes(int K, int T, int depth, int slev, int R) {
  if (depth > slev) depth = slev;
  size_t b = (size_t)R * T * 16;                      // tip buffer
  b += kTreeWSlots * 2 * (size_t)K * 128;             // matrix ring: slots x 2 sides
  b += (size_t)R * depth * 2 * K * 32 * 16;           // CLV stack
  b += (size_t)R * depth * 32 * 4;                    // scale-counter stack
  b += 64;                                            // mbarrier (+pad)
  return (b + 127) & ~(size_t)127;
}
// bytes of global scratch one warp needs for the stack levels beyond the shared-memory ones
__host__ __device__ inline size_t treew_spill_bytes(int K, int depth, int slev, int R) {
  return depth > slev ? (size_t)R * (depth - slev) * (2 * K * 32 * 16 + 128) : 0;
}
__host__ __device__ inline size_t treew_prog_bytes(int n_steps) {
  return (((size_t)(n_steps + 1) * sizeof(TreeWInstr)) + 1023) & ~(size_t)1023;
}

__device__ __forceinline__ void tma_store_2d(const void *tmap, const void *smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// The spilled stack levels are written and read back within microseconds while 5+ TB/s of CLV
// stores stream through the L2: they carry an evict_last policy so they stay resident instead
// of making a round trip to HBM (9 % of the DRAM traffic of the R = 2 retain run without it).
__device__ __forceinline__ uint64_t l2_keep_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_keep(double2 *p, double2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ double2 ld_keep(const double2 *p, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_keep_i32(int *p, int v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ int ld_keep_i32(const int *p, uint64_t pol) {
  int v;
  asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}

// x[r] = P v[r] for one rate class; pm = 8 double2 (row-major 4x4), warp-uniform address: each
// matrix row is read once and applied to the R vectors of the thread
template <int R>
__device__ __forceinline__ void matvec_u(const double2 *pm, const d4 (&v)[R], double (&x)[R][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 a = pm[2 * i], b = pm[2 * i + 1];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r][i] = ((a.x * v[r].x + a.y * v[r].y) + b.x * v[r].z) + b.y * v[r].w;
  }
}

constexpr int kTreeWFuseBytes = (kLnlBlock + 32) * 8;  // dynamic shared memory the fused final fold needs
constexpr int kTreeWInlineProg = 3072;  // bytes of program that can travel in the kernel parameters
struct TreeWArgs {
  TreeArgs t;              // shared fields (prog is a TreeWInstr array here)
  const char *tmaps;       // RETAIN: CUtensorMap per node slot (128 bytes each)
  // Small alignments (<= 1024 level-1 blocks): the last CTA to finish folds the groups into
  // the block partials and the final sum (the same canonical tree as fold_groups_kernel +
  // reduce1024_kernel) and publishes lnL + a sequence number into mapped host memory -- the
  // call is then pt_build + this kernel, and the host spins on the flag instead of a sync.
  int fuse_reduce, inline_prog;
  int64_t n_part;
  double *partials;                 // [n_part] level-1 block partials (kept for get_block_partials)
  unsigned int *done_counter;       // zero between calls
  volatile double *host_out;        // mapped: [0] = lnL, [1] = sequence number (as a double bit pattern)
  unsigned long long seq;
  __align__(16) unsigned char prog_inline[kTreeWInlineProg];
};

template <int K, int R, bool RETAIN>
__global__ void __launch_bounds__(R == 1 ? kTreeWMaxWarps * 32 : 256, 1) lk_treew_kernel(const TreeWArgs wa) {
  const TreeArgs &a = wa.t;
  constexpr int CH = 2 * K;                          // 16-byte chunks per pattern CLV
  constexpr int SWZ_SHIFT = (K == 4) ? 0 : (K == 2 ? 1 : 2);
  static_assert(K == 1 || K == 2 || K == 4, "rate classes per thread");
  static_assert(R == 1 || R == 2, "patterns per thread");

  extern __shared__ unsigned char smem_dyn[];
  unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n_steps = a.n_instr + 1;  // + root step
  int4 *sprog = reinterpret_cast<int4 *>(base);
  const int slev = min(a.stack_depth, a.smem_levels);
  const size_t sbytes = treew_stage_bytes(K, RETAIN, R), wbytes = treew_warp_bytes(K, a.T, a.stack_depth, a.smem_levels, R);
  unsigned char *ostage = base + treew_prog_bytes(n_steps) + (size_t)warp * sbytes;  // [R][32 rows][32K bytes], swizzled
  unsigned char *wb = base + treew_prog_bytes(n_steps) + (size_t)nwarps * sbytes + (size_t)warp * wbytes;
  uint8_t *tipbuf = wb;                                                 // [R][T*16]
  const uint32_t tip_bytes = (uint32_t)a.T * 16;
  double2 *ring = reinterpret_cast<double2 *>(tipbuf + (size_t)R * tip_bytes);  // [slots][2][K][8]
  double2 *stack = ring + kTreeWSlots * 2 * K * 8;                      // [slev][R][CH][32]
  int *stack_sc = reinterpret_cast<int *>(stack + (size_t)slev * R * CH * 32);  // [slev][R][32]
  uint64_t *bar = reinterpret_cast<uint64_t *>(stack_sc + (size_t)slev * R * 32);
  // levels >= slev: this warp's slice of the global scratch, per (level, r): [CH][32] chunks + [32] counters
  constexpr int SPILL_LEVEL = CH * 32 + 8;  // in double2 units
  double2 *spill = a.spill + ((size_t)blockIdx.x * kTreeWMaxWarps + warp) * (treew_spill_bytes(K, a.stack_depth, a.smem_levels, R) / 16);

  {
    const int4 *gprog = wa.inline_prog ? reinterpret_cast<const int4 *>(wa.prog_inline) : reinterpret_cast<const int4 *>(a.prog);
    for (int i = threadIdx.x; i < n_steps; i += blockDim.x) sprog[i] = gprog[i];
  }
  if (threadIdx.x == 0) sprog[n_steps] = make_int4(0, 0, 0, -1);  // harmless word past the end
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    fence_mbar_init();
  }
  __syncthreads();  // the only CTA-wide barrier

  // Work unit = R consecutive 32-pattern groups. The CTA owns a contiguous run of units; its
  // warps take them round-robin, so at any time the CTA's stores go to neighbouring pieces of a
  // node array.
  const int64_t ngroups = (a.N + 31) / 32;
  const int64_t u_begin = a.tile_begin / R, u_end = (min(ngroups, a.tile_end) + R - 1) / R;
  int64_t u_lo, u_hi;
  int u_step;
  if (a.interleave) {
    const int64_t per = (u_end - u_begin + gridDim.x - 1) / gridDim.x;
    u_lo = u_begin + (int64_t)blockIdx.x * per + warp;
    u_hi = min(u_end, u_begin + ((int64_t)blockIdx.x + 1) * per);
    u_step = nwarps;
  } else {
    const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, nw_total = (int64_t)gridDim.x * nwarps;
    const int64_t per = (u_end - u_begin + nw_total - 1) / nw_total;
    u_lo = u_begin + gw * per;
    u_hi = min(u_end, u_lo + per);
    u_step = 1;
  }
  const bool has_work = u_lo < u_hi;
  if (!has_work && !wa.fuse_reduce) return;
  if (has_work) {

  double pi[4], prob[K];
#pragma unroll
  for (int i = 0; i < 4; ++i) pi[i] = a.pi[i];
#pragma unroll
  for (int k = 0; k < K; ++k) prob[k] = a.probs[k];

  // the tip rows of the unit's R groups are contiguous in the group-major layout: one copy
  auto issue_tips = [&](int64_t u) {
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(&bar[0], R * tip_bytes);
      bulk_g2s(tipbuf, a.tips4 + (size_t)(u * R) * tip_bytes, R * tip_bytes, &bar[0]);
    }
  };
  // both matrix sets of `step` (contiguous in a.P, [side][k][4][4]) -> ring slot step % slots.
  // Always commits a group (an empty one past the root step) so wait_group counts stay uniform.
  auto fetch_matrices = [&](int step) {
    if (step <= a.n_instr) {
      const double2 *src = reinterpret_cast<const double2 *>(a.P) + (size_t)step * (2 * K * 8);
      double2 *dst = ring + (step % kTreeWSlots) * (2 * K * 8);
#pragma unroll
      for (int c = lane; c < 2 * K * 8; c += 32) cp_async16(dst + c, src + c);
    }
    cp_async_commit();
  };
  const int tb_byte = lane >> 1, tb_sh = (lane & 1) * 4;
  const int swz = (lane >> SWZ_SHIFT) & (CH - 1);
  // stack level `level` of pattern slice r: chunk c of this lane / the lane's scale counter.
  // Levels < slev live in shared memory, deeper ones in the global scratch (L2 evict_last).
  const uint64_t keep = l2_keep_policy();
  auto stack_put = [&](int level, int r, int c, double2 v) {
    if (level < slev) stack[(((size_t)level * R + r) * CH + c) * 32 + lane] = v;
    else st_keep(spill + ((size_t)(level - slev) * R + r) * SPILL_LEVEL + c * 32 + lane, v, keep);
  };
  auto stack_get = [&](int level, int r, int c) -> double2 {
    if (level < slev) return stack[(((size_t)level * R + r) * CH + c) * 32 + lane];
    return ld_keep(spill + ((size_t)(level - slev) * R + r) * SPILL_LEVEL + c * 32 + lane, keep);
  };
  auto stack_put_sc = [&](int level, int r, int v) {
    if (level < slev) stack_sc[((size_t)level * R + r) * 32 + lane] = v;
    else st_keep_i32(reinterpret_cast<int *>(spill + ((size_t)(level - slev) * R + r) * SPILL_LEVEL + CH * 32) + lane, v, keep);
  };
  auto stack_get_sc = [&](int level, int r) -> int {
    if (level < slev) return stack_sc[((size_t)level * R + r) * 32 + lane];
    return ld_keep_i32(reinterpret_cast<const int *>(spill + ((size_t)(level - slev) * R + r) * SPILL_LEVEL + CH * 32) + lane, keep);
  };

  issue_tips(u_lo);
  uint32_t seq = 0;
  for (int64_t u = u_lo; u < u_hi; u += u_step, ++seq) {
    __syncwarp();  // every lane is done with the previous unit's last ring slot
    fetch_matrices(0);
    fetch_matrices(1);
    int64_t pat[R];
    bool active[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      pat[r] = (u * R + r) * 32 + lane;
      active[r] = pat[r] < a.N;
    }
    mbar_wait(&bar[0], seq & 1);
    const uint8_t *tb = tipbuf;

    d4 cur[R][K];
    int cur_sc[R], sp = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cur_sc[r] = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) cur[r][k] = d4{0, 0, 0, 0};
    }

    int4 iw = sprog[0];   // kinds, lidx, ridx, out_slot
    int ml[R], mr[R];     // raw tip bytes of the coming step
    {
      const int lrow = ((iw.x & 3) == OPK_TIP) ? iw.y : 0, rrow = (((iw.x >> 2) & 3) == OPK_TIP) ? iw.z : 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ml[r] = tb[r * tip_bytes + lrow * 16 + tb_byte];
        mr[r] = tb[r * tip_bytes + rrow * 16 + tb_byte];
      }
    }
    int32_t *os = (RETAIN && iw.w >= 0) ? a.node_sc[iw.w] : nullptr;  // scale array of this step's result
    cp_async_wait<1>();  // step 0's matrices have landed
    if (a.n_instr == 0) {  // two-taxon tree: the root's tip bytes are already in registers
      __syncwarp();
      if (u + u_step < u_hi) issue_tips(u + u_step);
    }

    for (int step = 0; step < a.n_instr; ++step) {
      __syncwarp();  // all lanes have finished step-1: its ring slot is free, this step's copies are visible
      fetch_matrices(step + 2);  // two steps ahead (the root's matrix counts as step n_instr)
      const double2 *pmL = ring + (step % kTreeWSlots) * (2 * K * 8), *pmR = pmL + K * 8;
      const int lkind = iw.x & 3, rkind = (iw.x >> 2) & 3, push = iw.x & 16, lidx = iw.y, ridx = iw.z, oslot = iw.w;
      const int4 nw = sprog[step + 1];  // next step's word
      int32_t *os_next = (RETAIN && nw.w >= 0) ? a.node_sc[nw.w] : nullptr;
      if (push) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            stack_put(sp, r, 2 * k, make_double2(cur[r][k].x, cur[r][k].y));
            stack_put(sp, r, 2 * k + 1, make_double2(cur[r][k].z, cur[r][k].w));
          }
          stack_put_sc(sp, r, cur_sc[r]);
        }
        ++sp;
      }
      // ---- x = P_l L_l, y = P_r L_r for every rate class. Each operand kind has its own
      // copy of the matrix-vector code, so no operand is first copied into common registers.
      double x[K][R][4], y[K][R][4];
      int sc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) sc[r] = 0;
      auto side = [&](int kind, int idx, const int (&mbyte)[R], const double2 *pm, double (&o)[K][R][4]) {
        if (kind == OPK_TIP) {
          int m[R];
          bool hot = true;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            m[r] = (mbyte[r] >> tb_sh) & 15;
            hot = hot && (m[r] & (m[r] - 1)) == 0;
          }
          if (__all_sync(0xffffffffu, hot)) {
            // every lane's tip is a single state j: P L is column j of P, read directly (4 distinct
            // words of one 32-byte row per request: one wavefront, no arithmetic). 0/1 products and
            // additions of +0 are exact, so this is the general expression bit for bit.
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const double *col = reinterpret_cast<const double *>(pm) + (__ffs(m[r]) - 1);
#pragma unroll
              for (int k = 0; k < K; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) o[k][r][i] = col[k * 16 + i * 4];
            }
          } else {
            d4 t[R];
#pragma unroll
            for (int r = 0; r < R; ++r) t[r] = mask_vec(m[r]);
#pragma unroll
            for (int k = 0; k < K; ++k) matvec_u<R>(pm + k * 8, t, o[k]);
          }
        } else if (kind == OPK_CUR) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            d4 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = cur[r][k];
            matvec_u<R>(pm + k * 8, v, o[k]);
          }
#pragma unroll
          for (int r = 0; r < R; ++r) sc[r] += cur_sc[r];
        } else if (kind == OPK_POP) {
          --sp;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            d4 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const double2 p = stack_get(sp, r, 2 * k), q = stack_get(sp, r, 2 * k + 1);
              v[r] = d4{p.x, p.y, q.x, q.y};
            }
            matvec_u<R>(pm + k * 8, v, o[k]);
          }
#pragma unroll
          for (int r = 0; r < R; ++r) sc[r] += stack_get_sc(sp, r);
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            d4 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
              v[r] = active[r] ? ld256_stream(a.node_clv[idx] + pat[r] * (4 * K) + 4 * k) : d4{0, 0, 0, 0};
            matvec_u<R>(pm + k * 8, v, o[k]);
          }
#pragma unroll
          for (int r = 0; r < R; ++r) sc[r] += active[r] ? a.node_sc[idx][pat[r]] : 0;
        }
      };
      side(lkind, lidx, ml, pmL, x);
      side(rkind, ridx, mr, pmR, y);
      int h[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        h[r] = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          cur[r][k] = d4{x[k][r][0] * y[k][r][0], x[k][r][1] * y[k][r][1], x[k][r][2] * y[k][r][2], x[k][r][3] * y[k][r][3]};
          h[r] = max(h[r], max(max(hi32(cur[r][k].x), hi32(cur[r][k].y)), max(hi32(cur[r][k].z), hi32(cur[r][k].w))));
        }
      }
      // next step's tip bytes; consumed one iteration later
      {
        const int lrow = ((nw.x & 3) == OPK_TIP) ? nw.y : 0, rrow = (((nw.x >> 2) & 3) == OPK_TIP) ? nw.z : 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          ml[r] = tb[r * tip_bytes + lrow * 16 + tb_byte];
          mr[r] = tb[r * tip_bytes + rrow * 16 + tb_byte];
        }
      }
      iw = nw;
      if (step + 1 == a.n_instr) {  // the root's tip bytes are in registers: the tip buffer is free
        __syncwarp();
        if (u + u_step < u_hi) issue_tips(u + u_step);  // lands while the last step and the root join run
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (h[r] < kScaleHiThresh) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            cur[r][k].x *= 0x1p+256; cur[r][k].y *= 0x1p+256; cur[r][k].z *= 0x1p+256; cur[r][k].w *= 0x1p+256;
          }
          ++sc[r];
        }
        cur_sc[r] = sc[r];
      }
      if (RETAIN) {
        if (oslot >= 0) {
          if (lane == 0) bulk_wait_read0();  // the tiles stored last have left the staging buffers
          __syncwarp();
#pragma unroll
          for (int r = 0; r < R; ++r) {
            unsigned char *row = ostage + r * (32 * 32 * K) + lane * (32 * K);
#pragma unroll
            for (int k = 0; k < K; ++k) {
              *reinterpret_cast<double2 *>(row + (((2 * k) ^ swz) << 4)) = make_double2(cur[r][k].x, cur[r][k].y);
              *reinterpret_cast<double2 *>(row + (((2 * k + 1) ^ swz) << 4)) = make_double2(cur[r][k].z, cur[r][k].w);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const char *tmap = wa.tmaps + (size_t)oslot * 128;
#pragma unroll
            for (int r = 0; r < R; ++r)
              if ((u * R + r) * 32 < a.N) tma_store_2d(tmap, ostage + r * (32 * 32 * K), 0, (int)((u * R + r) * 32));
            bulk_commit();
          }
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (active[r]) os[pat[r]] = sc[r];
        }
        os = os_next;
      }
      cp_async_wait<1>();  // next step's matrices have landed (visible after the __syncwarp)
    }
    __syncwarp();
    // ---- root-edge join (root4_kernel's arithmetic): P applies to the b side only
    {
      const double2 *pm = ring + (a.n_instr % kTreeWSlots) * (2 * K * 8);
      const int akind = iw.x & 3, bkind = (iw.x >> 2) & 3;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        int c = 0;
        double l = 0.0, lk[K];
        // stack top (a POP operand of the root step), wherever it lives
        const int top = max(sp - 1, 0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          d4 av{0, 0, 0, 0}, bv[1] = {d4{0, 0, 0, 0}};
          // the POP operand (if any) was pushed before the CUR one was computed
          if (akind == OPK_TIP) av = mask_vec(ml[r] >> tb_sh);
          else if (akind == OPK_CUR) av = cur[r][k];
          else if (akind == OPK_POP) {
            const double2 p = stack_get(top, r, 2 * k), q = stack_get(top, r, 2 * k + 1);
            av = d4{p.x, p.y, q.x, q.y};
          } else if (active[r]) av = ld256_stream(a.node_clv[iw.y] + pat[r] * (4 * K) + 4 * k);
          if (bkind == OPK_TIP) bv[0] = mask_vec(mr[r] >> tb_sh);
          else if (bkind == OPK_CUR) bv[0] = cur[r][k];
          else if (bkind == OPK_POP) {
            const double2 p = stack_get(top, r, 2 * k), q = stack_get(top, r, 2 * k + 1);
            bv[0] = d4{p.x, p.y, q.x, q.y};
          } else if (active[r]) bv[0] = ld256_stream(a.node_clv[iw.z] + pat[r] * (4 * K) + 4 * k);
          double yy[1][4];
          matvec_u<1>(pm + k * 8, bv, yy);
          lk[k] = prob[k] * ((((pi[0] * av.x) * yy[0][0] + (pi[1] * av.y) * yy[0][1]) + (pi[2] * av.z) * yy[0][2]) + (pi[3] * av.w) * yy[0][3]);
        }
        if (akind == OPK_CUR) c += cur_sc[r];
        else if (akind == OPK_POP) c += stack_get_sc(top, r);
        else if (akind == OPK_STORED && active[r]) c += a.node_sc[iw.y][pat[r]];
        if (bkind == OPK_CUR) c += cur_sc[r];
        else if (bkind == OPK_POP) c += stack_get_sc(top, r);
        else if (bkind == OPK_STORED && active[r]) c += a.node_sc[iw.z][pat[r]];
        if (K == 1) l = lk[0];
        else if (K == 2) l = lk[0] + lk[1];
        else l = (lk[0] + lk[1]) + (lk[K > 2 ? 2 : 0] + lk[K > 2 ? 3 : 0]);
        double wl = 0.0;
        if (active[r]) {
          double lnl;
          if (a.pinvar >= 0.0) {
            const int m = a.inv[pat[r]];
            const double pv = (m & 1 ? pi[0] : 0.0) + (m & 2 ? pi[1] : 0.0) + (m & 4 ? pi[2] : 0.0) + (m & 8 ? pi[3] : 0.0);
            lnl = lnl_pinvar(l, c, a.pinvar, pv);
          } else {
            lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
          }
          if (a.site_lnl) a.site_lnl[pat[r]] = lnl;
          wl = (a.weights ? a.weights[pat[r]] : 1.0) * lnl;
        }
        // canonical level 0: a group of 32 consecutive patterns is one fold group
        const double gs = warp_fold(wl);
        if (lane == 0 && (u * R + r) < ngroups) a.groups[u * R + r] = gs;
      }
    }
  }
  if (RETAIN) {
    if (lane == 0) bulk_wait0();  // staged tiles must be read out before the CTA may exit
    __syncwarp();
  }
  }  // has_work
  if (wa.fuse_reduce) {
    // ---- last CTA: canonical fold of all group sums -> block partials -> lnL, straight to the host
    // (the warps' working storage is dead by now: the fold reuses the dynamic shared memory; the
    // host sizes it to at least kTreeWFuseBytes)
    __shared__ bool is_last;
    double *fvals = reinterpret_cast<double *>(base), *fwsum = fvals + kLnlBlock;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = atomicAdd(wa.done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      const int64_t n_groups = (a.N + 31) / 32;
      for (int b = threadIdx.x; b < kLnlBlock; b += blockDim.x) fvals[b] = 0.0;
      __syncthreads();
      for (int64_t b = warp; b < wa.n_part; b += nwarps) {
        const int64_t gi = b * 32 + lane;
        const double v = warp_fold(gi < n_groups ? __ldcg(a.groups + gi) : 0.0);
        if (lane == 0) {
          fvals[b] = v;
          wa.partials[b] = v;
        }
      }
      __syncthreads();
      const double r = block_fold_1024(fvals, fwsum);
      if (threadIdx.x == 0) {
        wa.host_out[0] = r;
        __threadfence_system();
        wa.host_out[1] = __longlong_as_double((long long)wa.seq);
        *wa.done_counter = 0;
        __threadfence_system();
      }
    }
  }
}

}  // namespace phylo
