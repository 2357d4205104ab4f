// Tree-fused 4-state pruning, warp-autonomous version (v8): thread = one PATTERN, all K rate
// classes; warp = one group of 32 consecutive patterns carried through the whole schedule.
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

constexpr int kTreeWMaxWarps = 12;  // 12 warps x 168 registers fill the register file
constexpr int kTreeWSlots = 3;      // matrix ring: the copy runs two steps ahead
constexpr int kTreeWSmemLevels = 2; // CLV stack levels kept in shared memory; deeper (rare) ones go to an L2-resident scratch

// bytes of shared memory one warp needs (multiple of 1024 so staging tiles stay 1024-aligned)
__host__ __device__ inline size_t treew_warp_bytes(int K, int T, int depth, bool retain, int slev, int obufs) {
  if (depth > slev) depth = slev;
  size_t b = retain ? (size_t)obufs * 32 * 32 * K : 0;  // store staging tiles (swizzled)
  b += (size_t)T * 16;                                // tip buffer
  b += kTreeWSlots * 2 * (size_t)K * 128;             // matrix ring: slots x 2 sides
  b += (size_t)depth * 2 * K * 32 * 16;               // CLV stack
  b += (size_t)depth * 32 * 4;                        // scale-counter stack
  b += 64;                                            // mbarrier (+pad)
  return (b + 1023) & ~(size_t)1023;
}
// bytes of global scratch one warp needs for the stack levels beyond kTreeWSmemLevels
__host__ __device__ inline size_t treew_spill_bytes(int K, int depth, int slev) {
  return depth > slev ? (size_t)(depth - slev) * (2 * K * 32 * 16 + 128) : 0;
}
__host__ __device__ inline size_t treew_prog_bytes(int n_steps) {
  return (((size_t)(n_steps + 1) * sizeof(TreeInstr)) + 1023) & ~(size_t)1023;
}

__device__ __forceinline__ void tma_store_2d(const void *tmap, const void *smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// x = P v for one rate class; pm = 8 double2 (row-major 4x4), warp-uniform address
__device__ __forceinline__ void matvec_u(const double2 *pm, const d4 &v, double (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 a = pm[2 * i], b = pm[2 * i + 1];
    x[i] = ((a.x * v.x + a.y * v.y) + b.x * v.z) + b.y * v.w;
  }
}

// In TreeInstr, out_clv carries the device address of the node's CUtensorMap (RETAIN).
template <int K, bool RETAIN>
__global__ void __launch_bounds__(kTreeWMaxWarps * 32, 1) lk_treew_kernel(const TreeArgs a) {
  constexpr int CH = 2 * K;                          // 16-byte chunks per pattern CLV
  constexpr int SWZ_SHIFT = (K == 4) ? 0 : (K == 2 ? 1 : 2);
  static_assert(K == 1 || K == 2 || K == 4, "rate classes per thread");

  extern __shared__ unsigned char smem_dyn[];
  unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n_steps = a.n_instr + 1;  // + root step
  int4 *sprog = reinterpret_cast<int4 *>(base);
  const size_t wbytes = treew_warp_bytes(K, a.T, a.stack_depth, RETAIN, a.smem_levels, a.obufs);
  unsigned char *wb = base + treew_prog_bytes(n_steps) + (size_t)warp * wbytes;
  unsigned char *ostage = wb;                                           // [2][32 rows][32K bytes], swizzled
  uint8_t *tipbuf = wb + (RETAIN ? a.obufs * 32 * 32 * K : 0);          // [T*16]
  const uint32_t tip_bytes = (uint32_t)a.T * 16;
  double2 *ring = reinterpret_cast<double2 *>(tipbuf + (size_t)tip_bytes);  // [slots][2][K][8]
  const int slev = min(a.stack_depth, a.smem_levels);
  double2 *stack = ring + kTreeWSlots * 2 * K * 8;                      // [slev][CH][32]
  int *stack_sc = reinterpret_cast<int *>(stack + (size_t)slev * CH * 32);  // [slev][32]
  uint64_t *bar = reinterpret_cast<uint64_t *>(stack_sc + (size_t)slev * 32);
  // levels >= slev: this warp's slice of the global scratch, per level [CH][32] chunks + [32] counters
  constexpr int SPILL_LEVEL = CH * 32 + 8;  // in double2 units
  double2 *spill = a.spill + ((size_t)blockIdx.x * kTreeWMaxWarps + warp) * (treew_spill_bytes(K, a.stack_depth, a.smem_levels) / 16);

  for (int i = threadIdx.x; i < 2 * n_steps; i += blockDim.x) sprog[i] = __ldg(reinterpret_cast<const int4 *>(a.prog) + i);
  if (threadIdx.x < 2) sprog[2 * n_steps + threadIdx.x] = make_int4(0, 0, 0, 0);  // harmless word past the end
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    fence_mbar_init();
  }
  __syncthreads();  // the only CTA-wide barrier

  // the CTA owns a contiguous run of 32-pattern groups; its warps take them round-robin, so at
  // any time the CTA's stores go to neighbouring 4K-byte pieces of a node array
  const int64_t ngroups = (a.N + 31) / 32;
  const int64_t g_end = min(ngroups, a.tile_end);
  int64_t g_lo, g_hi;
  int g_step;
  if (a.interleave) {
    const int64_t per = (g_end - a.tile_begin + gridDim.x - 1) / gridDim.x;
    g_lo = a.tile_begin + (int64_t)blockIdx.x * per + warp;
    g_hi = min(g_end, a.tile_begin + ((int64_t)blockIdx.x + 1) * per);
    g_step = nwarps;
  } else {
    const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, nw_total = (int64_t)gridDim.x * nwarps;
    const int64_t per = (g_end - a.tile_begin + nw_total - 1) / nw_total;
    g_lo = a.tile_begin + gw * per;
    g_hi = min(g_end, g_lo + per);
    g_step = 1;
  }
  if (g_lo >= g_hi) return;

  double pi[4], prob[K];
#pragma unroll
  for (int i = 0; i < 4; ++i) pi[i] = a.pi[i];
#pragma unroll
  for (int k = 0; k < K; ++k) prob[k] = a.probs[k];

  auto issue_tips = [&](int64_t g) {
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(&bar[0], tip_bytes);
      bulk_g2s(tipbuf, a.tips4 + (size_t)g * tip_bytes, tip_bytes, &bar[0]);
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

  issue_tips(g_lo);
  uint32_t seq = 0;
  uint32_t nstore = 0;  // retained tiles issued so far (staging buffer parity)
  for (int64_t g = g_lo; g < g_hi; g += g_step, ++seq) {
    __syncwarp();  // every lane is done with the previous group's last ring slot
    fetch_matrices(0);
    fetch_matrices(1);
    const int64_t pat = g * 32 + lane;
    const bool active = pat < a.N;
    mbar_wait(&bar[0], seq & 1);
    const uint8_t *tb = tipbuf;

    d4 cur[K];
#pragma unroll
    for (int k = 0; k < K; ++k) cur[k] = d4{0, 0, 0, 0};
    int cur_sc = 0, sp = 0;

    int4 iw = sprog[0];  // kinds, lidx, ridx, out_slot
    int ml, mr;          // raw tip bytes of the coming step
    {
      const int lrow = ((iw.x & 3) == OPK_TIP) ? iw.y : 0, rrow = (((iw.x >> 2) & 3) == OPK_TIP) ? iw.z : 0;
      ml = tb[lrow * 16 + tb_byte];
      mr = tb[rrow * 16 + tb_byte];
    }
    cp_async_wait<1>();  // step 0's matrices have landed
    if (a.n_instr == 0) {  // two-taxon tree: the root's tip bytes are already in registers
      __syncwarp();
      if (g + g_step < g_hi) issue_tips(g + g_step);
    }

    for (int step = 0; step < a.n_instr; ++step) {
      __syncwarp();  // all lanes have finished step-1: its ring slot is free, this step's copies are visible
      fetch_matrices(step + 2);  // two steps ahead (the root's matrix counts as step n_instr)
      const double2 *pmL = ring + (step % kTreeWSlots) * (2 * K * 8), *pmR = pmL + K * 8;
      const int lkind = iw.x & 3, rkind = (iw.x >> 2) & 3, push = iw.x & 16, lidx = iw.y, ridx = iw.z;
      const int4 ow = RETAIN ? sprog[2 * step + 1] : make_int4(0, 0, 0, 0);  // tensor map, out_sc
      const int4 nw = sprog[2 * step + 2];                                   // next step's word
      if (push) {
        if (sp < slev) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            stack[(sp * CH + 2 * k) * 32 + lane] = make_double2(cur[k].x, cur[k].y);
            stack[(sp * CH + 2 * k + 1) * 32 + lane] = make_double2(cur[k].z, cur[k].w);
          }
          stack_sc[sp * 32 + lane] = cur_sc;
        } else {
          double2 *lv = spill + (size_t)(sp - slev) * SPILL_LEVEL;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            lv[(2 * k) * 32 + lane] = make_double2(cur[k].x, cur[k].y);
            lv[(2 * k + 1) * 32 + lane] = make_double2(cur[k].z, cur[k].w);
          }
          reinterpret_cast<int *>(lv + CH * 32)[lane] = cur_sc;
        }
        ++sp;
      }
      // ---- x = P_l L_l, y = P_r L_r for every rate class. Each operand kind has its own
      // copy of the matrix-vector code, so no operand is first copied into common registers.
      double x[K][4], y[K][4];
      int sc = 0;
      auto side = [&](int kind, int idx, int mbyte, const double2 *pm, double (&o)[K][4]) {
        if (kind == OPK_TIP) {
          const int m = (mbyte >> tb_sh) & 15;
          if (__all_sync(0xffffffffu, (m & (m - 1)) == 0)) {
            // every lane's tip is a single state j: P L is column j of P, read directly (4 distinct
            // words of one 32-byte row per request: one wavefront, no arithmetic). 0/1 products and
            // additions of +0 are exact, so this is the general expression bit for bit.
            const double *col = reinterpret_cast<const double *>(pm) + (__ffs(m) - 1);
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
              for (int i = 0; i < 4; ++i) o[k][i] = col[k * 16 + i * 4];
          } else {
            const d4 t = mask_vec(m);
#pragma unroll
            for (int k = 0; k < K; ++k) matvec_u(pm + k * 8, t, o[k]);
          }
        } else if (kind == OPK_CUR) {
#pragma unroll
          for (int k = 0; k < K; ++k) matvec_u(pm + k * 8, cur[k], o[k]);
          sc += cur_sc;
        } else if (kind == OPK_POP) {
          --sp;
          if (sp < slev) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const double2 u = stack[(sp * CH + 2 * k) * 32 + lane], w = stack[(sp * CH + 2 * k + 1) * 32 + lane];
              matvec_u(pm + k * 8, d4{u.x, u.y, w.x, w.y}, o[k]);
            }
            sc += stack_sc[sp * 32 + lane];
          } else {
            const double2 *lv = spill + (size_t)(sp - slev) * SPILL_LEVEL;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const double2 u = lv[(2 * k) * 32 + lane], w = lv[(2 * k + 1) * 32 + lane];
              matvec_u(pm + k * 8, d4{u.x, u.y, w.x, w.y}, o[k]);
            }
            sc += reinterpret_cast<const int *>(lv + CH * 32)[lane];
          }
        } else {
          const double *src = a.node_clv[idx] + pat * (4 * K);
#pragma unroll
          for (int k = 0; k < K; ++k) matvec_u(pm + k * 8, active ? ld256_stream(src + 4 * k) : d4{0, 0, 0, 0}, o[k]);
          sc += active ? a.node_sc[idx][pat] : 0;
        }
      };
      side(lkind, lidx, ml, pmL, x);
      side(rkind, ridx, mr, pmR, y);
      int h = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        cur[k] = d4{x[k][0] * y[k][0], x[k][1] * y[k][1], x[k][2] * y[k][2], x[k][3] * y[k][3]};
        h = max(h, max(max(hi32(cur[k].x), hi32(cur[k].y)), max(hi32(cur[k].z), hi32(cur[k].w))));
      }
      // next step's tip bytes; consumed one iteration later
      {
        const int lrow = ((nw.x & 3) == OPK_TIP) ? nw.y : 0, rrow = (((nw.x >> 2) & 3) == OPK_TIP) ? nw.z : 0;
        ml = tb[lrow * 16 + tb_byte];
        mr = tb[rrow * 16 + tb_byte];
      }
      iw = nw;
      if (step + 1 == a.n_instr) {  // the root's tip bytes are in registers: the tip buffer is free
        __syncwarp();
        if (g + g_step < g_hi) issue_tips(g + g_step);  // lands while the last step and the root join run
      }
      if (h < kScaleHiThresh) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          cur[k].x *= 0x1p+256; cur[k].y *= 0x1p+256; cur[k].z *= 0x1p+256; cur[k].w *= 0x1p+256;
        }
        ++sc;
      }
      cur_sc = sc;
      if (RETAIN) {
        const uint64_t tmap = ((uint64_t)(uint32_t)ow.y << 32) | (uint32_t)ow.x;
        if (tmap != 0) {
          int32_t *os = reinterpret_cast<int32_t *>(((uint64_t)(uint32_t)ow.w << 32) | (uint32_t)ow.z);
          if (lane == 0) {  // the tile that used this staging buffer last has left it
            if (a.obufs == 2) bulk_wait_read1();
            else bulk_wait_read0();
          }
          __syncwarp();
          unsigned char *obuf = ostage + (nstore & (a.obufs - 1)) * (32 * 32 * K);
          ++nstore;
          unsigned char *row = obuf + lane * (32 * K);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            *reinterpret_cast<double2 *>(row + (((2 * k) ^ swz) << 4)) = make_double2(cur[k].x, cur[k].y);
            *reinterpret_cast<double2 *>(row + (((2 * k + 1) ^ swz) << 4)) = make_double2(cur[k].z, cur[k].w);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(reinterpret_cast<const void *>(tmap), obuf, 0, (int)(g * 32));
            bulk_commit();
          }
          if (active) os[pat] = sc;
        }
      }
      cp_async_wait<1>();  // next step's matrices have landed (visible after the __syncwarp)
    }
    __syncwarp();
    // ---- root-edge join (root4_kernel's arithmetic): P applies to the b side only
    {
      const double2 *pm = ring + (a.n_instr % kTreeWSlots) * (2 * K * 8);
      const int akind = iw.x & 3, bkind = (iw.x >> 2) & 3;
      int c = 0;
      double l = 0.0, lk[K];
      // stack top (a POP operand of the root step), wherever it lives
      const double2 *top = (sp - 1 < slev) ? stack + (size_t)max(sp - 1, 0) * CH * 32 : spill + (size_t)(sp - 1 - slev) * SPILL_LEVEL;
      const int *top_sc = (sp - 1 < slev) ? stack_sc + max(sp - 1, 0) * 32 : reinterpret_cast<const int *>(spill + (size_t)(sp - 1 - slev) * SPILL_LEVEL + CH * 32);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        d4 av{0, 0, 0, 0}, bv{0, 0, 0, 0};
        // the POP operand (if any) was pushed before the CUR one was computed
        if (akind == OPK_TIP) av = mask_vec(ml >> tb_sh);
        else if (akind == OPK_CUR) av = cur[k];
        else if (akind == OPK_POP) {
          const double2 u = top[(2 * k) * 32 + lane], w = top[(2 * k + 1) * 32 + lane];
          av = d4{u.x, u.y, w.x, w.y};
        } else if (active) av = ld256_stream(a.node_clv[iw.y] + pat * (4 * K) + 4 * k);
        if (bkind == OPK_TIP) bv = mask_vec(mr >> tb_sh);
        else if (bkind == OPK_CUR) bv = cur[k];
        else if (bkind == OPK_POP) {
          const double2 u = top[(2 * k) * 32 + lane], w = top[(2 * k + 1) * 32 + lane];
          bv = d4{u.x, u.y, w.x, w.y};
        } else if (active) bv = ld256_stream(a.node_clv[iw.z] + pat * (4 * K) + 4 * k);
        double y[4];
        matvec_u(pm + k * 8, bv, y);
        lk[k] = prob[k] * ((((pi[0] * av.x) * y[0] + (pi[1] * av.y) * y[1]) + (pi[2] * av.z) * y[2]) + (pi[3] * av.w) * y[3]);
      }
      if (akind == OPK_CUR) c += cur_sc;
      else if (akind == OPK_POP) c += top_sc[lane];
      else if (akind == OPK_STORED && active) c += a.node_sc[iw.y][pat];
      if (bkind == OPK_CUR) c += cur_sc;
      else if (bkind == OPK_POP) c += top_sc[lane];
      else if (bkind == OPK_STORED && active) c += a.node_sc[iw.z][pat];
      if (K == 1) l = lk[0];
      else if (K == 2) l = lk[0] + lk[1];
      else l = (lk[0] + lk[1]) + (lk[K > 2 ? 2 : 0] + lk[K > 2 ? 3 : 0]);
      double wl = 0.0;
      if (active) {
        double lnl;
        if (a.pinvar >= 0.0) {
          const int m = a.inv[pat];
          const double pv = (m & 1 ? pi[0] : 0.0) + (m & 2 ? pi[1] : 0.0) + (m & 4 ? pi[2] : 0.0) + (m & 8 ? pi[3] : 0.0);
          lnl = log((1.0 - a.pinvar) * ldexp(l, -kScaleExp * c) + a.pinvar * pv);
        } else {
          lnl = log(l) - (double)c * (kScaleExp * 0.6931471805599453094);
        }
        if (a.site_lnl) a.site_lnl[pat] = lnl;
        wl = (a.weights ? a.weights[pat] : 1.0) * lnl;
      }
      // canonical level 0: the warp's 32 consecutive patterns are one fold group
      const double gs = warp_fold(wl);
      if (lane == 0) a.groups[g] = gs;
    }
  }
  if (RETAIN) {
    if (lane == 0) bulk_wait0();  // staged tiles must be read out before the CTA may exit
    __syncwarp();
  }
}

}  // namespace phylo
