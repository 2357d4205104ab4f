// Fitch / non-additive parsimony and bitvector set-algebra kernels.
//
// Device layout: characters are bit-sliced. For each run of 32 characters
// ("word" w) and each state plane s < NP there is one 32-bit word buf[w*NP + s] whose bit c
// says "state s is in the set of character 32*w + c". DNA (NP=4) is therefore 16 bytes per
// 32 characters = 0.5 B/char and one 128-bit access per thread.
//
// The rule is the reference's bv_fitch (lib/bitvector/bv.c:148-160; same rule in
// lib/nonAdditive_c.ml:19-35):  m = a & b;  m == 0 ? (a | b, cost+1) : (m, cost+0),
// evaluated for 32 characters at once:  any = OR_s(a_s & b_s);  r_s = (a_s & b_s) | (~any &
// (a_s | b_s));  cost += popc(~any & valid).
#pragma once
#include "common.cuh"

namespace phylo {

template <int NP>
struct Planes {
  uint32_t v[NP];
};

template <int NP>
__device__ __forceinline__ Planes<NP> ld_planes(const uint32_t *buf, int64_t w) {
  Planes<NP> r;
  if (NP == 4) {
    const uint4 t = *(const uint4 *)(buf + w * 4);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int s = 0; s < NP; ++s) r.v[s] = buf[w * NP + s];
  }
  return r;
}
template <int NP>
__device__ __forceinline__ void st_planes(uint32_t *buf, int64_t w, const Planes<NP> &r) {
  if (NP == 4) {
    *(uint4 *)(buf + w * 4) = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
#pragma unroll
    for (int s = 0; s < NP; ++s) buf[w * NP + s] = r.v[s];
  }
}

__device__ __forceinline__ uint32_t valid_mask(int64_t w, int64_t N) {
  const int64_t rem = N - w * 32;
  return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// returns the changed-character mask
template <int NP>
__device__ __forceinline__ uint32_t fitch_rule(const Planes<NP> &a, const Planes<NP> &b,
                                               Planes<NP> &c) {
  uint32_t any = 0;
#pragma unroll
  for (int s = 0; s < NP; ++s) any |= a.v[s] & b.v[s];
#pragma unroll
  for (int s = 0; s < NP; ++s) c.v[s] = (a.v[s] & b.v[s]) | (~any & (a.v[s] | b.v[s]));
  return ~any;
}

__device__ __forceinline__ unsigned long long weighted_cost(uint32_t chg, int64_t w,
                                                            const uint32_t *__restrict__ wt) {
  unsigned long long c = 0;
  while (chg) {
    const int bit = __ffs(chg) - 1;
    chg &= chg - 1;
    c += wt[w * 32 + bit];
  }
  return c;
}

__device__ __forceinline__ void block_add_u64(unsigned long long v, unsigned long long *dst) {
  // warp shuffle -> one shared atomic per warp -> one global atomic per CTA (exact integers)
  __shared__ unsigned long long acc;
  if (threadIdx.x == 0) acc = 0;
  __syncthreads();
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(&acc, v);
  __syncthreads();
  if (threadIdx.x == 0 && acc) atomicAdd(dst, acc);
}

// measurement only (PHYLO_FITCH_TIMING=1): thread 0 of every CTA leaves %globaltimer stamps
__device__ __forceinline__ void fitch_stamp(unsigned long long *stamps, int i) {
  if (stamps != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[(size_t)blockIdx.x * 8 + i] = t;
  }
}

// ------------------------------------------------------- per-node median / distance ----
// STORE=true: bv_fitch (writes the parent set); STORE=false: bv_distance
// (lib/bitvector/bv.c:46-55). One thread per 32-character word, grid-strided.
template <int NP, bool STORE>
__global__ void __launch_bounds__(256)
fitch_median2_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                     uint32_t *__restrict__ c, int64_t nwords, int64_t N,
                     const uint32_t *__restrict__ wt, unsigned long long *__restrict__ cost) {
  unsigned long long local = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (int64_t)gridDim.x * blockDim.x) {
    const Planes<NP> pa = ld_planes<NP>(a, w), pb = ld_planes<NP>(b, w);
    Planes<NP> pc;
    const uint32_t chg = fitch_rule<NP>(pa, pb, pc) & valid_mask(w, N);
    if (STORE) st_planes<NP>(c, w, pc);
    local += wt ? weighted_cost(chg, w, wt) : (unsigned long long)__popc(chg);
  }
  block_add_u64(local, cost);
}

// run-time plane count (any alphabet up to 64 states)
template <bool STORE>
__global__ void __launch_bounds__(256)
fitch_median2_dyn_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                         uint32_t *__restrict__ c, int64_t nwords, int64_t N, int NP,
                         const uint32_t *__restrict__ wt, unsigned long long *__restrict__ cost) {
  unsigned long long local = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t any = 0;
    for (int s = 0; s < NP; ++s) any |= a[w * NP + s] & b[w * NP + s];
    if (STORE)
      for (int s = 0; s < NP; ++s) {
        const uint32_t x = a[w * NP + s], y = b[w * NP + s];
        c[w * NP + s] = (x & y) | (~any & (x | y));
      }
    const uint32_t chg = ~any & valid_mask(w, N);
    local += wt ? weighted_cost(chg, w, wt) : (unsigned long long)__popc(chg);
  }
  block_add_u64(local, cost);
}

// ---------------------------------------------------------------- whole-tree down-pass ----
// One launch for the entire post-order schedule: every thread carries its 32-character
// column through all n_ops medians and the root-edge join. A child that is an interior node
// was written earlier by this same thread, so no inter-thread synchronisation is needed and
// the re-read is served by L2. Node sets are written once (needed by the up-pass and by
// get_states). sched[o] = {parent, left, right} slot ids; slots -> buffers through `bufs`.
struct FitchStep {
  int parent, left, right;
};

template <int NP>
__global__ void __launch_bounds__(128)
fitch_tree_kernel(uint32_t *const *__restrict__ bufs, const FitchStep *__restrict__ sched, int n_ops,
                  int root_a, int root_b, int n_tips, int64_t nwords, int64_t N,
                  const uint32_t *__restrict__ wt, unsigned long long *__restrict__ node_cost,
                  unsigned long long *__restrict__ total) {
  extern __shared__ unsigned long long sh_cost[];  // n_ops + 1 per-CTA partial costs
  for (int o = threadIdx.x; o <= n_ops; o += blockDim.x) sh_cost[o] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int64_t w0 = (int64_t)blockIdx.x * blockDim.x; w0 < nwords;
       w0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = w0 + threadIdx.x;
    const bool act = w < nwords;
    const uint32_t valid = act ? valid_mask(w, N) : 0u;
    // (Measured alternatives, all slower than this plain walk: staging the tile's tip rows in
    // shared memory by bulk-TMA or cp.async and walking a compiled stack schedule on chip;
    // prefetching every tip row towards L2 first. See profiles/README.md.)
    for (int o = 0; o <= n_ops; ++o) {
      int il, ir, ip = -1;
      if (o < n_ops) {
        const FitchStep st = sched[o];
        il = st.left; ir = st.right; ip = st.parent;
      } else {
        il = root_a; ir = root_b;
      }
      unsigned long long c = 0;
      if (act) {
        const Planes<NP> pa = ld_planes<NP>(bufs[il], w), pb = ld_planes<NP>(bufs[ir], w);
        Planes<NP> pc;
        const uint32_t chg = fitch_rule<NP>(pa, pb, pc) & valid;
        if (ip >= 0) st_planes<NP>(bufs[ip], w, pc);
        c = wt ? weighted_cost(chg, w, wt) : (unsigned long long)__popc(chg);
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
      if (lane == 0 && c) atomicAdd(&sh_cost[o], c);
    }
  }
  __syncthreads();
  unsigned long long t = 0;
  for (int o = threadIdx.x; o <= n_ops; o += blockDim.x) {
    const unsigned long long c = sh_cost[o];
    if (c) {
      atomicAdd(&node_cost[o], c);
      t += c;
    }
  }
  if (t) atomicAdd(total, t);
}

// ------------------------------------------------ whole-tree down-pass, register walk ----
// Latency-optimised variant (default for <= 8 planes): the schedule is compiled on the host
// into a depth-first plan (build_fused_plan, shared with the likelihood tree kernels) whose
// operands are GLOBAL (a tip or a set that already exists in HBM), CUR (the previous result,
// still in registers) or POP (parked on a per-thread shared-memory stack). A thread never
// re-reads what it wrote, so the only loads left are the tips -- and those do not depend on
// any computation: they are issued D steps ahead into a register ring, D*2 128-bit loads in
// flight per thread instead of the 2 (behind two dependent pointer loads) of
// fitch_tree_kernel. At 1 M characters (31 k columns, ~7 warps per SM) that is the
// difference between a latency-bound 50 us and a bandwidth-shaped walk.
struct __align__(16) FitchInstr {
  int kinds;            // lkind | rkind << 2 | push_first << 4   (OPK_TIP doubles as "global operand")
  int pad_;
  const uint32_t *l, *r;  // global operands
  uint32_t *out;          // parent set, or NULL (root-edge join)
};
static_assert(sizeof(FitchInstr) == 32, "FitchInstr is two 16-byte words");

template <int NP, int D>
__global__ void __launch_bounds__(128)
fitch_treep_kernel(const FitchInstr *__restrict__ prog, int n_steps, int depth, int64_t nwords, int64_t N,
                   const uint32_t *__restrict__ wt, unsigned long long *__restrict__ node_cost,
                   unsigned long long *__restrict__ total) {
  extern __shared__ __align__(16) unsigned char fsm[];
  FitchInstr *sprog = reinterpret_cast<FitchInstr *>(fsm);                                 // [n_steps]
  unsigned long long *sh_cost = reinterpret_cast<unsigned long long *>(sprog + n_steps);   // [n_steps]
  uint32_t *stack = reinterpret_cast<uint32_t *>(sh_cost + n_steps + (n_steps & 1));       // [depth][NP][128]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 2 * n_steps; i += blockDim.x)
    reinterpret_cast<int4 *>(sprog)[i] = __ldg(reinterpret_cast<const int4 *>(prog) + i);
  for (int o = tid; o < n_steps; o += blockDim.x) sh_cost[o] = 0;
  __syncthreads();
  auto ldg_planes = [&](const uint32_t *buf, int64_t w) {
    Planes<NP> r;
    if (NP == 4) {
      const uint4 t = __ldg(reinterpret_cast<const uint4 *>(buf + w * 4));
      r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
#pragma unroll
      for (int s = 0; s < NP; ++s) r.v[s] = __ldg(buf + w * NP + s);
    }
    return r;
  };
  for (int64_t w0 = (int64_t)blockIdx.x * blockDim.x; w0 < nwords; w0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = w0 + tid;
    const bool act = w < nwords;
    const int64_t wc = act ? w : nwords - 1;  // inactive lanes load a valid column and discard it
    const uint32_t valid = act ? valid_mask(w, N) : 0u;
    Planes<NP> pl[D], pr[D], cur;
#pragma unroll
    for (int s = 0; s < NP; ++s) cur.v[s] = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i < n_steps) {
        const FitchInstr in = sprog[i];
        if ((in.kinds & 3) == OPK_TIP) pl[i] = ldg_planes(in.l, wc);
        if (((in.kinds >> 2) & 3) == OPK_TIP) pr[i] = ldg_planes(in.r, wc);
      }
    }
    int sp = 0;
    for (int base = 0; base < n_steps; base += D) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int step = base + i;
        if (step < n_steps) {
          const FitchInstr in = sprog[step];
          const int lkind = in.kinds & 3, rkind = (in.kinds >> 2) & 3;
          if (in.kinds & 16) {
#pragma unroll
            for (int s = 0; s < NP; ++s) stack[(sp * NP + s) * 128 + tid] = cur.v[s];
            ++sp;
          }
          Planes<NP> a, b, c;
          if (lkind == OPK_TIP) a = pl[i];
          else if (lkind == OPK_CUR) a = cur;
          else {
            --sp;
#pragma unroll
            for (int s = 0; s < NP; ++s) a.v[s] = stack[(sp * NP + s) * 128 + tid];
          }
          if (rkind == OPK_TIP) b = pr[i];
          else if (rkind == OPK_CUR) b = cur;
          else {
            --sp;
#pragma unroll
            for (int s = 0; s < NP; ++s) b.v[s] = stack[(sp * NP + s) * 128 + tid];
          }
          const uint32_t chg = fitch_rule<NP>(a, b, c) & valid;
          if (in.out != nullptr && act) st_planes<NP>(in.out, w, c);
          cur = c;
          if (wt) {
            unsigned long long cw = weighted_cost(chg, wc, wt);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) cw += __shfl_down_sync(0xffffffffu, cw, off);
            if (lane == 0 && cw) atomicAdd(&sh_cost[step], cw);
          } else {
            const unsigned cs = __reduce_add_sync(0xffffffffu, (unsigned)__popc(chg));
            if (lane == 0 && cs) atomicAdd(&sh_cost[step], (unsigned long long)cs);
          }
          if (step + D < n_steps) {  // refill this ring position with the operands of step + D
            const FitchInstr nx = sprog[step + D];
            if ((nx.kinds & 3) == OPK_TIP) pl[i] = ldg_planes(nx.l, wc);
            if (((nx.kinds >> 2) & 3) == OPK_TIP) pr[i] = ldg_planes(nx.r, wc);
          }
        }
      }
    }
  }
  __syncthreads();
  unsigned long long t = 0;
  for (int o = tid; o < n_steps; o += blockDim.x) {
    const unsigned long long c = sh_cost[o];
    if (c) {
      atomicAdd(&node_cost[o], c);
      t += c;
    }
  }
  if (t) atomicAdd(total, t);
}

// ----------------------------------------------- whole-tree down-pass, on-chip tiles ----
// The whole-tree kernel for 4 state planes (DNA). A CTA owns a tile of 32 columns; ALL input
// sets of the tile (tips and sets already resident) are fetched at once with cp.async (n_in
// independent 512-byte rows in flight per CTA), then the medians are evaluated out of shared
// memory in two phases: the host cuts the tree into whole subtrees of at most ~n/8 medians,
// deals them to the 8 warps (phase 1; a lane only ever touches its own column, so a warp runs
// its list without any synchronisation), and the few medians above the cut run on warp 0 after
// ONE __syncthreads (phase 2). (A first version synchronised once per tree level: 27 barriers
// for the bench tree, 12 k cycles per tile.) Results go to shared memory (for the parent) and to
// HBM (once). What sits on the chain of dependent medians is kept to the boolean rule itself:
// where an operand is the result of the warp's previous median it is taken from registers (the
// host flags it; the shared-memory copy is still written for any later reader), the table rows,
// the cost counter and the descriptor of the NEXT median are fetched before the current one is
// evaluated (descriptors run two ahead). Measured per median on the spine: 180 cycles when every
// operand made the shared-memory round trip, see profiles/README.md for this version.
// The last CTA to finish publishes the per-op costs straight into mapped host memory -- every
// 64-bit word carries the call's tag in its top 16 bits, so the host needs no ordering between
// words and the kernel no system-scope fence -- and re-zeroes the accumulators: one kernel
// launch is the whole call.
constexpr int kFitchTileWarps = 8;
constexpr int kFitchInlineProg = 3200;  // bytes of program that fit into the kernel parameters
constexpr int kFitchAccCopies = 16;  // CTAs spread their atomics over this many accumulator sets (L2 serialises same-address atomics)
struct __align__(16) FitchTileOp {
  uint32_t l_off, r_off;  // byte offsets of the operands' rows in the tile table; the result replaces the left row.
                          // l_off bit 0: the left operand is the result of the warp's previous median (registers)
  uint32_t *out;          // parent set in HBM, or NULL (root-edge join)
};
struct FitchTileArgs {
  const uint32_t *const *in_ptr;   // [n_in]
  const FitchTileOp *ops;          // [n_ops] phase-1 lists of warp 0, 1, ..., then phase 2 (the root join last)
  const int *task_start;           // [kFitchTileWarps + 2]: phase-1 op ranges per warp, then the phase-2 range
  int n_in, n_ops;
  int64_t nwords, N;
  const uint32_t *wt;
  unsigned long long *acc;         // [kFitchAccCopies][n_ops + 2]: per-op costs, (unused), CTA counter (all zero between calls)
  unsigned long long *host_out;    // mapped host memory [n_ops + 2]: per-op costs (tagged), (unused), then the call's sequence number
  unsigned long long seq;          // untagged protocol: written last, the host spins on it instead of a stream sync
  unsigned long long tag;          // != 0: every published word is cost | tag (tag = 16 bits << 48); no fences, no seq word
  int inline_prog;                 // 1: the program travels in `prog` (kernel parameter space), no H2D copy
  int prefetch_next;               // 1: a CTA has only a few tiles: prefetch the next tile's rows into L2
  unsigned long long *stamps;      // measurement only (PHYLO_FITCH_TIMING=1), else NULL
  __align__(16) unsigned char prog[kFitchInlineProg];  // ops | in_ptr | task_start
};

// Every operand is consumed exactly once (the host checks the schedule is a forest and gives
// every use of an input its own row), so a median is written over its left operand: the table
// holds only the n_in input rows (T x 512 bytes for a whole tree), 5-6 CTAs per SM for 64 taxa.
// CNT = uint16_t (unweighted: <= 32 per tile and op, <= 2047 tiles per CTA) or unsigned long long (weighted)
template <typename CNT>
__global__ void __launch_bounds__(256)
fitch_tile_kernel(const FitchTileArgs a) {
  extern __shared__ __align__(16) unsigned char tsm[];
  unsigned char *table = tsm;                                                              // [n_in][32] uint4
  FitchTileOp *sops = reinterpret_cast<FitchTileOp *>(table + (size_t)a.n_in * 512);       // [n_ops]
  // per-lane cost counters [n_ops][32]: plain read-modify-write by the owning lane -- no warp
  // reduction and no atomics on the per-level critical path (an op always maps to one warp)
  CNT *sh_cnt = reinterpret_cast<CNT *>(sops + a.n_ops);
  const uint32_t **sin = reinterpret_cast<const uint32_t **>(sh_cnt + (size_t)((a.n_ops + 3) & ~3) * 32);  // [n_in]
  __shared__ bool is_last;
  __shared__ int stask[kFitchTileWarps + 2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kFitchTileWarps;
  fitch_stamp(a.stamps, 0);
  const int64_t ntiles = (a.nwords + 31) / 32;
  const uint32_t mine = smem_u32(table) + lane * 16;                 // this lane's column inside any table row
  const uint32_t ops_s = smem_u32(sops);
  {
    const FitchTileOp *gops = a.inline_prog ? reinterpret_cast<const FitchTileOp *>(a.prog) : a.ops;
    const uint32_t *const *gin = a.inline_prog ? reinterpret_cast<const uint32_t *const *>(a.prog + sizeof(FitchTileOp) * a.n_ops) : a.in_ptr;
    const int *gts = a.inline_prog ? reinterpret_cast<const int *>(a.prog + sizeof(FitchTileOp) * a.n_ops + 8 * (size_t)a.n_in) : a.task_start;
    // the first tile's rows are requested before anything else: the program copy below runs under their latency
    if ((int64_t)blockIdx.x < ntiles) {
      const int64_t w = (int64_t)blockIdx.x * 32 + lane;
      for (int i = warp; i < a.n_in; i += nwarps)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mine + (uint32_t)i * 512u), "l"(gin[i] + w * 4) : "memory");
    }
    cp_async_commit();
    for (int i = tid; i < a.n_ops; i += blockDim.x) sops[i] = gops[i];
    for (int i = tid; i < a.n_in; i += blockDim.x) sin[i] = gin[i];
    if (tid < kFitchTileWarps + 2) stask[tid] = gts[tid];
  }
  for (int i = tid; i < a.n_ops * 32; i += blockDim.x) sh_cnt[i] = 0;
  __syncthreads();
  const int p1_lo = stask[warp], p1_hi = stask[warp + 1], p2_lo = stask[kFitchTileWarps], p2_hi = stask[kFitchTileWarps + 1];
  auto lds = [](uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
  };
  auto sts = [](uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  };
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t w = tile * 32 + lane;  // buffers are padded to whole tiles: always in bounds
    const uint32_t valid = valid_mask(w, a.N);
    if (tile != blockIdx.x) {
      for (int i = warp; i < a.n_in; i += nwarps)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mine + (uint32_t)i * 512u), "l"(sin[i] + w * 4) : "memory");
      cp_async_commit();
    }
    cp_async_wait<0>();
    __syncthreads();
    if (tile == blockIdx.x) fitch_stamp(a.stamps, 1);
    // few tiles per CTA (latency regime): the next tile's rows are pulled into L2 while this one is evaluated
    if (a.prefetch_next && tile + gridDim.x < ntiles) {
      const size_t off = (size_t)(tile + gridDim.x) * 512;  // a tile's share of a row: 32 columns x 16 bytes
      for (int i = tid; i < a.n_in * 4; i += blockDim.x)
        prefetch_l2(reinterpret_cast<const char *>(sin[i >> 2]) + off + (i & 3) * 128);
    }
    // a run of ops executed by this warp in order; op = {l_off, r_off, out lo, out hi}
    auto run = [&](int lo, int hi) {
      if (lo >= hi) return;
      const int last = hi - 1;
      uint4 op = lds(ops_s + lo * 16), nx = lds(ops_s + min(lo + 1, last) * 16);
      uint4 cur = make_uint4(0, 0, 0, 0);
      uint4 pa = lds(mine + op.x), pb = lds(mine + op.y);  // the first op of a run never takes the register operand
      CNT cn = sh_cnt[lo * 32 + lane];
      for (int o = lo; o <= last; ++o) {
        const uint4 nn = lds(ops_s + min(o + 2, last) * 16);
        const bool lcur = (op.x & 1u) != 0;
        const uint4 x = lcur ? cur : pa, y = pb;
        // the next median's rows and counter: none of them is written by this median (the host flags that case)
        if (!(nx.x & 1u)) pa = lds(mine + nx.x);
        pb = lds(mine + nx.y);
        const CNT cnext = sh_cnt[min(o + 1, last) * 32 + lane];
        const uint32_t any = (x.x & y.x) | (x.y & y.y) | (x.z & y.z) | (x.w & y.w);
        cur.x = (x.x & y.x) | (~any & (x.x | y.x));
        cur.y = (x.y & y.y) | (~any & (x.y | y.y));
        cur.z = (x.z & y.z) | (~any & (x.z | y.z));
        cur.w = (x.w & y.w) | (~any & (x.w | y.w));
        sts(mine + (op.x & ~1u), cur);
        if ((op.z | op.w) != 0u) {
          const unsigned long long out = ((unsigned long long)op.z | ((unsigned long long)op.w << 32)) + (unsigned long long)w * 16ull;
          asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out), "r"(cur.x), "r"(cur.y), "r"(cur.z), "r"(cur.w) : "memory");
        }
        const uint32_t chg = ~any & valid;
        if (sizeof(CNT) == 8) sh_cnt[o * 32 + lane] = cn + (CNT)weighted_cost(chg, w, a.wt);
        else sh_cnt[o * 32 + lane] = cn + (CNT)__popc(chg);
        cn = cnext;
        op = nx;
        nx = nn;
      }
    };
    run(p1_lo, p1_hi);
    __syncthreads();
    if (tile == blockIdx.x) fitch_stamp(a.stamps, 2);
    if (warp == 0) run(p2_lo, p2_hi);
    __syncthreads();  // the table is rewritten by the next tile's inputs
    if (tile == blockIdx.x) fitch_stamp(a.stamps, 3);
  }
  fitch_stamp(a.stamps, 4);
  // per-op totals: one warp per op folds its 32 lane counters (four ops at a time: the reductions overlap)
  unsigned long long *acc = a.acc + (size_t)(blockIdx.x % kFitchAccCopies) * (a.n_ops + 2);
  for (int o0 = warp; o0 < a.n_ops; o0 += 4 * nwarps) {
    unsigned long long c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int o = o0 + k * nwarps;
      c[k] = 0;
      if (o < a.n_ops) {
        if (sizeof(CNT) == 8) {
          c[k] = sh_cnt[o * 32 + lane];
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) c[k] += __shfl_down_sync(0xffffffffu, c[k], off);
        } else {
          c[k] = __reduce_add_sync(0xffffffffu, (unsigned)sh_cnt[o * 32 + lane]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane == 0 && c[k]) atomicAdd(&acc[o0 + k * nwarps], c[k]);
  }
  // ---- last CTA publishes to mapped host memory and restores the all-zero invariant.
  // bar.sync orders the CTA's atomics before thread 0's fence + counter increment
  // (cumulativity), so only one thread pays for the fence.
  __syncthreads();
  fitch_stamp(a.stamps, 5);
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(&a.acc[a.n_ops + 1], 1ull) == (unsigned long long)gridDim.x - 1;
  }
  __syncthreads();
  fitch_stamp(a.stamps, 6);
  if (is_last) {
    __threadfence();
    for (int o = tid; o < a.n_ops; o += blockDim.x) {
      unsigned long long v = 0;
#pragma unroll
      for (int c = 0; c < kFitchAccCopies; ++c) {
        v += __ldcg(&a.acc[(size_t)c * (a.n_ops + 2) + o]);
        a.acc[(size_t)c * (a.n_ops + 2) + o] = 0;
      }
      *reinterpret_cast<volatile unsigned long long *>(&a.host_out[o]) = v | a.tag;
    }
    if (tid == 0) a.acc[a.n_ops + 1] = 0;
    if (a.tag == 0) {  // untagged (weighted costs may need all 64 bits): results, system fence, then the sequence number
      __threadfence_system();
      __syncthreads();
      if (tid == 0) *reinterpret_cast<volatile unsigned long long *>(&a.host_out[a.n_ops + 1]) = a.seq;
    }
    fitch_stamp(a.stamps, 7);
  }
}

// ------------------------------------------------------------------------- up-pass ----
// Final sets, walking the schedule backwards (parents before children). Rule (SURVEY.md
// section 8 a11; the reference leaves Node.final_states TODO, lib/node.ml:260-268), per character
// with parent-final A, prelim P, child prelims L, R:
//   (P & A) == A -> A;  else (L & R) == 0 -> P | A;  else P | (A & (L | R)).
// The two ends of the root edge take A = root set = (a & b) if non-empty else (a | b).
// up[o] = {node, parent_or_-1, left, right}; leaves keep their observed sets.
struct FitchUpStep {
  int node, parent, left, right;
};

// Final-state rule of Fitch's second pass for one node (verified against exhaustive
// MPR sets in tests/golden/fitch_bruteforce.json): A = parent's final set, P = the node's own
// preliminary set, L / R = the children's preliminary sets.
//   A subset of P            -> A
//   else L & R empty (union) -> P | A
//   else                     -> P | (A & (L | R))
template <int NP>
__device__ __forceinline__ Planes<NP> fitch_final_rule(const Planes<NP> &P, const Planes<NP> &A, const Planes<NP> &L,
                                                       const Planes<NP> &R) {
  uint32_t notsub = 0, lr = 0;
#pragma unroll
  for (int s = 0; s < NP; ++s) {
    notsub |= A.v[s] & ~P.v[s];
    lr |= L.v[s] & R.v[s];
  }
  Planes<NP> F;
#pragma unroll
  for (int s = 0; s < NP; ++s) {
    const uint32_t f2 = P.v[s] | A.v[s];
    const uint32_t f3 = P.v[s] | (A.v[s] & (L.v[s] | R.v[s]));
    F.v[s] = (~notsub & A.v[s]) | (notsub & ((~lr & f2) | (lr & f3)));
  }
  return F;
}

// One node of the up-pass as its own kernel: NodeData.median_3 for the non-additive plugin (the
// reference only sketches it: `bv_CAML_fitch_median3(vb0, vb1, vb2, vb3)` is commented out in
// lib/bitvector/bv.h:94). P may alias nothing; out is a different buffer.
template <int NP>
__global__ void __launch_bounds__(256)
fitch_final1_kernel(const uint32_t *__restrict__ P, const uint32_t *__restrict__ A, const uint32_t *__restrict__ L,
                    const uint32_t *__restrict__ R, uint32_t *__restrict__ out, int64_t nwords) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    const Planes<NP> p = ld_planes<NP>(P, w), a = ld_planes<NP>(A, w), l = ld_planes<NP>(L, w), r = ld_planes<NP>(R, w);
    st_planes<NP>(out, w, fitch_final_rule<NP>(p, a, l, r));
  }
}

template <int NP>
__global__ void __launch_bounds__(128)
fitch_uppass_kernel(uint32_t *const *__restrict__ prelim, uint32_t *const *__restrict__ fin,
                    const FitchUpStep *__restrict__ up, int n_up, int root_a, int root_b,
                    int64_t nwords) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (int64_t)gridDim.x * blockDim.x) {
    Planes<NP> rootset;
    {
      const Planes<NP> a = ld_planes<NP>(prelim[root_a], w), b = ld_planes<NP>(prelim[root_b], w);
      fitch_rule<NP>(a, b, rootset);
    }
    for (int o = 0; o < n_up; ++o) {
      const FitchUpStep st = up[o];
      const Planes<NP> P = ld_planes<NP>(prelim[st.node], w);
      const Planes<NP> L = ld_planes<NP>(prelim[st.left], w), R = ld_planes<NP>(prelim[st.right], w);
      const Planes<NP> A = st.parent < 0 ? rootset : ld_planes<NP>(fin[st.parent], w);
      st_planes<NP>(fin[st.node], w, fitch_final_rule<NP>(P, A, L, R));
    }
  }
}

// -------------------------------------------- general-TCM median on state sets (Sankoff side) ----
// CostMatrix.find_median_general / find_median_metric (lib/costMatrix.ml:68-86, :107-124): for
// two state SETS a and b,  cost = min over i in a, j in b, k in candidates of M[i][k] + M[j][k]
// and the median is the set of all k that reach the minimum (candidates: every state, or the
// states of a | b in the metric variant). With <= 6 states the whole function is a table of
// 2^S x 2^S entries (cost | median << 16) built on the host from M; the kernel gathers each
// character's two masks out of the bit-sliced planes, looks the pair up in shared memory, and
// scatters the median back into planes. One thread per 32-character word.
template <int NP, bool STORE>
__global__ void __launch_bounds__(256)
tcm_median2_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t *__restrict__ c,
                   int64_t nwords, int64_t N, int S, const uint32_t *__restrict__ table,
                   const uint32_t *__restrict__ wt, unsigned long long *__restrict__ cost) {
  extern __shared__ uint32_t stab[];
  const int entries = 1 << (2 * S);
  for (int i = threadIdx.x; i < entries; i += blockDim.x) stab[i] = table[i];
  __syncthreads();
  const uint32_t smask = (1u << S) - 1u;
  unsigned long long local = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    const Planes<NP> pa = ld_planes<NP>(a, w), pb = ld_planes<NP>(b, w);
    Planes<NP> pc;
#pragma unroll
    for (int s = 0; s < NP; ++s) pc.v[s] = 0;
    const uint32_t valid = valid_mask(w, N);
    for (int ch = 0; ch < 32; ++ch) {
      if (!((valid >> ch) & 1)) break;  // valid characters are the low bits
      uint32_t ma = 0, mb = 0;
#pragma unroll
      for (int s = 0; s < NP; ++s) {
        ma |= ((pa.v[s] >> ch) & 1u) << s;
        mb |= ((pb.v[s] >> ch) & 1u) << s;
      }
      const uint32_t e = stab[((ma & smask) << S) | (mb & smask)];
      local += (unsigned long long)(e & 0xffffu) * (wt ? wt[w * 32 + ch] : 1u);
      if (STORE) {
        const uint32_t med = e >> 16;
#pragma unroll
        for (int s = 0; s < NP; ++s) pc.v[s] |= ((med >> s) & 1u) << ch;
      }
    }
    if (STORE) st_planes<NP>(c, w, pc);
  }
  block_add_u64(local, cost);
}

// ---------------------------------------------------------------- layout transcoding ----
// reference layout (one character per W-bit element, lib/bitvector/bv.h:29-55) <-> planes.
// A warp turns 32 consecutive elements into NP plane words with NP ballots.
template <typename Elt>
__global__ void __launch_bounds__(256)
fitch_encode_kernel(const Elt *__restrict__ codes, uint32_t *__restrict__ buf, int64_t N,
                    int64_t nwords, int NP, unsigned long long *__restrict__ n_bad,
                    const uint64_t *__restrict__ lut) {  // lut != NULL: 1-byte symbols -> state sets
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long bad = 0;
  for (int64_t w = warp0; w < nwords; w += nwarps) {
    const int64_t c = w * 32 + lane;
    uint64_t e = c < N ? (uint64_t)codes[c] : 0;
    if (lut && c < N) e = lut[e & 0xff];
    const uint64_t keep = NP >= 64 ? ~0ull : ((1ull << NP) - 1);
    if (c < N && (e & keep) == 0) ++bad;
    uint32_t mine = 0, mine2 = 0;
    for (int s = 0; s < NP; ++s) {
      const uint32_t bits = __ballot_sync(0xffffffffu, (e >> s) & 1);
      if ((s & 31) == lane) {
        if (s < 32) mine = bits; else mine2 = bits;
      }
    }
    if (lane < NP) buf[w * NP + lane] = mine;
    if (lane + 32 < NP) buf[w * NP + 32 + lane] = mine2;
  }
  if (bad) atomicAdd(n_bad, bad);
}

template <typename Elt>
__global__ void __launch_bounds__(256)
fitch_decode_kernel(const uint32_t *__restrict__ buf, Elt *__restrict__ codes, int64_t N,
                    int64_t nwords, int NP) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp0; w < nwords; w += nwarps) {
    uint64_t e = 0;
    for (int s = 0; s < NP; ++s) e |= (uint64_t)((buf[w * NP + s] >> lane) & 1u) << s;
    const int64_t c = w * 32 + lane;
    if (c < N) codes[c] = (Elt)e;
  }
}

// Bit-sliced planes uploaded as they are (phylo_fitch_set_tips with elt_bytes == 0): nothing to transcode,
// only the checks the encoder would have made: a character with no state at all is invalid input, and
// lanes beyond N are cleared. blockIdx.y = taxon.
__global__ void __launch_bounds__(256)
fitch_planes_check_kernel(uint32_t *const *__restrict__ bufs, int64_t nwords, int64_t N, int NP,
                          unsigned long long *__restrict__ n_bad) {
  uint32_t *buf = bufs[blockIdx.y];
  unsigned long long bad = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t valid = valid_mask(w, N);
    uint32_t any = 0;
    for (int s = 0; s < NP; ++s) any |= buf[w * NP + s];
    bad += __popc(~any & valid);
    if (valid != 0xffffffffu)
      for (int s = 0; s < NP; ++s) buf[w * NP + s] &= valid;
  }
  block_add_u64(bad, n_bad);
}

// ------------------------------------------------------------------ bitvector set ops ----
// bv_union / bv_inter (lib/bitvector/bv.c:93-99, :112-118): plane-wise OR / AND.
template <bool UNION>
__global__ void __launch_bounds__(256)
bv_binop_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                uint32_t *__restrict__ c, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    c[i] = UNION ? (a[i] | b[i]) : (a[i] & b[i]);
}

// mode 0: bv_popcount (bv.c:102-108): total number of set state bits.
// mode 1: bv_saturation (bv.c:121-129): characters whose set intersects `mask`.
// mode 2: bv_poly_saturation (bv.c:133-144): characters with exactly `n` states set.
__global__ void __launch_bounds__(256)
bv_count_kernel(const uint32_t *__restrict__ a, int64_t nwords, int64_t N, int NP, int mode,
                unsigned long long mask, int n, unsigned long long *__restrict__ out) {
  unsigned long long local = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t valid = valid_mask(w, N);
    if (mode == 0) {
      for (int s = 0; s < NP; ++s) local += __popc(a[w * NP + s] & valid);
    } else if (mode == 1) {
      uint32_t hit = 0;
      for (int s = 0; s < NP; ++s)
        if ((mask >> s) & 1) hit |= a[w * NP + s];
      local += __popc(hit & valid);
    } else {
      // per-character population count across planes: bit-serial counter in 7 bit-planes
      uint32_t cnt[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int s = 0; s < NP; ++s) {
        uint32_t carry = a[w * NP + s];
        for (int bpl = 0; bpl < 7 && carry; ++bpl) {
          const uint32_t t = cnt[bpl] & carry;
          cnt[bpl] ^= carry;
          carry = t;
        }
      }
      uint32_t eq = valid;
      for (int bpl = 0; bpl < 7; ++bpl) eq &= ((n >> bpl) & 1) ? cnt[bpl] : ~cnt[bpl];
      local += __popc(eq);
    }
  }
  block_add_u64(local, out);
}

// bv_compare (lib/bitvector/bv.c:72-89): lexicographic over characters on the element
// values. Finds the first differing character index (min-reduce); the host finishes.
__global__ void __launch_bounds__(256)
bv_firstdiff_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, int64_t nwords,
                    int64_t N, int NP, unsigned long long *__restrict__ first) {
  unsigned long long best = ~0ull;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t diff = 0;
    for (int s = 0; s < NP; ++s) diff |= a[w * NP + s] ^ b[w * NP + s];
    diff &= valid_mask(w, N);
    if (diff) best = min(best, (unsigned long long)(w * 32 + (__ffs(diff) - 1)));
  }
  if (best != ~0ull) atomicMin(first, best);
}

}  // namespace phylo
