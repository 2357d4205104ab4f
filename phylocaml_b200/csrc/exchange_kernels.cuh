// The path's only exchange -- one scalar per evaluation -- done on the device over peer-mapped memory
// (NVLink / NVSwitch P2P stores), without NCCL and without a host round trip:
//
//   every rank owns a MAILBOX in its own HBM with one slot per rank; all mailboxes are mapped into
//   every rank (cudaIpc handles between processes, plain peer access inside one process). After a
//   rank's tree kernel has left its level-1 block partials (one double per 1024 patterns) in device
//   memory, a single CTA (a) copies them into slot[rank] of EVERY mailbox, then publishes a sequence
//   number there (release, system scope), (b) waits until all slots of its OWN mailbox carry that
//   sequence number (acquire), and (c) folds the concatenation of the ranks' partials, in rank order,
//   with the canonical 1024-fold -- the same tree a single GPU builds over the whole
//   alignment, so lnL is BIT-IDENTICAL for any rank count (shard boundaries are multiples of 1024
//   patterns). The result goes to mapped host memory; the host spins on the sequence number.
//
// Integer variant (Fitch / TCM lengths): one uint64 per rank, exact sum.
#pragma once
#include "common.cuh"

namespace phylo {

constexpr int kXMaxWorld = 16;
constexpr int kXMaxPartials = 8190;                      // per rank: 8 M patterns
constexpr int kXSlotDoubles = 2 + kXMaxPartials;          // [0] sequence number, [1] count, then the payload
// two sets of slots, used alternately (sequence number parity): a rank that has finished exchange q may start
// q + 1 while a slower peer is still reading the slots of q; it cannot get two ahead, because finishing q + 1
// needs that peer's q + 1 publication
constexpr size_t kXMailboxBytes = sizeof(double) * kXSlotDoubles * kXMaxWorld * 2;
constexpr long long kXTimeoutCycles = 4000000000ll;      // ~2 s: a peer that never arrives is an error, not a hang

struct XchgPeers {
  double *box[kXMaxWorld];  // mailbox of rank q as mapped into this process
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// mode 0: fp64 partials + canonical fold; mode 1: one uint64 per rank (bits in a double slot), integer sum.
// host_out (mapped): [0] result, [1] sequence number, or ~0 on a timeout.
__global__ void __launch_bounds__(1024) exchange_kernel(const double *__restrict__ mine, int n_mine, int rank, int world,
                                                        XchgPeers peers, unsigned long long seq, int mode,
                                                        volatile double *host_out) {
  __shared__ double vals[kLnlBlock];
  __shared__ double wsum[32];
  __shared__ double lvl2[kLnlBlock];
  __shared__ int cnt[kXMaxWorld + 1];
  __shared__ int timed_out;
  const int tid = threadIdx.x;
  const size_t set = (size_t)(seq & 1) * kXMaxWorld;
  if (tid == 0) timed_out = 0;
  // (a) publish
  for (int q = 0; q < world; ++q) {
    double *slot = peers.box[q] + (set + rank) * kXSlotDoubles;
    for (int i = tid; i < n_mine; i += blockDim.x) slot[2 + i] = mine[i];
  }
  // bar.sync makes the CTA's payload stores precede the publishing threads' release (cumulativity): one
  // system-scope release per peer instead of a system fence in each of the 1024 threads
  __syncthreads();
  if (tid < world) {
    unsigned long long *slot = (unsigned long long *)(peers.box[tid] + (set + rank) * kXSlotDoubles);
    slot[1] = (unsigned long long)n_mine;
    st_release_sys(slot, seq);
  }
  // (b) wait for every rank's slot of the own mailbox
  const double *own = peers.box[rank];
  if (tid < world) {
    const unsigned long long *slot = (const unsigned long long *)(own + (set + tid) * kXSlotDoubles);
    const long long t0 = clock64();
    while (ld_acquire_sys(slot) != seq) {
      if (clock64() - t0 > kXTimeoutCycles) { timed_out = 1; break; }
      __nanosleep(20);
    }
    cnt[tid + 1] = (int)ld_acquire_sys(slot + 1);
  }
  __syncthreads();  // the acquiring threads' view reaches the rest of the CTA through the barrier
  if (timed_out) {
    if (tid == 0) {
      host_out[0] = 0.0;
      __threadfence_system();
      const unsigned long long bad = ~0ull;
      host_out[1] = *(const double *)&bad;
    }
    return;
  }
  double result = 0.0;
  if (mode == 1) {
    if (tid == 0) {
      unsigned long long sum = 0;
      for (int r = 0; r < world; ++r) {
        const double v = ld_volatile_f64(own + (set + r) * kXSlotDoubles + 2);
        sum += *(const unsigned long long *)&v;
      }
      result = *(const double *)&sum;
    }
  } else {
    // (c) canonical fold over the concatenation, rank order
    if (tid == 0) {
      cnt[0] = 0;
      for (int r = 0; r < world; ++r) cnt[r + 1] += cnt[r];  // prefix offsets
    }
    __syncthreads();
    const int n_tot = cnt[world];
    const int nb = (n_tot + kLnlBlock - 1) / kLnlBlock;  // <= 128
    for (int b = 0; b < nb; ++b) {
      const int g = b * kLnlBlock + tid;
      double v = 0.0;
      if (g < n_tot) {
        int r = 0;
        while (g >= cnt[r + 1]) ++r;
        v = ld_volatile_f64(own + (set + r) * kXSlotDoubles + 2 + (g - cnt[r]));
      }
      vals[tid] = v;
      __syncthreads();
      const double f = block_fold_1024(vals, wsum);
      if (tid == 0) lvl2[b] = f;
      __syncthreads();
    }
    if (nb == 1) {
      result = lvl2[0];
    } else {
      vals[tid] = tid < nb ? lvl2[tid] : 0.0;
      __syncthreads();
      result = block_fold_1024(vals, wsum);
    }
  }
  if (tid == 0) {
    host_out[0] = result;
    __threadfence_system();
    host_out[1] = *(const double *)&seq;
  }
}

}  // namespace phylo
