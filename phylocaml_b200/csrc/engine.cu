// C ABI implementation (include/phylo_engine.h) of the B200 tree-scoring engine.
// Host-side handle, device arenas and kernel dispatch; all arithmetic lives in the kernels
// (lk_kernels.cuh: P(t), per-node pruning, DMMA kernels, root joins; lk_treew_kernel.cuh /
// lk_tree_kernel.cuh: single-launch tree-fused pruning; lk_edge_kernels.cuh: branch-length
// loop; fitch_kernels.cuh: Fitch / bitvector; compress_kernels.cuh: site-pattern compression).
// No CPU fallback anywhere: every entry point that computes launches CUDA kernels on the
// handle's device.
#include <cuda.h>  // CUtensorMap types only; the driver entry point is resolved at run time
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a tool is attached

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <string>
#include <numeric>
#include <vector>

#include "fitch_kernels.cuh"
#include "lk_kernels.cuh"
#include "lk_tree_kernel.cuh"
#include "lk_treew_kernel.cuh"
#include "lk_treem_kernel.cuh"
#include "lk_edge_kernels.cuh"
#include "compress_kernels.cuh"
#include "exchange_kernels.cuh"
#include "lk_grad_kernels.cuh"
#include "sankoff_kernels.cuh"
#include "phylo_engine.h"

using namespace phylo;

static std::string g_create_error;

// layout of the pinned (and device-mapped) scalar block phylo_engine::hScalar (128 doubles)
enum : int {
  HS_LNL = 0,        // lnL of the last evaluation (D2H copy target)
  HS_BAD = 1,        // invalid-mask counter read-back
  HS_TREE_OUT = 16,  // [0] lnL, [1] sequence number written by the tree kernel's last CTA (mapped)
  HS_EDGE_OUT = 24,  // 3 x kEdgeMaxT folded sums of phylo_lk_edge_eval (mapped)
  HS_XCHG_OUT = 80,  // [0] result, [1] sequence number of the device-side scalar exchange (mapped)
  HS_XCHG_IN = 84,   // the uint64 a rank contributes to an integer exchange (mapped, read by the kernel)
  HS_DOUBLES = 128
};

struct LkNode {
  double *clv = nullptr;
  int32_t *scale = nullptr;
  bool valid = false;
};

// kernel classes for the CUDA-event profiler (phylo_engine_profile_*)
enum KClass {
  KC_PT_BUILD = 0, KC_TREE_FUSED, KC_PRUNE_II, KC_PRUNE_TI, KC_PRUNE_TT, KC_ROOT, KC_REDUCE, KC_TIPS_PREPARE,
  KC_FITCH_TREE, KC_FITCH_NODE, KC_FITCH_UPPASS, KC_FITCH_TRANSCODE, KC_BV, KC_EDGE, KC_COMPRESS, KC_COUNT
};
static const char *kClassNames[KC_COUNT] = {
    "pt_build", "tree_fused", "prune_inner_inner", "prune_tip_inner", "prune_tip_tip", "root_lnl", "reduce1024",
    "tips_prepare", "fitch_tree", "fitch_median2", "fitch_uppass", "fitch_transcode", "bv_setops", "edge_loop", "compress"};

// compiled program of the last tile-kernel call (fitch_tile_compile) and what it depends on
struct FitchTileCache {
  bool valid = false, weighted = false, retain = true;
  std::vector<phylo_op> ops;
  int root_a = -1, root_b = -1;
  std::vector<int> pos;                  // schedule op -> position in the program (= index of its cost)
  std::vector<int> slots;                // every node slot the program reads or writes ...
  std::vector<const uint32_t *> ptrs;    // ... and the buffer it had when the program was compiled
  size_t smem = 0;
  FitchTileArgs a;
};

struct phylo_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = kSMs;

  // ---- profiler: one CUDA-event pair around every launch while enabled
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_pool;
  std::vector<int> prof_pending;  // class of pair i (events 2i, 2i+1)
  double prof_ms[KC_COUNT] = {0};
  uint64_t prof_n[KC_COUNT] = {0};

  // ---- likelihood model (MlModel.t, lib/mlModel.ml:53-63)
  int S = 0, K = 0;
  bool sym = false, has_model = false;
  double pinvar = -1.0;
  double *dU = nullptr, *dLam = nullptr, *dUi = nullptr, *dPi = nullptr, *dRates = nullptr,
         *dProbs = nullptr;
  // ---- likelihood data
  int T = 0, cap = 0, mask_dev_bytes = 1;
  int64_t N = 0;
  int opt_fitch_walk = 1;  // Fitch tree kernel: 0 = L2 walk, 1 = auto, 2 = register walk, 3 = on-chip tiles
  unsigned long long *dAcc = nullptr;  // tile kernel accumulators (all zero between calls)
  uint32_t *dTcm = nullptr;            // general-TCM median table: 2^S x 2^S entries (cost | median << 16)
  int tcmS = 0;
  size_t capAcc = 0, tileSmem = 0;
  uint32_t *dTipSlab = nullptr;            // Fitch: the T tip plane buffers as rows of one allocation (one 2-D upload)
  FitchTileCache tileCache;                // compiled tile-kernel program of the last fitch_score_tree call
  unsigned long long *hCostDev = nullptr;  // device view of the mapped hCost
  unsigned long long *dStamps = nullptr;   // PHYLO_FITCH_TIMING=1 only
  unsigned long long tileSeq = 0, treeSeq = 0;
  unsigned int *dTreeDone = nullptr;  // CTA counter of the fused final fold (zero between calls)
  bool fused_result_ready = false;    // the last fused evaluation already left lnL in hScalar[0]
  int tileOcc = 1;
  bool tileWeighted = false;
  int opt_fused = 1;  // 0 = one kernel per node, 1 = tree-fused (warp-autonomous where eligible), 2 = tile kernel only
  bool opt_retain = true;
  // TMA tensor maps of the node CLVs (warp-autonomous tree kernel stores through them)
  typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn encodeTiled = nullptr;
  CUtensorMap *dTmaps = nullptr;  // [cap]
  void *dSpill = nullptr;         // deep stack levels of the warp-autonomous tree kernel
  size_t capSpill = 0;
  bool tmapDirty = true;
  int64_t tipStride = 0;  // elements per tip row (N rounded up to 1024)
  double *dTT = nullptr;        // tip+tip table of finished vectors (20 / 61 states), S*S rows of K*S doubles
  int32_t *dTTsc = nullptr;     // their scale counters
  size_t capTT = 0;
  int opt_tt_table = 1;         // PHYLO_TT_TABLE=0 keeps the DMMA kernel for tip+tip (A/B switch)
  uint64_t *dSymTab = nullptr;  // 256 state masks by symbol byte (phylo_engine_set_symbol_table) or NULL
  bool symtab_fits_byte = false;
  size_t hostPitch = 0;   // bytes between taxon rows of the host alignment being uploaded (0: N * mask_bytes)
  double **dNodeClv = nullptr;   // device tables of node buffers (tree-fused kernel)
  int32_t **dNodeSc = nullptr;
  bool nodeTabDirty = true;
  void *dProg = nullptr, *hProg = nullptr;
  size_t capProg = 0;
  cudaStream_t copyStream = nullptr;   // H2D of alignment slabs, overlapped with compute
  cudaStream_t copyStream2 = nullptr;  // odd slabs: two DMA queues keep PCIe busier than one
  bool tips8_valid = false;            // dTips (one byte per cell) is current; false after a packed upload until needed
  cudaStream_t auxStream = nullptr;    // odd slabs compute here so slab kernels overlap their tails
  cudaEvent_t auxDone = nullptr;
  std::vector<cudaEvent_t> slabEvents;
  void *dRaw = nullptr;   // raw tip upload staging (kept across set_tips calls)
  size_t capRaw = 0;
  unsigned long long *dBad = nullptr;
  uint8_t *dTips4 = nullptr;  // 4-state only: nibble-packed copy of dTips for the tree-fused kernel
  void *dTips = nullptr;  // T*tipStride masks, device width
  void *dInv = nullptr;   // N masks: AND over tips
  double *dWeights = nullptr;
  std::vector<LkNode> nodes;
  // node-slot allocator behind phylo_lk_node_alloc / _release (the OCaml custom blocks' finalizers
  // return slots here): released slots keep their device buffers, so a search loop does no cudaMalloc
  std::vector<char> lkOwned;
  std::vector<int> lkFree;
  int lkNext = 0, cap0 = 0;        // next never-used slot; the capacity the caller asked for
  uint64_t lkGen = 0;              // bumped whenever every slot is dropped (new alignment shape / model alphabet)
  double *dP = nullptr;  // transition matrices [branch][K][S][S]
  std::vector<double> ptLastT;  // branch lengths e->dP was last built for (cleared when the model or dP changes)
  int ptLastInterleave = -1;
  size_t capP = 0;       // branches
  // ---- Sankoff (cost-vector parsimony, sankoff_kernels.cuh)
  int skT = 0, skCap = 0, skS = 0;
  int64_t skN = 0;
  uint32_t *dSkTips = nullptr, *dSkW = nullptr;  // [T][N] state masks; integer weights or NULL
  int *dSkM = nullptr;                           // [S][S]
  std::vector<int *> skVec;                      // per slot: [S][N] cost planes (interior slots, on demand)
  std::vector<char> skValid;
  int **dSkTab = nullptr;
  bool skTabDirty = true, skHasM = false;
  unsigned long long *dSkTotal = nullptr;
  // host copy of the eigensystem (Q = V diag(lam) Vinv) for the model-parameter derivatives
  std::vector<double> hV, hVinv, hLam, hPi, hRates, hProbs;
  // device-side scalar exchange (exchange_kernels.cuh)
  double *xMailbox = nullptr;   // this engine's mailbox (kXMailboxBytes)
  XchgPeers xPeers{};           // every rank's mailbox as mapped here
  std::vector<void *> xOpened;  // mappings opened from IPC handles (closed on destroy)
  int xWorld = 0, xRank = -1;
  unsigned long long xSeq = 0;
  bool defer_scalar = false;    // PHYLO_OPT_DEFER_SCALAR
  double *dFrag = nullptr;  // the same matrices as DMMA A-fragment tables (tree-fused 20/61-state kernel)
  size_t capFrag = 0;       // doubles
  double *dT = nullptr, *hT = nullptr;  // branch lengths (device / pinned)
  double *dSite = nullptr, *dWSite = nullptr;
  // branch-length loop (lk_edge_kernels.cuh): eigenvector matrices in the orientation the sum table needs,
  // the sum table of the prepared edge, per-(t, derivative) block partials
  double *dUL = nullptr, *dUR = nullptr, *dSum = nullptr, *dEdgePart = nullptr, *dEdgeOut = nullptr, *dEdgeT = nullptr;
  char *dEdgeArena = nullptr;  // lk_edge_lnl_batch: edge descriptors + fold levels
  size_t capEdgeArena = 0;
  double *hEdgeRes = nullptr;  // pinned results of lk_edge_lnl_batch, grow-only (a pinned allocation per call cost more than the joins of a small neighbourhood)
  int capEdgeRes = 0;
  int32_t *dSumSc = nullptr;
  bool edge_ready = false;
  int edge_a = -1, edge_b = -1;
  double *dGroups = nullptr;                   // per-32-pattern sums (tree-fused kernel)
  double *dPart = nullptr, *dPart2 = nullptr;  // reduction levels
  int64_t nPart = 0;
  double *hScalar = nullptr;  // pinned
  bool lk_evaluated = false;

  // ---- Fitch data
  int fT = 0, fcap = 0, fNP = 0, fNPdev = 0, felt = 1;
  int64_t fN = 0, fWords = 0;
  std::vector<uint32_t *> fPre, fFin;
  std::vector<char> fValid;        // slot holds preliminary sets of the LOADED alignment (not a stale buffer)
  std::vector<char> fFinValid;     // same for the final (up-pass) sets
  std::vector<char> fOwned;        // allocator state, as for the likelihood slots
  std::vector<int> fFreeList;
  int fNext = 0, fcap0 = 0;
  uint64_t fGen = 0;
  uint32_t **dPreTab = nullptr, **dFinTab = nullptr;
  bool tabDirty = true;
  uint32_t *dFW = nullptr;  // integer weights, padded to fWords*32
  unsigned long long *dCost = nullptr;  // [cap+2]: per-op costs, then total
  size_t capCost = 0;
  void *dSched = nullptr;
  void *hSched = nullptr;
  size_t capSched = 0;
  unsigned long long *hCost = nullptr;  // pinned mirror of dCost
  std::vector<uint64_t> nodeCost;
  void *dStage = nullptr;  // transcoding staging (one node in reference layout)
  size_t capStage = 0;
};

// ----------------------------------------------------------------------- helpers ----
static int fail(phylo_engine *e, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t err_ = (call);                                                           \
    if (err_ != cudaSuccess)                                                             \
      return fail(e, PHYLO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err_), \
                  __FILE__, __LINE__);                                                   \
  } while (0)

#define LAUNCH_CHECK()                                                                   \
  do {                                                                                   \
    ++e->launches;                                                                       \
    cudaError_t err_ = cudaGetLastError();                                               \
    if (err_ != cudaSuccess)                                                             \
      return fail(e, PHYLO_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                 \
                  cudaGetErrorString(err_), __FILE__, __LINE__);                         \
  } while (0)

// RAII: an NVTX range named after the kernel class around the launches in its scope (nsys / ncu --nvtx
// timelines show "tree_fused", "prune_inner_inner", "fitch_tree", ... instead of mangled names) and, while
// the engine's profiler is on, an event pair on the engine's stream
struct ProfScope {
  phylo_engine *e;
  int idx = -1;
  ProfScope(phylo_engine *e_, int cls) : e(e_) {
    nvtxRangePushA(kClassNames[cls]);
    if (!e->prof_on) return;
    idx = (int)e->prof_pending.size();
    while (e->prof_pool.size() < (size_t)(2 * idx + 2)) {
      cudaEvent_t ev;
      if (cudaEventCreate(&ev) != cudaSuccess) { idx = -1; return; }
      e->prof_pool.push_back(ev);
    }
    e->prof_pending.push_back(cls);
    cudaEventRecord(e->prof_pool[2 * idx], e->stream);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(e->prof_pool[2 * idx + 1], e->stream);
    nvtxRangePop();
  }
};

// fold finished event pairs into the per-class totals (call after a stream sync)
static void prof_resolve(phylo_engine *e) {
  if (e->prof_pending.empty()) return;
  cudaStreamSynchronize(e->stream);
  for (size_t i = 0; i < e->prof_pending.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->prof_pool[2 * i], e->prof_pool[2 * i + 1]) == cudaSuccess) {
      e->prof_ms[e->prof_pending[i]] += ms;
      e->prof_n[e->prof_pending[i]] += 1;
    }
  }
  e->prof_pending.clear();
}

// per-call variant: event pairs are only folded once a few thousand are pending (or when the
// totals are read), so small evaluations do not pay an event synchronisation each
static void prof_resolve_lazy(phylo_engine *e) {
  if (e->prof_pending.size() >= 4096) prof_resolve(e);
}

template <typename T>
static void dfree(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

static int grid_for(int64_t work_items, int per_block, int max_blocks) {
  int64_t g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

static int lk_finish_reduce(phylo_engine *e, double *slot);

// ------------------------------------------------------------------------ engine ----
extern "C" int phylo_engine_create(int device, phylo_engine **out) {
  phylo_engine *e = nullptr;
  if (!out) return fail(nullptr, PHYLO_ERR_ARG, "phylo_engine_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t st = cudaGetDeviceCount(&ndev);
  if (st != cudaSuccess || ndev == 0)
    return fail(nullptr, PHYLO_ERR_CUDA,
                "phylo_engine_create: no usable CUDA device (%s); this engine has no CPU fallback",
                st != cudaSuccess ? cudaGetErrorString(st) : "device count is 0");
  if (device < 0 || device >= ndev)
    return fail(nullptr, PHYLO_ERR_ARG, "phylo_engine_create: device %d out of range [0,%d)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(nullptr, PHYLO_ERR_UNSUPPORTED,
                "phylo_engine_create: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  e = new phylo_engine();
  e->device = device;
  if (const char *v = std::getenv("PHYLO_TT_TABLE")) e->opt_tt_table = std::atoi(v);
  e->sm_count = prop.multiProcessorCount;
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      e->encodeTiled = (phylo_engine::EncodeTiledFn)fn;
    cudaGetLastError();
  }
  if (cudaMallocHost(&e->hScalar, sizeof(double) * HS_DOUBLES) != cudaSuccess) {
    delete e;
    return fail(nullptr, PHYLO_ERR_CUDA, "phylo_engine_create: cudaMallocHost failed");
  }
  *out = e;
  return PHYLO_OK;
}

static void lk_free_data(phylo_engine *e) {
  for (auto &n : e->nodes) {
    dfree(n.clv);
    dfree(n.scale);
  }
  e->nodes.clear();
  dfree(e->dTips);
  dfree(e->dTips4);
  dfree(e->dNodeClv);
  dfree(e->dNodeSc);
  dfree(e->dTmaps);
  dfree(e->dSum); dfree(e->dSumSc); dfree(e->dEdgePart);
  e->edge_ready = false;
  e->nodeTabDirty = true;
  e->tmapDirty = true;
  dfree(e->dInv);
  dfree(e->dWeights);
  dfree(e->dSite);
  dfree(e->dWSite);
  dfree(e->dPart);
  dfree(e->dGroups);
  dfree(e->dPart2);
  e->T = 0; e->N = 0; e->cap = 0; e->lk_evaluated = false;
  e->lkOwned.clear(); e->lkFree.clear(); e->lkNext = 0; e->cap0 = 0;
  ++e->lkGen;  // slots handed out so far are gone: late releases of them are ignored
}

static void sankoff_free_data(phylo_engine *e) {
  for (auto &v : e->skVec) dfree(v);
  e->skVec.clear(); e->skValid.clear();
  dfree(e->dSkTips); dfree(e->dSkW); dfree(e->dSkTab);
  e->skTabDirty = true;
  e->skT = 0; e->skN = 0; e->skCap = 0;
}

static void fitch_free_data(phylo_engine *e) {
  if (e->dTipSlab)  // the tips' plane buffers are rows of one allocation
    for (int t = 0; t < e->fT && t < (int)e->fPre.size(); ++t) e->fPre[t] = nullptr;
  dfree(e->dTipSlab);
  for (auto &p : e->fPre) dfree(p);
  for (auto &p : e->fFin) dfree(p);
  e->fPre.clear();
  e->fFin.clear();
  dfree(e->dPreTab);
  dfree(e->dFinTab);
  dfree(e->dFW);
  e->fT = 0; e->fN = 0; e->fcap = 0;
  e->fValid.clear(); e->fFinValid.clear(); e->fOwned.clear(); e->fFreeList.clear(); e->fNext = 0; e->fcap0 = 0;
  ++e->fGen;
}

extern "C" void phylo_engine_destroy(phylo_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  lk_free_data(e);
  fitch_free_data(e);
  sankoff_free_data(e);
  dfree(e->dSkM); dfree(e->dSkTotal);
  dfree(e->dU); dfree(e->dLam); dfree(e->dUi); dfree(e->dPi); dfree(e->dRates); dfree(e->dProbs);
  dfree(e->dP); dfree(e->dFrag); dfree(e->dT);
  for (void *p : e->xOpened) cudaIpcCloseMemHandle(p);
  dfree(e->xMailbox);
  dfree(e->dTT); dfree(e->dTTsc); dfree(e->dSymTab); dfree(e->dCost); dfree(e->dSched); dfree(e->dStage); dfree(e->dProg); dfree(e->dRaw); dfree(e->dBad); dfree(e->dSpill); dfree(e->dAcc); dfree(e->dStamps); dfree(e->dTreeDone); dfree(e->dTcm); dfree(e->dUL); dfree(e->dUR); dfree(e->dEdgeOut); dfree(e->dEdgeT); dfree(e->dEdgeArena);
  if (e->hProg) cudaFreeHost(e->hProg);
  for (auto ev : e->prof_pool) cudaEventDestroy(ev);
  for (auto ev : e->slabEvents) cudaEventDestroy(ev);
  if (e->copyStream) cudaStreamDestroy(e->copyStream);
  if (e->copyStream2) cudaStreamDestroy(e->copyStream2);
  if (e->auxStream) cudaStreamDestroy(e->auxStream);
  if (e->auxDone) cudaEventDestroy(e->auxDone);
  if (e->hT) cudaFreeHost(e->hT);
  if (e->hScalar) cudaFreeHost(e->hScalar);
  if (e->hEdgeRes) cudaFreeHost(e->hEdgeRes);
  if (e->hSched) cudaFreeHost(e->hSched);
  if (e->hCost) cudaFreeHost(e->hCost);
  delete e;
}

extern "C" const char *phylo_last_error(const phylo_engine *e) {
  return e ? e->err.c_str() : g_create_error.c_str();
}

extern "C" int phylo_engine_set_stream(phylo_engine *e, void *s) {
  if (!e) return PHYLO_ERR_ARG;
  e->stream = (cudaStream_t)s;
  return PHYLO_OK;
}

extern "C" int phylo_engine_sync(phylo_engine *e) {
  if (!e) return PHYLO_ERR_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  return PHYLO_OK;
}

extern "C" uint64_t phylo_engine_launch_count(const phylo_engine *e) { return e ? e->launches : 0; }

extern "C" int phylo_engine_set_symbol_table(phylo_engine *e, const uint64_t *table256) {
  if (!e) return PHYLO_ERR_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  if (!table256) {
    dfree(e->dSymTab);
    return PHYLO_OK;
  }
  if (!e->dSymTab) CK(cudaMalloc(&e->dSymTab, sizeof(uint64_t) * 256));
  CK(cudaMemcpy(e->dSymTab, table256, sizeof(uint64_t) * 256, cudaMemcpyHostToDevice));
  e->symtab_fits_byte = true;
  for (int i = 0; i < 256; ++i)
    if (table256[i] > 0xff) e->symtab_fits_byte = false;
  return PHYLO_OK;
}

extern "C" int phylo_engine_profile(phylo_engine *e, int enable) {
  if (!e) return PHYLO_ERR_ARG;
  CK(cudaSetDevice(e->device));
  prof_resolve(e);
  e->prof_on = enable != 0;
  return PHYLO_OK;
}
extern "C" int phylo_engine_profile_reset(phylo_engine *e) {
  if (!e) return PHYLO_ERR_ARG;
  CK(cudaSetDevice(e->device));
  prof_resolve(e);
  for (int c = 0; c < KC_COUNT; ++c) { e->prof_ms[c] = 0; e->prof_n[c] = 0; }
  return PHYLO_OK;
}
extern "C" int phylo_engine_profile_get(phylo_engine *e, int kernel_class, double *ms_total, uint64_t *launches) {
  if (!e) return PHYLO_ERR_ARG;
  if (kernel_class < 0 || kernel_class >= KC_COUNT) return fail(e, PHYLO_ERR_ARG, "profile_get: class %d out of range", kernel_class);
  CK(cudaSetDevice(e->device));
  prof_resolve(e);
  if (ms_total) *ms_total = e->prof_ms[kernel_class];
  if (launches) *launches = e->prof_n[kernel_class];
  return PHYLO_OK;
}
extern "C" int phylo_kernel_class_count(void) { return KC_COUNT; }
extern "C" const char *phylo_kernel_class_name(int kernel_class) {
  return (kernel_class >= 0 && kernel_class < KC_COUNT) ? kClassNames[kernel_class] : nullptr;
}

extern "C" int phylo_host_alloc(void **out, uint64_t bytes) {
  if (!out) return PHYLO_ERR_ARG;
  return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? PHYLO_OK : PHYLO_ERR_CUDA;
}
extern "C" int phylo_host_free(void *p) {
  return cudaFreeHost(p) == cudaSuccess ? PHYLO_OK : PHYLO_ERR_CUDA;
}

// ---------------------------------------------------------------- P(t) plumbing ----
static int ensure_pt_capacity(phylo_engine *e, size_t branches, int S, int K) {
  if (branches + 1 <= e->capP && e->dP) return PHYLO_OK;
  size_t nb = std::max(branches + 1, (size_t)64);  // +1: the fused kernel copies matrix sets in pairs
  CK(cudaStreamSynchronize(e->stream));
  dfree(e->dP);
  dfree(e->dT);
  if (e->hT) { cudaFreeHost(e->hT); e->hT = nullptr; }
  e->capP = 0;
  e->ptLastT.clear();
  CK(cudaMalloc(&e->dP, sizeof(double) * nb * K * S * S));
  CK(cudaMalloc(&e->dT, sizeof(double) * nb));
  CK(cudaMallocHost(&e->hT, sizeof(double) * nb));
  e->capP = nb;
  return PHYLO_OK;
}

// launches pt_build for branches [0, nb) whose lengths are already in e->hT
static int build_pt(phylo_engine *e, int nb, int interleave = 0) {
  // the same branch lengths under the same model as the last build (the same tree scored again: new weights,
  // a re-upload of the tips, a benchmark loop): e->dP already holds these matrices -- no copy, no launch
  if ((int)e->ptLastT.size() == nb && e->ptLastInterleave == interleave &&
      std::memcmp(e->ptLastT.data(), e->hT, sizeof(double) * nb) == 0)
    return PHYLO_OK;
  e->ptLastT.assign(e->hT, e->hT + nb);
  e->ptLastInterleave = interleave;
  CK(cudaMemcpyAsync(e->dT, e->hT, sizeof(double) * nb, cudaMemcpyHostToDevice, e->stream));
  const int threads = std::min(256, std::max(32, ((e->S * e->S + 31) / 32) * 32));
  ProfScope prof(e, KC_PT_BUILD);
  pt_build_kernel<<<nb * e->K, threads, sizeof(double) * e->S, e->stream>>>(
      e->dU, e->dLam, e->sym ? nullptr : e->dUi, e->dRates, e->dT, e->S, e->K, e->dP, interleave);
  LAUNCH_CHECK();
  return PHYLO_OK;
}

static int compose_common(phylo_engine *e, const double *U, const double *D, const double *Ui,
                          double t, int n, double *P_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (!U || !D || !P_out || n < 1 || n > 64) return fail(e, PHYLO_ERR_ARG, "compose: bad arguments (n=%d)", n);
  CK(cudaSetDevice(e->device));
  double *dbuf = nullptr;
  const size_t nn = (size_t)n * n;
  // layout: U | lam | Ui | rate(=1) | t | P
  CK(cudaMalloc(&dbuf, sizeof(double) * (3 * nn + n + 2)));
  std::vector<double> h(2 * nn + n + 2);
  std::memcpy(h.data(), U, sizeof(double) * nn);
  for (int i = 0; i < n; ++i) h[nn + i] = D[(size_t)i * n + i];
  if (Ui) std::memcpy(h.data() + nn + n, Ui, sizeof(double) * nn);
  h[2 * nn + n] = 1.0;
  h[2 * nn + n + 1] = t;
  cudaError_t st = cudaMemcpyAsync(dbuf, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess) {
    const int threads = std::min(256, std::max(32, ((n * n + 31) / 32) * 32));
    pt_build_kernel<<<1, threads, sizeof(double) * n, e->stream>>>(
        dbuf, dbuf + nn, Ui ? dbuf + nn + n : nullptr, dbuf + 2 * nn + n, dbuf + 2 * nn + n + 1, n, 1,
        dbuf + 2 * nn + n + 2, 0);
    ++e->launches;
    st = cudaGetLastError();
  }
  if (st == cudaSuccess)
    st = cudaMemcpyAsync(P_out, dbuf + 2 * nn + n + 2, sizeof(double) * nn, cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(dbuf);
  if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "compose: %s", cudaGetErrorString(st));
  return PHYLO_OK;
}

extern "C" int phylo_compose_sym(phylo_engine *e, const double *U, const double *D, double t, int n,
                                 double *P_out) {
  return compose_common(e, U, D, nullptr, t, n, P_out);
}
extern "C" int phylo_compose_gtr(phylo_engine *e, const double *U, const double *D, const double *Ui,
                                 double t, int n, double *P_out) {
  if (!Ui) return e ? fail(e, PHYLO_ERR_ARG, "compose_gtr: Ui is NULL") : PHYLO_ERR_ARG;
  return compose_common(e, U, D, Ui, t, n, P_out);
}

// ------------------------------------------------------------------- likelihood ----
extern "C" int phylo_lk_set_model(phylo_engine *e, int S, int K, const double *U, const double *D,
                                  const double *Ui, const double *priors, const double *rates,
                                  const double *probs, double pinvar) {
  if (!e) return PHYLO_ERR_ARG;
  if (S < 2 || S > 64 || K < 1 || K > 64 || !U || !D || !priors || !rates || !probs)
    return fail(e, PHYLO_ERR_ARG, "lk_set_model: bad arguments (S=%d K=%d)", S, K);
  if (pinvar >= 1.0) return fail(e, PHYLO_ERR_ARG, "lk_set_model: pinvar %g >= 1", pinvar);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  if (e->T > 0 && S != e->S) lk_free_data(e);  // another alphabet: the loaded tips are void
  if (e->T > 0 && K != e->K) {  // CLV shape changes: drop interior buffers
    for (auto &n : e->nodes) { dfree(n.clv); dfree(n.scale); n.valid = false; }
  }
  for (auto &n : e->nodes) n.valid = false;  // CLVs of the previous model are stale
  dfree(e->dU); dfree(e->dLam); dfree(e->dUi); dfree(e->dPi); dfree(e->dRates); dfree(e->dProbs);
  dfree(e->dP); e->capP = 0;
  dfree(e->dFrag); e->capFrag = 0;
  const size_t ss = (size_t)S * S;
  std::vector<double> lam(S);
  for (int i = 0; i < S; ++i) lam[i] = D[(size_t)i * S + i];
  CK(cudaMalloc(&e->dU, sizeof(double) * ss));
  CK(cudaMalloc(&e->dLam, sizeof(double) * S));
  CK(cudaMalloc(&e->dPi, sizeof(double) * S));
  CK(cudaMalloc(&e->dRates, sizeof(double) * K));
  CK(cudaMalloc(&e->dProbs, sizeof(double) * K));
  CK(cudaMemcpy(e->dU, U, sizeof(double) * ss, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->dLam, lam.data(), sizeof(double) * S, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->dPi, priors, sizeof(double) * S, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->dRates, rates, sizeof(double) * K, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->dProbs, probs, sizeof(double) * K, cudaMemcpyHostToDevice));
  if (Ui) {
    CK(cudaMalloc(&e->dUi, sizeof(double) * ss));
    CK(cudaMemcpy(e->dUi, Ui, sizeof(double) * ss, cudaMemcpyHostToDevice));
  }
  {
    // P = UL diag(e) UR: GTR UL = U, UR = Ui (lib/mlmodel.c:325-342); symmetric UL = U^T, UR = U (:280-302)
    std::vector<double> ul(ss), ur(ss);
    for (int i = 0; i < S; ++i)
      for (int m = 0; m < S; ++m) {
        ul[(size_t)i * S + m] = Ui ? U[(size_t)i * S + m] : U[(size_t)m * S + i];
        ur[(size_t)m * S + i] = Ui ? Ui[(size_t)m * S + i] : U[(size_t)m * S + i];
      }
    // The edge sum table c_km = (sum_i pi_i a_ki UL[i][m]) (sum_j UR[m][j] b_kj) is a pruning update
    // with the "transition matrices" M1[m][i] = pi_i UL[i][m] and M2[m][j] = UR[m][j] for every
    // rate class: phylo_lk_edge_prepare runs the ordinary pruning kernels on them.
    std::vector<double> m1((size_t)K * ss), m2((size_t)K * ss);
    for (int k = 0; k < K; ++k)
      for (int m = 0; m < S; ++m)
        for (int i = 0; i < S; ++i) {
          m1[(size_t)k * ss + (size_t)m * S + i] = priors[i] * ul[(size_t)i * S + m];
          m2[(size_t)k * ss + (size_t)m * S + i] = ur[(size_t)m * S + i];
        }
    dfree(e->dUL); dfree(e->dUR);
    CK(cudaMalloc(&e->dUL, sizeof(double) * K * ss));
    CK(cudaMalloc(&e->dUR, sizeof(double) * K * ss));
    CK(cudaMemcpy(e->dUL, m1.data(), sizeof(double) * K * ss, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->dUR, m2.data(), sizeof(double) * K * ss, cudaMemcpyHostToDevice));
  }
  e->hV.assign(ss, 0.0); e->hVinv.assign(ss, 0.0);
  for (int i = 0; i < S; ++i)
    for (int m = 0; m < S; ++m) {
      e->hV[(size_t)i * S + m] = Ui ? U[(size_t)i * S + m] : U[(size_t)m * S + i];
      e->hVinv[(size_t)m * S + i] = Ui ? Ui[(size_t)m * S + i] : U[(size_t)m * S + i];
    }
  e->hLam = lam;
  e->hPi.assign(priors, priors + S);
  e->hRates.assign(rates, rates + K);
  e->hProbs.assign(probs, probs + K);
  dfree(e->dSum); dfree(e->dSumSc);
  e->edge_ready = false;
  e->S = S; e->K = K; e->sym = (Ui == nullptr); e->pinvar = pinvar < 0 ? -1.0 : pinvar;
  e->has_model = true;
  return PHYLO_OK;
}

static int dev_mask_bytes(int S) { return S <= 8 ? 1 : (S <= 32 ? 4 : 8); }

template <typename InT>
static int launch_tips_prepare(phylo_engine *e, const void *raw, unsigned long long *dBad, int64_t p_lo, int64_t p_hi,
                               cudaStream_t cs) {
  const int g = grid_for(p_hi - p_lo, 256, e->sm_count * 8);
  const uint64_t *lut = sizeof(InT) == 1 ? e->dSymTab : nullptr;  // symbols are bytes (checked in lk_prepare_shape)
  ProfScope prof(e, KC_TIPS_PREPARE);
  switch (e->mask_dev_bytes) {
    case 1:
      tips_prepare_kernel<InT, uint8_t><<<g, 256, 0, cs>>>((const InT *)raw, (uint8_t *)e->dTips, e->tipStride,
                                                                 (uint8_t *)e->dInv, e->T, e->N, e->S, dBad, p_lo, p_hi, lut);
      break;
    case 4:
      tips_prepare_kernel<InT, uint32_t><<<g, 256, 0, cs>>>((const InT *)raw, (uint32_t *)e->dTips, e->tipStride,
                                                                  (uint32_t *)e->dInv, e->T, e->N, e->S, dBad, p_lo, p_hi, lut);
      break;
    default:
      tips_prepare_kernel<InT, uint64_t><<<g, 256, 0, cs>>>((const InT *)raw, (uint64_t *)e->dTips, e->tipStride,
                                                                  (uint64_t *)e->dInv, e->T, e->N, e->S, dBad, p_lo, p_hi, lut);
  }
  LAUNCH_CHECK();
  return PHYLO_OK;
}

// validation + (re)allocation for an alignment of this shape; keeps every device allocation
// when the shape is unchanged (a tree-search loop re-uploads often)
static int lk_prepare_shape(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                            const double *weights, int capacity, const char *who) {
  if (!e->has_model) return fail(e, PHYLO_ERR_STATE, "%s: call phylo_lk_set_model first", who);
  if (T < 2 || N < 1 || !masks || capacity < T ||
      !(mask_bytes == 0 || mask_bytes == 1 || mask_bytes == 2 || mask_bytes == 4 || mask_bytes == 8))
    return fail(e, PHYLO_ERR_ARG, "%s: bad arguments (T=%d N=%lld mask_bytes=%d capacity=%d)", who, T,
                (long long)N, mask_bytes, capacity);
  if (mask_bytes == 0 && (e->S != 4 || e->dSymTab))
    return fail(e, PHYLO_ERR_ARG, "%s: packed 4-bit masks (mask_bytes 0) need a 4-state model and no symbol table", who);
  if (e->dSymTab && mask_bytes != 1)
    return fail(e, PHYLO_ERR_ARG, "%s: a symbol table is set, the alignment must be 1 byte per cell (got %d)", who, mask_bytes);
  if (!e->dSymTab && mask_bytes > 0 && mask_bytes * 8 < e->S)
    return fail(e, PHYLO_ERR_ARG, "%s: %d-bit masks cannot hold %d states", who, mask_bytes * 8, e->S);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  const bool reuse = e->dTips && e->T == T && e->N == N && e->cap0 == capacity &&
                     e->mask_dev_bytes == dev_mask_bytes(e->S) && (weights != nullptr) == (e->dWeights != nullptr);
  // bytes of the raw upload in its staging buffer (packed rows are padded to 16 bytes)
  const size_t raw_bytes = mask_bytes ? (size_t)T * N * mask_bytes : (size_t)T * ((((size_t)N + 1) / 2 + 15) & ~(size_t)15);
  if (reuse) {
    for (auto &n : e->nodes) n.valid = false;
    e->lk_evaluated = false;
    e->edge_ready = false;
  } else {
    lk_free_data(e);
    e->T = T; e->N = N; e->cap = capacity; e->cap0 = capacity;
    e->mask_dev_bytes = dev_mask_bytes(e->S);
    e->nodes.assign(capacity, LkNode());
    e->lkOwned.assign(capacity, 0);
    e->lkFree.clear();
    e->lkNext = T;
    e->tipStride = ((N + kLnlBlock - 1) / kLnlBlock) * kLnlBlock;
    CK(cudaMalloc(&e->dTips, (size_t)T * e->tipStride * e->mask_dev_bytes));
    // padding cells read as "all states" (never used in a sum, but keeps them harmless)
    CK(cudaMemsetAsync(e->dTips, 0xff, (size_t)T * e->tipStride * e->mask_dev_bytes, e->stream));
    if (e->S == 4) CK(cudaMalloc(&e->dTips4, (size_t)T * e->tipStride / 2));
    CK(cudaMalloc(&e->dNodeClv, sizeof(double *) * capacity));
    CK(cudaMalloc(&e->dNodeSc, sizeof(int32_t *) * capacity));
    CK(cudaMalloc(&e->dTmaps, sizeof(CUtensorMap) * capacity));
    e->nodeTabDirty = true;
    e->tmapDirty = true;
    CK(cudaMalloc(&e->dInv, (size_t)N * e->mask_dev_bytes));
    if (weights) CK(cudaMalloc(&e->dWeights, sizeof(double) * N));
    e->nPart = (N + kLnlBlock - 1) / kLnlBlock;
    CK(cudaMalloc(&e->dPart, sizeof(double) * e->nPart));
    CK(cudaMalloc(&e->dGroups, sizeof(double) * e->nPart * 32));
    CK(cudaMalloc(&e->dPart2, sizeof(double) * ((e->nPart + kLnlBlock - 1) / kLnlBlock + 1) * 2));
    CK(cudaMalloc(&e->dSite, sizeof(double) * N));
  }
  // the raw upload lands in a staging buffer; tips_prepare writes the padded device rows
  if (raw_bytes > e->capRaw) {
    dfree(e->dRaw);
    e->capRaw = 0;
    CK(cudaMalloc(&e->dRaw, raw_bytes));
    e->capRaw = raw_bytes;
  }
  e->tips8_valid = false;
  if (!e->dBad) CK(cudaMalloc(&e->dBad, sizeof(unsigned long long)));
  CK(cudaMemsetAsync(e->dBad, 0, sizeof(unsigned long long), e->stream));
  if (weights) CK(cudaMemcpyAsync(e->dWeights, weights, sizeof(double) * N, cudaMemcpyHostToDevice, e->stream));
  return PHYLO_OK;
}

// patterns [p_lo, p_hi) of the host alignment -> device: H2D on `copy` (NULL: the engine's
// own stream), then mask conversion / validation / nibble packing on the engine's stream
static int lk_upload_slab(phylo_engine *e, const void *masks, int mask_bytes, int64_t p_lo, int64_t p_hi,
                          cudaStream_t copy, cudaEvent_t ready, cudaStream_t cs) {
  if (mask_bytes == 0) {
    // packed nibbles: bytes [p_lo / 2, ceil(p_hi / 2)) of every row, then straight into the group-major tiles
    const size_t nb = ((size_t)e->N + 1) / 2, dpitch = (nb + 15) & ~(size_t)15, host_pitch = e->hostPitch ? e->hostPitch : nb;
    const size_t b_lo = (size_t)p_lo / 2, b_hi = std::min(nb, ((size_t)p_hi + 1) / 2);
    CK(cudaMemcpy2DAsync((char *)e->dRaw + b_lo, dpitch, (const char *)masks + b_lo, host_pitch, b_hi - b_lo, e->T,
                         cudaMemcpyHostToDevice, copy ? copy : cs));
    if (copy) {
      CK(cudaEventRecord(ready, copy));
      CK(cudaStreamWaitEvent(cs, ready, 0));
    }
    const int64_t g_lo = p_lo / 32, g_hi = (p_hi >= e->N) ? e->tipStride / 32 : p_hi / 32;
    ProfScope prof(e, KC_TIPS_PREPARE);
    dim3 grid((unsigned)((g_hi - g_lo + 31) / 32), (unsigned)((e->T + 31) / 32));
    tips_nibbles_to_groups_kernel<<<grid, 256, 0, cs>>>((const uint8_t *)e->dRaw, dpitch, e->dTips4, e->T, e->N, g_lo, g_hi);
    LAUNCH_CHECK();
    tips_groups_check_kernel<<<grid_for((g_hi - g_lo) * 16, 256, e->sm_count * 8), 256, 0, cs>>>(
        e->dTips4, e->T, e->N, g_lo, g_hi, (uint8_t *)e->dInv, e->dBad);
    LAUNCH_CHECK();
    return PHYLO_OK;
  }
  const size_t pitch = (size_t)e->N * mask_bytes;
  const size_t host_pitch = e->hostPitch ? e->hostPitch : pitch;
  CK(cudaMemcpy2DAsync((char *)e->dRaw + (size_t)p_lo * mask_bytes, pitch, (const char *)masks + (size_t)p_lo * mask_bytes,
                       host_pitch, (size_t)(p_hi - p_lo) * mask_bytes, e->T, cudaMemcpyHostToDevice, copy ? copy : cs));
  if (copy) {
    CK(cudaEventRecord(ready, copy));
    CK(cudaStreamWaitEvent(cs, ready, 0));
  }
  int rc;
  switch (mask_bytes) {
    case 1: rc = launch_tips_prepare<uint8_t>(e, e->dRaw, e->dBad, p_lo, p_hi, cs); break;
    case 2: rc = launch_tips_prepare<uint16_t>(e, e->dRaw, e->dBad, p_lo, p_hi, cs); break;
    case 4: rc = launch_tips_prepare<uint32_t>(e, e->dRaw, e->dBad, p_lo, p_hi, cs); break;
    default: rc = launch_tips_prepare<uint64_t>(e, e->dRaw, e->dBad, p_lo, p_hi, cs);
  }
  if (rc != PHYLO_OK) return rc;
  if (e->dTips4) {
    ProfScope prof(e, KC_TIPS_PREPARE);
    const int64_t b_lo = p_lo / 2, b_hi = (p_hi >= e->N) ? e->tipStride / 2 : p_hi / 2;
    tips_pack4_kernel<<<grid_for((int64_t)e->T * (b_hi - b_lo), 256, e->sm_count * 8), 256, 0, cs>>>(
        (const uint8_t *)e->dTips, e->dTips4, e->T, e->N, e->tipStride, b_lo, b_hi);
    LAUNCH_CHECK();
  }
  if (p_hi >= e->N) e->tips8_valid = true;  // the last slab of a one-byte-per-cell upload
  return PHYLO_OK;
}

// one byte per cell tip rows for the per-node kernels after a packed upload (lazy)
static int lk_ensure_tips8(phylo_engine *e) {
  if (e->tips8_valid) return PHYLO_OK;
  ProfScope prof(e, KC_TIPS_PREPARE);
  tips_groups_to_bytes_kernel<<<grid_for((int64_t)e->T * e->tipStride / 2, 256, e->sm_count * 8), 256, 0, e->stream>>>(
      e->dTips4, (uint8_t *)e->dTips, e->T, e->tipStride);
  LAUNCH_CHECK();
  e->tips8_valid = true;
  return PHYLO_OK;
}

// reads back the invalid-mask counter (call after a stream sync)
static int lk_check_bad(phylo_engine *e, const char *who) {
  CK(cudaMemcpyAsync(e->hScalar + HS_BAD, e->dBad, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const unsigned long long bad = *(unsigned long long *)(e->hScalar + HS_BAD);
  if (bad) {
    lk_free_data(e);
    return fail(e, PHYLO_ERR_DATA, "%s: %llu tip cells have none of the %d state bits set", who, bad, e->S);
  }
  return PHYLO_OK;
}

extern "C" int phylo_lk_set_tips_pitched(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                                         uint64_t host_pitch_bytes, const double *weights, int capacity) {
  if (!e) return PHYLO_ERR_ARG;
  if (host_pitch_bytes != 0 && N > 0 &&
      host_pitch_bytes < (mask_bytes ? (uint64_t)N * (uint64_t)mask_bytes : ((uint64_t)N + 1) / 2))
    return fail(e, PHYLO_ERR_ARG, "lk_set_tips: host pitch %llu is shorter than a row of %lld masks",
                (unsigned long long)host_pitch_bytes, (long long)N);
  int rc;
  if ((rc = lk_prepare_shape(e, T, N, masks, mask_bytes, weights, capacity, "lk_set_tips")) != PHYLO_OK) return rc;
  e->hostPitch = (size_t)host_pitch_bytes;
  rc = lk_upload_slab(e, masks, mask_bytes, 0, N, nullptr, nullptr, e->stream);
  e->hostPitch = 0;
  if (rc != PHYLO_OK) { lk_free_data(e); return rc; }
  return lk_check_bad(e, "lk_set_tips");
}

extern "C" int phylo_lk_set_tips(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                                 const double *weights, int capacity) {
  return phylo_lk_set_tips_pitched(e, T, N, masks, mask_bytes, 0, weights, capacity);
}

static int lk_ensure_node(phylo_engine *e, int slot) {
  LkNode &n = e->nodes[slot];
  if (n.clv) return PHYLO_OK;
  CK(cudaMalloc(&n.clv, sizeof(double) * (size_t)e->N * e->K * e->S));
  CK(cudaMalloc(&n.scale, sizeof(int32_t) * (size_t)e->N));
  e->nodeTabDirty = true;
  e->tmapDirty = true;
  return PHYLO_OK;
}

// ---- node-slot lifetime (the drop-in's answer to the reference's custom blocks with a finalizer,
// lib/bitvector/bv.c:183-189,229-244: a node value owns native memory and gives it back when the GC
// collects it). Slots T..cap-1 name interior nodes; alloc hands out a released slot first (its CLV
// buffer is still attached), then a never-used one, and only then grows the slot table.
static int lk_grow_slots(phylo_engine *e, int new_cap) {
  CK(cudaStreamSynchronize(e->stream));
  double **nclv = nullptr;
  int32_t **nsc = nullptr;
  CUtensorMap *ntm = nullptr;
  if (cudaMalloc(&nclv, sizeof(double *) * new_cap) != cudaSuccess || cudaMalloc(&nsc, sizeof(int32_t *) * new_cap) != cudaSuccess ||
      cudaMalloc(&ntm, sizeof(CUtensorMap) * new_cap) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(nclv); cudaFree(nsc); cudaFree(ntm);
    return fail(e, PHYLO_ERR_CUDA, "lk_node_alloc: cannot grow the slot tables to %d slots", new_cap);
  }
  dfree(e->dNodeClv); dfree(e->dNodeSc); dfree(e->dTmaps);
  e->dNodeClv = nclv; e->dNodeSc = nsc; e->dTmaps = ntm;
  e->nodes.resize(new_cap);
  e->lkOwned.resize(new_cap, 0);
  e->cap = new_cap;
  e->nodeTabDirty = true;
  e->tmapDirty = true;
  return PHYLO_OK;
}

extern "C" int phylo_lk_node_alloc(phylo_engine *e, int *slot_out, uint64_t *generation_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (!slot_out) return fail(e, PHYLO_ERR_ARG, "lk_node_alloc: slot_out is NULL");
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_node_alloc: no tips loaded");
  CK(cudaSetDevice(e->device));
  int slot = -1;
  if (!e->lkFree.empty()) {
    slot = e->lkFree.back();
    e->lkFree.pop_back();
  } else {
    while (e->lkNext < e->cap && (e->lkOwned[e->lkNext] || e->nodes[e->lkNext].valid)) ++e->lkNext;  // slots a schedule named directly
    if (e->lkNext >= e->cap) {
      const int rc = lk_grow_slots(e, std::max(e->cap * 2, e->cap + 16));
      if (rc != PHYLO_OK) return rc;
    }
    slot = e->lkNext++;
  }
  e->lkOwned[slot] = 1;
  e->nodes[slot].valid = false;
  *slot_out = slot;
  if (generation_out) *generation_out = e->lkGen;
  return PHYLO_OK;
}

extern "C" int phylo_lk_node_release(phylo_engine *e, int slot, uint64_t generation) {
  if (!e) return PHYLO_ERR_ARG;
  if (generation != e->lkGen) return PHYLO_OK;  // the alignment this slot belonged to is gone already
  if (slot < e->T || slot >= e->cap || !e->lkOwned[slot])
    return fail(e, PHYLO_ERR_ARG, "lk_node_release: slot %d was not handed out by phylo_lk_node_alloc", slot);
  e->lkOwned[slot] = 0;
  e->nodes[slot].valid = false;
  if (e->edge_a == slot || e->edge_b == slot) e->edge_ready = false;
  e->lkFree.push_back(slot);
  return PHYLO_OK;
}

extern "C" int phylo_lk_node_stats(phylo_engine *e, int *capacity, int *in_use, int *with_buffers) {
  if (!e) return PHYLO_ERR_ARG;
  int used = 0, buf = 0;
  for (int s = e->T; s < e->cap; ++s) { used += e->lkOwned[s] != 0; buf += e->nodes[s].clv != nullptr; }
  if (capacity) *capacity = e->cap - e->T;
  if (in_use) *in_use = used;
  if (with_buffers) *with_buffers = buf;
  return PHYLO_OK;
}

extern "C" int phylo_lk_shape(phylo_engine *e, int *n_taxa, int64_t *n_patterns, int *n_slots) {
  if (!e) return PHYLO_ERR_ARG;
  if (n_taxa) *n_taxa = e->T;
  if (n_patterns) *n_patterns = e->N;
  if (n_slots) *n_slots = e->cap;
  return PHYLO_OK;
}

struct Operand {
  const void *src;
  const int32_t *scale;
  bool tip;
};

static int lk_operand(phylo_engine *e, int slot, Operand *o, const char *who) {
  if (slot < 0 || slot >= e->cap) return fail(e, PHYLO_ERR_ARG, "%s: node slot %d out of range [0,%d)", who, slot, e->cap);
  if (slot < e->T) {
    if (!e->tips8_valid) {
      const int rc = lk_ensure_tips8(e);
      if (rc != PHYLO_OK) return rc;
    }
    o->src = (const char *)e->dTips + (size_t)slot * e->tipStride * e->mask_dev_bytes;
    o->scale = nullptr;
    o->tip = true;
    return PHYLO_OK;
  }
  if (!e->nodes[slot].valid) return fail(e, PHYLO_ERR_STATE, "%s: node slot %d has no CLV yet", who, slot);
  o->src = e->nodes[slot].clv;
  o->scale = e->nodes[slot].scale;
  o->tip = false;
  return PHYLO_OK;
}

// persistent-style grid: as many CTAs as fit on the chip at once (148 SMs x occupancy),
// each grid-striding over the items
template <typename Kern>
static int resident_grid(phylo_engine *e, Kern kern, int threads, size_t smem, int64_t needed) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  return (int)std::max<int64_t>(1, std::min<int64_t>(needed, (int64_t)e->sm_count * occ));
}

template <int K>
static void launch_prune4(phylo_engine *e, const double *Pl, const double *Pr, const Operand &l,
                          const Operand &r, double *out, int32_t *osc) {
#define P4(LT, RT, U)                                                                           \
  do {                                                                                          \
    auto kern = prune4_kernel<K, LT, RT, U>;                                                    \
    const int g = resident_grid(e, kern, 256, 0, (e->N * K + 256 * U - 1) / (256 * U));         \
    kern<<<g, 256, 0, e->stream>>>(Pl, Pr, l.src, l.scale, r.src, r.scale, out, osc, e->N);     \
  } while (0)
  if (l.tip && r.tip) {
    constexpr int U = 4;
    const size_t smem = 256 * K * sizeof(d4) + 256 * sizeof(int);
    auto kern = prune4_tt_kernel<K, U>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int g = resident_grid(e, kern, 256, smem, (e->N * K + 256 * U - 1) / (256 * U));
    kern<<<g, 256, smem, e->stream>>>(Pl, Pr, (const uint8_t *)l.src, (const uint8_t *)r.src, out, osc, e->N);
  } else if (l.tip) P4(true, false, 4);
  else if (r.tip) P4(false, true, 4);
  else P4(false, false, 2);
#undef P4
}

template <int ST, typename MaskT>
static cudaError_t launch_prune_any(phylo_engine *e, const double *Pl, const double *Pr, const Operand &l,
                                    const Operand &r, double *out, int32_t *osc) {
  constexpr int TP = 128;
  const int S = e->S, S4 = (S + 3) & ~3, SP = S | 1;
  const size_t smem = sizeof(double) * (2 * (size_t)S * S4 + 2 * (size_t)TP * SP);
  auto kern = prune_any_kernel<ST, MaskT, TP>;
  cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (st != cudaSuccess) return st;
  const int per_sm = std::max(1, (int)(200 * 1024 / (smem + 1024)));
  const int g = grid_for(e->N, TP, e->sm_count * std::min(per_sm, 8));
  kern<<<g, TP, smem, e->stream>>>(Pl, Pr, l.src, l.scale, l.tip, r.src, r.scale, r.tip, out, osc, e->N, S, e->K);
  return cudaSuccess;
}

// fp64 tensor-core path for S = 20 / 61; returns false when the A fragments of all K rate
// classes do not fit in shared memory (then the FMA kernel above is used)
template <int S, typename MaskT>
static bool launch_prune_mma(phylo_engine *e, const double *Pl, const double *Pr, const Operand &l,
                             const Operand &r, double *out, int32_t *osc, cudaError_t *st) {
  constexpr int MT = (S + 7) / 8, KS = (S + 3) / 4;
  const size_t smem = sizeof(double) * 2 * (size_t)e->K * MT * KS * 32;
  if (smem > 200 * 1024) return false;
#define MMA_LAUNCH(LT, RT)                                                                              \
  {                                                                                                     \
    auto kern = prune_mma_kernel<S, MaskT, LT, RT>;                                                     \
    *st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    if (*st != cudaSuccess) return true;                                                                \
    const int g = resident_grid(e, kern, 256, smem, (e->N + 63) / 64);                                  \
    kern<<<g, 256, smem, e->stream>>>(Pl, Pr, l.src, l.scale, r.src, r.scale, out, osc, e->N, e->K);    \
  }
  if (l.tip && r.tip) MMA_LAUNCH(true, true)
  else if (l.tip) MMA_LAUNCH(true, false)
  else if (r.tip) MMA_LAUNCH(false, true)
  else MMA_LAUNCH(false, false)
#undef MMA_LAUNCH
  return true;
}

// tip+tip for 20 / 61 states: table of the S*S finished vectors + a copy kernel (lk_kernels.cuh);
// returns false when the table cannot be allocated (the DMMA kernel is used then)
template <int S, typename MaskT>
static bool launch_prune_tt_table(phylo_engine *e, const double *Pl, const double *Pr, const Operand &l,
                                  const Operand &r, double *out, int32_t *osc, cudaError_t *st) {
  const size_t rows = (size_t)S * S, need = rows * e->K * S;
  if (need > e->capTT) {
    cudaStreamSynchronize(e->stream);
    dfree(e->dTT); dfree(e->dTTsc);
    e->capTT = 0;
    if (cudaMalloc(&e->dTT, sizeof(double) * need) != cudaSuccess ||
        cudaMalloc(&e->dTTsc, sizeof(int32_t) * rows) != cudaSuccess) {
      cudaGetLastError();
      dfree(e->dTT); dfree(e->dTTsc);
      return false;
    }
    e->capTT = need;
  }
  tt_table_kernel<S><<<(unsigned)rows, 128, 0, e->stream>>>(Pl, Pr, e->K, e->dTT, e->dTTsc);
  ++e->launches;
  auto kern = prune_tt_copy_kernel<S, MaskT>;
  const int g = resident_grid(e, kern, 256, 0, (e->N + 63) / 64);
  kern<<<g, 256, 0, e->stream>>>(Pl, Pr, e->dTT, e->dTTsc, (const MaskT *)l.src, (const MaskT *)r.src, out, osc,
                                 e->N, e->K);
  *st = cudaSuccess;
  return true;
}

static int lk_launch_prune(phylo_engine *e, const double *Pl, const double *Pr, const Operand &l,
                           const Operand &r, double *out, int32_t *osc) {
  ProfScope prof(e, (l.tip && r.tip) ? KC_PRUNE_TT : ((l.tip || r.tip) ? KC_PRUNE_TI : KC_PRUNE_II));
  if (e->S == 4 && (e->K == 1 || e->K == 2 || e->K == 4 || e->K == 8 || e->K == 16)) {
    switch (e->K) {
      case 1: launch_prune4<1>(e, Pl, Pr, l, r, out, osc); break;
      case 2: launch_prune4<2>(e, Pl, Pr, l, r, out, osc); break;
      case 4: launch_prune4<4>(e, Pl, Pr, l, r, out, osc); break;
      case 8: launch_prune4<8>(e, Pl, Pr, l, r, out, osc); break;
      default: launch_prune4<16>(e, Pl, Pr, l, r, out, osc);
    }
  } else {
    cudaError_t st = cudaSuccess;
    bool tt_done = false;
    if (l.tip && r.tip && e->opt_tt_table) {
      // 20 states only: measured 99.7 -> 83.1 us per tip+tip update (config 4); for 61 states the
      // table path was slower than the DMMA kernel's one-hot lookups (63.5 -> 74.5 us) and is not used.
      // A chunk-per-thread copy kernel + a second kernel for the ambiguous patterns was slower for both
      // (105 / 104 us) and was removed.
      if (e->S == 20) tt_done = launch_prune_tt_table<20, uint32_t>(e, Pl, Pr, l, r, out, osc, &st);
    }
    if (tt_done) {}
    else if (e->S == 20 && launch_prune_mma<20, uint32_t>(e, Pl, Pr, l, r, out, osc, &st)) {}
    else if (e->S == 61 && launch_prune_mma<61, uint64_t>(e, Pl, Pr, l, r, out, osc, &st)) {}
    else if (e->S == 20) st = launch_prune_any<20, uint32_t>(e, Pl, Pr, l, r, out, osc);
    else if (e->S == 61) st = launch_prune_any<61, uint64_t>(e, Pl, Pr, l, r, out, osc);
    else if (e->mask_dev_bytes == 1) st = launch_prune_any<0, uint8_t>(e, Pl, Pr, l, r, out, osc);
    else if (e->mask_dev_bytes == 4) st = launch_prune_any<0, uint32_t>(e, Pl, Pr, l, r, out, osc);
    else st = launch_prune_any<0, uint64_t>(e, Pl, Pr, l, r, out, osc);
    if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "prune_any setup: %s", cudaGetErrorString(st));
  }
  LAUNCH_CHECK();
  return PHYLO_OK;
}

template <int K>
static void launch_root4(phylo_engine *e, const double *Pr, const Operand &a, const Operand &b) {
  const int g = grid_for(e->nPart, 1, e->sm_count * 4);
#define R4(AT, BT)                                                                                \
  root4_kernel<K, AT, BT><<<g, 256, 0, e->stream>>>(Pr, e->dPi, e->dProbs, e->pinvar,             \
                                                    (const uint8_t *)e->dInv, a.src, a.scale, b.src, \
                                                    b.scale, e->dWeights, e->dSite, e->dPart, e->N)
  if (a.tip && b.tip) R4(true, true);
  else if (a.tip) R4(true, false);
  else if (b.tip) R4(false, true);
  else R4(false, false);
#undef R4
}

template <int ST, typename MaskT>
static cudaError_t launch_root_any(phylo_engine *e, const double *Pr, const Operand &a, const Operand &b) {
  constexpr int TP = 128;
  const int S = e->S, S4 = (S + 3) & ~3, SP = S | 1;
  const size_t smem = sizeof(double) * ((size_t)S * S4 + 2 * (size_t)TP * SP + S);
  auto kern = root_any_kernel<ST, MaskT, TP>;
  cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (st != cudaSuccess) return st;
  const int g = grid_for(e->nPart, 1, e->sm_count * 2);
  kern<<<g, TP, smem, e->stream>>>(Pr, e->dPi, e->dProbs, e->pinvar, (const MaskT *)e->dInv, a.src, a.scale,
                                   a.tip, b.src, b.scale, b.tip, e->dWeights, e->dSite, e->dPart, e->N, S, e->K);
  return cudaSuccess;
}

// fp64 tensor-core root join for S = 20 / 61 (root_mma_kernel): site values first, then the
// canonical fold over them; returns false when the fragments do not fit in shared memory
template <int S, typename MaskT>
static bool launch_root_mma(phylo_engine *e, const double *Pr, const Operand &a, const Operand &b, cudaError_t *st) {
  constexpr int MT = (S + 7) / 8, KS = (S + 3) / 4;
  const size_t smem = sizeof(double) * ((size_t)e->K * MT * KS * 32 + S);
  if (smem > 200 * 1024) return false;
  if (!e->dWSite) {
    *st = cudaMalloc(&e->dWSite, sizeof(double) * (size_t)e->N);
    if (*st != cudaSuccess) return true;
  }
#define ROOT_MMA(AT, BT)                                                                                        \
  {                                                                                                             \
    auto kern = root_mma_kernel<S, MaskT, AT, BT>;                                                              \
    *st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                   \
    if (*st != cudaSuccess) return true;                                                                        \
    const int g = resident_grid(e, kern, 256, smem, (e->N + 63) / 64);                                          \
    kern<<<g, 256, smem, e->stream>>>(Pr, e->dPi, e->dProbs, e->pinvar, (const MaskT *)e->dInv, a.src, a.scale, \
                                      b.src, b.scale, e->dWeights, e->dSite, e->dWSite, e->N, e->K);            \
  }
  if (a.tip && b.tip) ROOT_MMA(true, true)
  else if (a.tip) ROOT_MMA(true, false)
  else if (b.tip) ROOT_MMA(false, true)
  else ROOT_MMA(false, false)
#undef ROOT_MMA
  *st = cudaSuccess;
  return true;
}

// root-edge join with transition matrices Pr ([K][S][S] on device) -> *lnl_host (pinned slot)
static int lk_root_eval(phylo_engine *e, const double *Pr, const Operand &a, const Operand &b, double *slot) {
  {
  ProfScope prof(e, KC_ROOT);
  if (e->S == 4 && (e->K == 1 || e->K == 2 || e->K == 4 || e->K == 8 || e->K == 16)) {
    switch (e->K) {
      case 1: launch_root4<1>(e, Pr, a, b); break;
      case 2: launch_root4<2>(e, Pr, a, b); break;
      case 4: launch_root4<4>(e, Pr, a, b); break;
      case 8: launch_root4<8>(e, Pr, a, b); break;
      default: launch_root4<16>(e, Pr, a, b);
    }
  } else {
    cudaError_t st;
    bool mma = false;
    if (e->S == 20) mma = launch_root_mma<20, uint32_t>(e, Pr, a, b, &st);
    else if (e->S == 61) mma = launch_root_mma<61, uint64_t>(e, Pr, a, b, &st);
    if (mma) {
      if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "root_mma setup: %s", cudaGetErrorString(st));
      LAUNCH_CHECK();
      // level 1 of the canonical fold: the weighted site values -> per-1024-block partials
      reduce1024_kernel<<<(int)e->nPart, 256, 0, e->stream>>>(e->dWSite, e->N, e->dPart);
    }
    else if (e->S == 20) st = launch_root_any<20, uint32_t>(e, Pr, a, b);
    else if (e->S == 61) st = launch_root_any<61, uint64_t>(e, Pr, a, b);
    else if (e->mask_dev_bytes == 1) st = launch_root_any<0, uint8_t>(e, Pr, a, b);
    else if (e->mask_dev_bytes == 4) st = launch_root_any<0, uint32_t>(e, Pr, a, b);
    else st = launch_root_any<0, uint64_t>(e, Pr, a, b);
    if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "root_any setup: %s", cudaGetErrorString(st));
  }
  LAUNCH_CHECK();
  }
  return lk_finish_reduce(e, slot);
}

// remaining levels of the canonical reduction over the level-1 partials, then D2H of the sum
static int lk_finish_reduce(phylo_engine *e, double *slot) {
  ProfScope prof(e, KC_REDUCE);
  const double *cur = e->dPart;
  int64_t n = e->nPart;
  double *bufs[2] = {e->dPart2, e->dPart2 + ((e->nPart + kLnlBlock - 1) / kLnlBlock + 1)};
  int flip = 0;
  do {
    const int64_t nb = (n + kLnlBlock - 1) / kLnlBlock;
    reduce1024_kernel<<<(int)nb, 256, 0, e->stream>>>(cur, n, bufs[flip]);
    LAUNCH_CHECK();
    cur = bufs[flip];
    flip ^= 1;
    n = nb;
  } while (n > 1);
  CK(cudaMemcpyAsync(slot, cur, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  return PHYLO_OK;
}

extern "C" int phylo_engine_set_option(phylo_engine *e, int option, int64_t value) {
  if (!e) return PHYLO_ERR_ARG;
  switch (option) {
    case PHYLO_OPT_FUSED_TREE: e->opt_fused = (value == 2) ? 2 : (value != 0); return PHYLO_OK;
    case PHYLO_OPT_RETAIN_CLV: e->opt_retain = value != 0; return PHYLO_OK;
    case PHYLO_OPT_FITCH_WALK: e->opt_fitch_walk = (value < 0 || value > 3) ? 1 : (int)value; return PHYLO_OK;
    case PHYLO_OPT_DEFER_SCALAR: e->defer_scalar = value != 0; return PHYLO_OK;
    default: return fail(e, PHYLO_ERR_ARG, "set_option: unknown option %d", option);
  }
}
extern "C" int phylo_engine_get_option(phylo_engine *e, int option, int64_t *value) {
  if (!e || !value) return PHYLO_ERR_ARG;
  switch (option) {
    case PHYLO_OPT_FUSED_TREE: *value = e->opt_fused; return PHYLO_OK;
    case PHYLO_OPT_RETAIN_CLV: *value = e->opt_retain; return PHYLO_OK;
    case PHYLO_OPT_FITCH_WALK: *value = e->opt_fitch_walk; return PHYLO_OK;
    case PHYLO_OPT_DEFER_SCALAR: *value = e->defer_scalar; return PHYLO_OK;
    default: return fail(e, PHYLO_ERR_ARG, "get_option: unknown option %d", option);
  }
}

// A schedule compiled for the tree-fused kernels: one step per median in an order that keeps
// the running result in registers, every operand tagged TIP / STORED / CUR / POP.
struct PlanStep {
  int lkind, rkind, lidx, ridx, push_first, out_slot;
  double t_left, t_right;
};
struct FusedPlan {
  std::vector<PlanStep> steps;  // n_ops medians + the root-edge join (out_slot = -1)
  int depth = 0;                // stack levels needed
};

// Re-derives a depth-first order of the (already validated, post-order) schedule in which
// the child with the larger stack need is evaluated first, and assigns every operand one of
// TIP / STORED / CUR (just computed, in registers) / POP (parked on the stack).
// Returns false when the schedule is not a plain tree (a result used twice or never).
static bool build_fused_plan(int cap, int T, const phylo_op *ops, int n_ops, int ra, int rb, double rt,
                             FusedPlan &pl) {
  if (n_ops > 100000) return false;
  std::vector<int> producer(cap, -1), uses(cap, 0), need(cap, 0);
  for (int o = 0; o < n_ops; ++o) {
    if (producer[ops[o].parent] != -1) return false;
    producer[ops[o].parent] = o;
    ++uses[ops[o].left];
    ++uses[ops[o].right];
  }
  ++uses[ra];
  ++uses[rb];
  if (ra == rb) return false;
  for (int s = 0; s < cap; ++s)
    if (producer[s] >= 0 && uses[s] != 1) return false;
  // an op that reads a slot which a LATER op overwrites means "the old contents": only the
  // sequential per-node path honours that
  for (int o = 0; o < n_ops; ++o)
    if (producer[ops[o].left] > o || producer[ops[o].right] > o) return false;
  auto computed = [&](int s) { return producer[s] >= 0; };
  for (int o = 0; o < n_ops; ++o) {
    const int l = ops[o].left, r = ops[o].right, p = ops[o].parent;
    const bool cl = computed(l), cr = computed(r);
    if (cl && cr) need[p] = need[l] == need[r] ? need[l] + 1 : std::max(need[l], need[r]);
    else need[p] = cl ? need[l] : (cr ? need[r] : 0);
  }
  bool live = false;
  int depth = 0, maxdepth = 0;
  auto kind = [&](int child, bool both, int first, int &idx) {
    idx = child;
    if (child < T) return (int)OPK_TIP;
    if (!computed(child)) return (int)OPK_STORED;
    return (both && child == first) ? (int)OPK_POP : (int)OPK_CUR;
  };
  // iterative post-order with explicit frames (caterpillars are thousands deep)
  struct Frame { int slot, stage; };
  auto emit_subtree = [&](int root) {
    if (!computed(root)) return;
    std::vector<Frame> st;
    st.push_back({root, 0});
    while (!st.empty()) {
      Frame &f = st.back();
      const phylo_op &op = ops[producer[f.slot]];
      const bool cl = computed(op.left), cr = computed(op.right);
      const int first = (cl && cr) ? (need[op.left] >= need[op.right] ? op.left : op.right)
                                   : (cl ? op.left : (cr ? op.right : -1));
      const int second = (cl && cr) ? (first == op.left ? op.right : op.left) : -1;
      if (f.stage == 0) {
        f.stage = 1;
        if (first >= 0) { st.push_back({first, 0}); continue; }
      }
      if (f.stage == 1) {
        f.stage = 2;
        if (second >= 0) { st.push_back({second, 0}); continue; }
      }
      PlanStep in{};
      in.lkind = kind(op.left, cl && cr, first, in.lidx);
      in.rkind = kind(op.right, cl && cr, first, in.ridx);
      in.push_first = (!cl && !cr && live) ? 1 : 0;
      in.out_slot = op.parent;
      in.t_left = op.t_left;
      in.t_right = op.t_right;
      if (in.push_first) maxdepth = std::max(maxdepth, ++depth);
      if (in.lkind == OPK_POP || in.rkind == OPK_POP) --depth;
      live = true;
      pl.steps.push_back(in);
      st.pop_back();
    }
  };
  const bool ca = computed(ra), cb = computed(rb);
  const int first = (ca && cb) ? (need[ra] >= need[rb] ? ra : rb) : -1;
  if (ca && cb) {
    emit_subtree(first);
    emit_subtree(first == ra ? rb : ra);
  } else {
    emit_subtree(ra);
    emit_subtree(rb);
  }
  PlanStep root{};
  root.lkind = kind(ra, ca && cb, first, root.lidx);
  root.rkind = kind(rb, ca && cb, first, root.ridx);
  root.out_slot = -1;
  root.t_left = root.t_right = rt;
  pl.steps.push_back(root);
  pl.depth = std::max(maxdepth, 1);
  return (int)pl.steps.size() == n_ops + 1;
}

// Host-only view of the plan compiler for tests (no CUDA call): the steps the tree-fused
// likelihood kernels and the Fitch register walk would execute for this schedule.
extern "C" int phylo_plan_compile(const phylo_op *ops, int n_ops, int T, int capacity, int root_a, int root_b,
                                  int32_t *steps_out, int *depth_out) {
  if (n_ops < 0 || (n_ops > 0 && !ops) || T < 1 || capacity < T || !steps_out) return PHYLO_ERR_ARG;
  if (root_a < 0 || root_a >= capacity || root_b < 0 || root_b >= capacity) return PHYLO_ERR_ARG;
  for (int o = 0; o < n_ops; ++o)
    if (ops[o].parent < T || ops[o].parent >= capacity || ops[o].left < 0 || ops[o].left >= capacity ||
        ops[o].right < 0 || ops[o].right >= capacity)
      return PHYLO_ERR_ARG;
  FusedPlan pl;
  if (!build_fused_plan(capacity, T, ops, n_ops, root_a, root_b, 0.0, pl)) return PHYLO_ERR_UNSUPPORTED;
  for (size_t i = 0; i < pl.steps.size(); ++i) {
    const PlanStep &st = pl.steps[i];
    int32_t *row = steps_out + 6 * i;
    row[0] = st.lkind; row[1] = st.lidx; row[2] = st.rkind; row[3] = st.ridx; row[4] = st.push_first; row[5] = st.out_slot;
  }
  if (depth_out) *depth_out = pl.depth;
  return PHYLO_OK;
}

static size_t tree_smem_bytes(int K, int T, int depth, int n_steps) {
  const size_t tile = (size_t)kTreeR * kTreeThreads / K;
  return 2 * tile * 8 + 4 * 8 + (size_t)depth * kTreeR * kTreeThreads * (sizeof(d4) + sizeof(int)) +
         (size_t)(n_steps + 2) * sizeof(TreeInstr) + (size_t)(kTreeThreads / 32) * 2 * 2 * 16 * K * 8 + 128 +
         (size_t)T * (tile / 2);
}

template <int K>
static cudaError_t launch_tree(phylo_engine *e, TreeArgs args, size_t smem, bool retain, int64_t tile_begin,
                               int64_t tile_end, cudaStream_t cs) {
  args.tile_begin = tile_begin;
  args.tile_end = tile_end;
  const int64_t ntiles = tile_end - tile_begin;
  cudaError_t st;
  int occ = 1;
  if (retain) {
    auto kern = lk_tree4_kernel<K, true>;
    if ((st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return st;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTreeThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    const int g = (int)std::min<int64_t>(ntiles, (int64_t)e->sm_count * occ);
    kern<<<g, kTreeThreads, smem, cs>>>(args);
  } else {
    auto kern = lk_tree4_kernel<K, false>;
    if ((st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return st;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTreeThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    const int g = (int)std::min<int64_t>(ntiles, (int64_t)e->sm_count * occ);
    kern<<<g, kTreeThreads, smem, cs>>>(args);
  }
  ++e->launches;
  return cudaGetLastError();
}

static cudaError_t launch_tree_k(phylo_engine *e, const TreeArgs &a, size_t smem, int64_t tile_begin, int64_t tile_end,
                                 cudaStream_t cs) {
  switch (e->K) {
    case 1: return launch_tree<1>(e, a, smem, e->opt_retain, tile_begin, tile_end, cs);
    case 2: return launch_tree<2>(e, a, smem, e->opt_retain, tile_begin, tile_end, cs);
    case 4: return launch_tree<4>(e, a, smem, e->opt_retain, tile_begin, tile_end, cs);
    default: return launch_tree<8>(e, a, smem, e->opt_retain, tile_begin, tile_end, cs);
  }
}

// ---- warp-autonomous tree kernel (lk_treew_kernel.cuh): geometry and launch
struct TreeWGeom {
  int warps = 0;     // warps per CTA
  size_t smem = 0;   // dynamic shared memory per CTA
  int slev = kTreeWSmemLevels, R = 1, interleave = 1;
};
// R = 2 (two patterns per thread share every matrix read) pays when 8 warps of it fit: it needs
// twice the per-warp shared memory, and 4 or 8 warps to keep the four SM sub-partitions even.
// Otherwise R = 1 with as many warps as fit. PHYLO_TREEW_TUNE="slev,R,interleave" overrides
// (experiments; R = 0: automatic).
static TreeWGeom treew_geometry(const phylo_engine *e, int depth, int n_steps, bool retain) {
  const size_t kMaxSmem = 227 * 1024, fixed = treew_prog_bytes(n_steps) + 1024;  // +1024: manual alignment
  const int64_t ngroups = (e->N + 31) / 32;
  int want_slev = 0, want_R = 0, want_il = 1;
  if (const char *t = getenv("PHYLO_TREEW_TUNE")) sscanf(t, "%d,%d,%d", &want_slev, &want_R, &want_il);
  auto fit = [&](int R, int slev, int max_warps) {
    TreeWGeom g;
    g.R = R; g.slev = slev; g.interleave = want_il != 0;
    const size_t wb = treew_stage_bytes(e->K, retain, R) + treew_warp_bytes(e->K, e->T, depth, slev, R);
    if (fixed + wb > kMaxSmem) return g;
    int w = (int)std::min<size_t>(max_warps, (kMaxSmem - fixed) / wb);
    // few units (small alignments): spread them over the SMs instead of filling CTAs
    const int64_t units = (ngroups + R - 1) / R;
    w = (int)std::max<int64_t>(1, std::min<int64_t>(w, (units + e->sm_count - 1) / e->sm_count));
    g.warps = w;
    g.smem = std::max(fixed + (size_t)w * wb, (size_t)kTreeWFuseBytes + 1024);  // the fused final fold reuses it
    return g;
  };
  if (want_R == 1 || want_R == 2) return fit(want_R, want_slev > 0 ? want_slev : kTreeWSmemLevels, want_R == 1 ? kTreeWMaxWarps : 8);
  if (want_slev > 0) return fit(1, want_slev, kTreeWMaxWarps);
  if (ngroups >= (int64_t)e->sm_count * 16) {  // enough work for 8 double-width warps per SM
    for (int slev = kTreeWSmemLevels; slev >= 1; --slev) {
      TreeWGeom g = fit(2, slev, 8);
      if (g.warps == 8) return g;
    }
  }
  return fit(1, kTreeWSmemLevels, kTreeWMaxWarps);
}

template <int K, int R>
static cudaError_t launch_treew(phylo_engine *e, const TreeWArgs &wargs, const TreeWGeom &geo, bool retain, int64_t g_begin,
                                int64_t g_end, cudaStream_t cs) {
  TreeWArgs args = wargs;
  args.t.tile_begin = g_begin;
  args.t.tile_end = g_end;
  const int64_t units = (g_end - g_begin + R - 1) / R, ctas = (units + geo.warps - 1) / geo.warps;
  const int g = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, e->sm_count));
  cudaError_t st;
  if (retain) {
    auto kern = lk_treew_kernel<K, R, true>;
    if ((st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem)) != cudaSuccess) return st;
    kern<<<g, geo.warps * 32, geo.smem, cs>>>(args);
  } else {
    auto kern = lk_treew_kernel<K, R, false>;
    if ((st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem)) != cudaSuccess) return st;
    kern<<<g, geo.warps * 32, geo.smem, cs>>>(args);
  }
  ++e->launches;
  return cudaGetLastError();
}

static cudaError_t launch_treew_k(phylo_engine *e, const TreeArgs &a, const TreeWGeom &geo, int64_t g_begin, int64_t g_end,
                                  cudaStream_t cs, const TreeWArgs *fuse = nullptr) {
  TreeWArgs w;
  if (fuse) w = *fuse;
  else { w.fuse_reduce = 0; w.inline_prog = 0; }
  w.t = a;
  w.tmaps = (const char *)e->dTmaps;
#define TREEW(KV)                                                                             \
  (geo.R == 2 ? launch_treew<KV, 2>(e, w, geo, e->opt_retain, g_begin, g_end, cs)             \
              : launch_treew<KV, 1>(e, w, geo, e->opt_retain, g_begin, g_end, cs))
  switch (e->K) {
    case 1: return TREEW(1);
    case 2: return TREEW(2);
    default: return TREEW(4);
  }
#undef TREEW
}

// (re)builds the TMA tensor maps of all allocated node CLVs: [N rows][4K doubles], box =
// 32 rows, swizzle = row bytes (128/64/32)
static int lk_build_tmaps(phylo_engine *e) {
  std::vector<CUtensorMap> h(e->cap);
  std::memset(h.data(), 0, sizeof(CUtensorMap) * e->cap);
  const CUtensorMapSwizzle swz = e->K == 4 ? CU_TENSOR_MAP_SWIZZLE_128B
                                           : (e->K == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  for (int s = 0; s < e->cap; ++s) {
    if (!e->nodes[s].clv) continue;
    if (e->S != 4) {
      // tree-fused 20-state kernel: the CLV [N][K][S] seen as dims (S, N, K); a box (S, 8, K) is one 8-pattern
      // group in the kernel's shared-memory order [k][pattern][S]
      const cuuint64_t gdim[3] = {(cuuint64_t)e->S, (cuuint64_t)e->N, (cuuint64_t)e->K};
      const cuuint64_t gstride[2] = {(cuuint64_t)e->K * e->S * 8, (cuuint64_t)e->S * 8};
      const cuuint32_t box[3] = {(cuuint32_t)e->S, 8, (cuuint32_t)e->K};
      const cuuint32_t estr[3] = {1, 1, 1};
      const CUresult r = e->encodeTiled(&h[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, e->nodes[s].clv, gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(e, PHYLO_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed for node slot %d (CUresult %d)", s, (int)r);
      continue;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)(4 * e->K), (cuuint64_t)e->N};
    const cuuint64_t gstride[1] = {(cuuint64_t)(32 * e->K)};
    const cuuint32_t box[2] = {(cuuint32_t)(4 * e->K), 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = e->encodeTiled(&h[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, e->nodes[s].clv, gdim, gstride, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, PHYLO_ERR_CUDA, "cuTensorMapEncodeTiled failed for node slot %d (CUresult %d)", s, (int)r);
  }
  CK(cudaMemcpy(e->dTmaps, h.data(), sizeof(CUtensorMap) * e->cap, cudaMemcpyHostToDevice));
  e->tmapDirty = false;
  return PHYLO_OK;
}

// returns PHYLO_OK with *done = true when the fused kernel handled the evaluation
// host_masks != NULL: the alignment is still on the host; it is uploaded in slabs on a second
// stream while earlier slabs are already being scored (phylo_lk_score_alignment)
static int lk_score_tree_fused(phylo_engine *e, const phylo_op *ops, int n_ops, int ra, int rb, double rt,
                               bool *done, const void *host_masks = nullptr, int mask_bytes = 0) {
  *done = false;
  if (!e->opt_fused || e->S != 4 || e->mask_dev_bytes != 1 || !e->dTips4) return PHYLO_OK;
  if (!(e->K == 1 || e->K == 2 || e->K == 4 || e->K == 8)) return PHYLO_OK;
  FusedPlan pl;
  if (!build_fused_plan(e->cap, e->T, ops, n_ops, ra, rb, rt, pl)) return PHYLO_OK;
  const size_t kMaxSmem = 227 * 1024;
  // warp-autonomous kernel (thread = pattern) for K <= 4; the tile kernel (thread = pattern x
  // rate class) for K = 8, or when a warp's working set does not fit
  TreeWGeom geo;
  if (e->opt_fused == 1 && e->K <= 4 && e->encodeTiled)
    geo = treew_geometry(e, pl.depth, (int)pl.steps.size(), e->opt_retain);
  const bool useW = geo.warps > 0;
  const size_t smem = useW ? geo.smem : tree_smem_bytes(e->K, e->T, pl.depth, (int)pl.steps.size());
  if (smem > kMaxSmem) return PHYLO_OK;  // very deep / very wide trees: per-node kernels instead
  int rc;
  const int nb = 2 * n_ops + 1;
  if ((rc = ensure_pt_capacity(e, nb, e->S, e->K)) != PHYLO_OK) return rc;
  if (e->opt_retain)
    for (int o = 0; o < n_ops; ++o)
      if ((rc = lk_ensure_node(e, ops[o].parent)) != PHYLO_OK) return rc;
  const size_t pbytes = sizeof(TreeInstr) * pl.steps.size();
  if (pbytes > e->capProg) {
    CK(cudaStreamSynchronize(e->stream));
    dfree(e->dProg);
    if (e->hProg) { cudaFreeHost(e->hProg); e->hProg = nullptr; }
    e->capProg = 0;
    CK(cudaMalloc(&e->dProg, pbytes * 2));
    CK(cudaMallocHost(&e->hProg, pbytes * 2));
    e->capProg = pbytes * 2;
  }
  CK(cudaStreamSynchronize(e->stream));  // pinned staging (hT, hProg) is about to be rewritten
  if (e->nodeTabDirty) {
    std::vector<double *> hc(e->cap);
    std::vector<int32_t *> hs(e->cap);
    for (int s = 0; s < e->cap; ++s) { hc[s] = e->nodes[s].clv; hs[s] = e->nodes[s].scale; }
    CK(cudaMemcpy(e->dNodeClv, hc.data(), sizeof(double *) * e->cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->dNodeSc, hs.data(), sizeof(int32_t *) * e->cap, cudaMemcpyHostToDevice));
    e->nodeTabDirty = false;
  }
  if (useW && e->opt_retain && e->tmapDirty && (rc = lk_build_tmaps(e)) != PHYLO_OK) return rc;
  {
    TreeInstr *hp = (TreeInstr *)e->hProg;
    TreeWInstr *hw = (TreeWInstr *)e->hProg;
    for (size_t i = 0; i < pl.steps.size(); ++i) {
      const PlanStep &st = pl.steps[i];
      if (useW) {
        hw[i] = TreeWInstr{st.lkind | (st.rkind << 2) | (st.push_first << 4), st.lidx, st.ridx,
                           e->opt_retain ? st.out_slot : -1};
      } else {
        TreeInstr in{};
        in.kinds = st.lkind | (st.rkind << 2) | (st.push_first << 4);
        in.lidx = st.lidx;
        in.ridx = st.ridx;
        in.out_slot = st.out_slot;
        in.out_clv = (e->opt_retain && st.out_slot >= 0) ? e->nodes[st.out_slot].clv : nullptr;
        in.out_sc = (e->opt_retain && st.out_slot >= 0) ? e->nodes[st.out_slot].scale : nullptr;
        hp[i] = in;
      }
      if (i + 1 < pl.steps.size()) { e->hT[2 * i] = st.t_left; e->hT[2 * i + 1] = st.t_right; }
      else e->hT[2 * i] = st.t_left;
    }
  }
  // small alignments on the warp-autonomous kernel: program in the kernel parameters, final
  // reduction and result publication inside the kernel (see TreeWArgs)
  const bool fuse = useW && !host_masks && e->nPart <= kLnlBlock;
  TreeWArgs fargs;
  fargs.fuse_reduce = 0;
  fargs.inline_prog = 0;
  if (fuse) {
    if (!e->dTreeDone) {
      CK(cudaMalloc(&e->dTreeDone, sizeof(unsigned int)));
      CK(cudaMemsetAsync(e->dTreeDone, 0, sizeof(unsigned int), e->stream));
    }
    fargs.fuse_reduce = 1;
    fargs.n_part = e->nPart;
    fargs.partials = e->dPart;
    fargs.done_counter = e->dTreeDone;
    double *dev_out = nullptr;
    CK(cudaHostGetDevicePointer((void **)&dev_out, e->hScalar + HS_TREE_OUT, 0));
    fargs.host_out = dev_out;
    fargs.seq = ++e->treeSeq;
    const size_t wbytes = sizeof(TreeWInstr) * pl.steps.size();
    if (wbytes <= (size_t)kTreeWInlineProg) {
      std::memcpy(fargs.prog_inline, e->hProg, wbytes);
      fargs.inline_prog = 1;
    }
  }
  if (!fargs.inline_prog) CK(cudaMemcpyAsync(e->dProg, e->hProg, pbytes, cudaMemcpyHostToDevice, e->stream));
  if ((rc = build_pt(e, nb, useW ? 0 : 1)) != PHYLO_OK) return rc;
  TreeArgs a;
  a.prog = (const TreeInstr *)e->dProg;
  a.n_instr = n_ops;
  a.P = e->dP;
  a.tips4 = e->dTips4;
  a.tip_stride = e->tipStride;
  a.T = e->T;
  a.N = e->N;
  a.node_clv = e->dNodeClv;
  a.node_sc = e->dNodeSc;
  a.pi = e->dPi;
  a.probs = e->dProbs;
  a.pinvar = e->pinvar;
  a.inv = (const uint8_t *)e->dInv;
  a.weights = e->dWeights;
  a.site_lnl = e->dSite;
  a.groups = e->dGroups;
  a.stack_depth = pl.depth;
  a.spill = nullptr;
  if (useW) {
    // two regions: consecutive slabs of phylo_lk_score_alignment run on two streams
    const size_t region = treew_spill_bytes(e->K, pl.depth, geo.slev, geo.R) * kTreeWMaxWarps * (size_t)e->sm_count, need = 2 * region;
    a.smem_levels = geo.slev;
    a.interleave = geo.interleave;
    if (need > e->capSpill) {
      CK(cudaStreamSynchronize(e->stream));
      dfree(e->dSpill);
      e->capSpill = 0;
      CK(cudaMalloc(&e->dSpill, need));
      e->capSpill = need;
    }
    a.spill = (double2 *)e->dSpill;
  }
  const int64_t tile = useW ? 32 : (int64_t)kTreeR * kTreeThreads / e->K, ntiles = (e->N + tile - 1) / tile;
  if (!host_masks) {
    ProfScope prof(e, KC_TREE_FUSED);
    if (fuse) {
      volatile double *flag = e->hScalar + HS_TREE_OUT + 1;
      *flag = -1.0;  // sequence numbers are positive integers' bit patterns: never this value
    }
    cudaError_t st = useW ? launch_treew_k(e, a, geo, 0, ntiles, e->stream, fuse ? &fargs : nullptr)
                          : launch_tree_k(e, a, smem, 0, ntiles, e->stream);
    if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "tree-fused launch: %s", cudaGetErrorString(st));
  } else {
    // slabs of whole 1024-pattern blocks: >= ~2 waves of tiles each, at most 16 slabs
    const int64_t blocks = e->nPart;
    // (64 blocks = 2048 warp-groups of 32 patterns, ~1.7 waves of the tree kernel; consecutive slabs run on two
    // streams, so their tails overlap. A rank of an 8-GPU run holds 512 blocks: 8 slabs instead of the 2 that a
    // 256-block minimum gave it -- with 2 slabs upload and scoring barely overlapped: e2e 6.2 ms for 3.2 ms of work)
    int nslab = (int)std::max<int64_t>(1, std::min<int64_t>(16, blocks / 64));
    // The warp-autonomous kernel hands a launch's units (R groups of 32 patterns) to sm_count x warps warps in
    // rounds; a CTA leaves when its slowest warp is done. A slab of 244 blocks (4 M patterns in 16 slabs) is 3.3
    // rounds: every CTA stays for 4 while 5 of its 8 warps idle through the last one. So a slab is a whole number
    // of rounds (74 blocks each with 148 SMs x 8 warps x 2 groups) whenever that is close to the size wanted.
    // The first slab's upload is the part of the transfer nothing hides: the slabs grow from one round to the
    // size wanted (q, 2 q, ... blocks).
    std::vector<int64_t> slab_end;  // empty: nslab equal parts
    if (useW) {
      const int64_t round_units = (int64_t)e->sm_count * geo.warps, block_units = kLnlBlock / (32 * geo.R);
      const int64_t q = round_units / std::gcd(round_units, block_units);  // blocks in the smallest whole-round slab
      const int64_t want = blocks / nslab;
      if (nslab > 1 && q <= want + want / 2) {
        const int64_t full = std::max<int64_t>(1, (want + q / 2) / q) * q;
        for (int64_t cur = 0, step = q; cur < blocks; step = std::min(step + q, full)) {
          cur = std::min(blocks, cur + step);
          if (blocks - cur < q / 2) cur = blocks;  // a small rest joins the last slab
          slab_end.push_back(cur);
        }
        nslab = (int)slab_end.size();
      }
    }
    if (!e->copyStream) CK(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
    if (!e->copyStream2) CK(cudaStreamCreateWithFlags(&e->copyStream2, cudaStreamNonBlocking));
    if (!e->auxStream) CK(cudaStreamCreateWithFlags(&e->auxStream, cudaStreamNonBlocking));
    if (!e->auxDone) CK(cudaEventCreateWithFlags(&e->auxDone, cudaEventDisableTiming));
    while ((int)e->slabEvents.size() < nslab + 1) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->slabEvents.push_back(ev);
    }
    // the aux stream starts after the set-up work already queued on the engine's stream
    // (program upload, pt_build, weights); the copy stream needs nothing: the staging buffer is
    // idle (the engine's stream was synchronised above)
    CK(cudaEventRecord(e->slabEvents[nslab], e->stream));
    CK(cudaStreamWaitEvent(e->auxStream, e->slabEvents[nslab], 0));
    for (int sidx = 0; sidx < nslab; ++sidx) {
      // consecutive slabs alternate between two compute streams so that the tail of one
      // slab's kernel overlaps the head of the next
      cudaStream_t cs = (sidx & 1) ? e->auxStream : e->stream;
      const int64_t b_lo = slab_end.empty() ? blocks * sidx / nslab : (sidx ? slab_end[sidx - 1] : 0);
      const int64_t b_hi = slab_end.empty() ? blocks * (sidx + 1) / nslab : slab_end[sidx];
      const int64_t p_lo = b_lo * kLnlBlock, p_hi = std::min<int64_t>(e->N, b_hi * kLnlBlock);
      if ((rc = lk_upload_slab(e, host_masks, mask_bytes, p_lo, p_hi, (sidx & 1) ? e->copyStream2 : e->copyStream,
                               e->slabEvents[sidx], cs)) != PHYLO_OK)
        return rc;
      if (useW) a.spill = (double2 *)((char *)e->dSpill + (size_t)(sidx & 1) * (e->capSpill / 2));
      cudaError_t st = useW ? launch_treew_k(e, a, geo, p_lo / tile, (p_hi + tile - 1) / tile, cs)
                            : launch_tree_k(e, a, smem, p_lo / tile, (p_hi + tile - 1) / tile, cs);
      if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "tree-fused launch: %s", cudaGetErrorString(st));
    }
    CK(cudaEventRecord(e->auxDone, e->auxStream));
    CK(cudaStreamWaitEvent(e->stream, e->auxDone, 0));
  }
  if (e->opt_retain)
    for (int o = 0; o < n_ops; ++o) e->nodes[ops[o].parent].valid = true;
  else
    for (int o = 0; o < n_ops; ++o) e->nodes[ops[o].parent].valid = false;
  e->fused_result_ready = false;
  if (fuse && e->defer_scalar) {  // the caller will combine the block partials on the device (phylo_lk_exchange_reduce)
    *done = true;
    return PHYLO_OK;
  }
  if (fuse) {
    // spin on the sequence number the last CTA writes after lnL (mapped host memory); a
    // blocking stream sync would cost more than these kernels. Fallback after 2 ms.
    const double want = [&] { double d; const long long q = (long long)fargs.seq; std::memcpy(&d, &q, 8); return d; }();
    volatile double *out = e->hScalar + HS_TREE_OUT;
    bool seen = false;
    const auto t0 = std::chrono::steady_clock::now();
    for (int spin = 0;; ++spin) {
      double f = out[1];
      if (std::memcmp(&f, &want, 8) == 0) { seen = true; break; }
      if ((spin & 1023) == 1023 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
    }
    if (!seen) {
      CK(cudaStreamSynchronize(e->stream));
      double f = out[1];
      if (std::memcmp(&f, &want, 8) != 0) return fail(e, PHYLO_ERR_CUDA, "lk_score_tree: the tree kernel did not publish its result");
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    e->hScalar[0] = out[0];
    e->fused_result_ready = true;
    *done = true;
    return PHYLO_OK;
  }
  {
    ProfScope prof(e, KC_REDUCE);
    fold_groups_kernel<<<(int)e->nPart, 32, 0, e->stream>>>(e->dGroups, (e->N + 31) / 32, e->dPart);
    LAUNCH_CHECK();
  }
  if (e->defer_scalar) {  // the block partials are what phylo_lk_exchange_reduce needs; no local sum, no D2H
    *done = true;
    return PHYLO_OK;
  }
  if ((rc = lk_finish_reduce(e, e->hScalar)) != PHYLO_OK) return rc;
  *done = true;
  return PHYLO_OK;
}

// ---- tree-fused evaluation for 20 / 61 states (lk_treem_kernel.cuh): one launch for the whole
// schedule, every interior CLV written once into its node slot, nothing read back from HBM but
// the (L2-resident) values parked across a subtree.
template <int S, typename MaskT, int R, int NW, int KT>
static cudaError_t launch_treem(phylo_engine *e, TreeMArgs args, size_t smem) {
  auto kern = lk_treem_kernel<S, MaskT, R, NW, KT>;
  static const bool paired = [] { const char *v = getenv("PHYLO_TREEM_PAIRED"); return !(v && v[0] == '0'); }();
  args.paired = paired ? 1 : 0;
  static const bool st_swap = [] { const char *v = getenv("PHYLO_TREEM_STSWAP"); return !(v && v[0] == '0'); }();
  args.st_swap = st_swap ? 1 : 0;
  args.prog_in_smem = smem + treem_prog_bytes(args.n_steps) <= 227 * 1024;
  if (args.prog_in_smem) smem += treem_prog_bytes(args.n_steps);
  cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (st != cudaSuccess) return st;
  const int64_t ngroups = args.g_end - args.g_begin;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(e->sm_count, (ngroups + NW * R - 1) / (NW * R)));
  kern<<<grid, NW * 32, smem, e->stream>>>(args);
  return cudaGetLastError();
}

// host_masks != NULL (phylo_lk_score_alignment): the alignment is still on the host; it is uploaded in three slabs
// on a copy stream while the slabs before are scored.
static int lk_score_tree_fusedm(phylo_engine *e, const phylo_op *ops, int n_ops, int ra, int rb, double rt, bool *done,
                                const void *host_masks = nullptr, int mask_bytes = 0) {
  *done = false;
  if (!e->opt_fused || !(e->S == 20 || e->S == 61) || n_ops < 1) return PHYLO_OK;
  if (e->S == 20 && !e->encodeTiled) return PHYLO_OK;  // its results leave through TMA tensor maps
  const size_t kMaxSmem = 227 * 1024;
  // 20 states: 16 warps x 2 groups (128 registers; A fragments straight from shared memory); 61 states: 8 warps x
  // 2 groups (the 16 k-steps of B fragments want 255 registers)
  int R = 0;
  const int NW = e->S == 20 ? 16 : 8;
  for (int r : {2, 1}) {
    const size_t need = e->S == 20 ? treem_smem_bytes<20>(e->K, r, NW) : treem_smem_bytes<61>(e->K, r, NW);
    if (need <= kMaxSmem) { R = r; break; }
  }
  if (!R || !NW) return PHYLO_OK;  // too many rate classes for the on-chip tables: per-node kernels
  FusedPlan pl;
  if (!build_fused_plan(e->cap, e->T, ops, n_ops, ra, rb, rt, pl)) return PHYLO_OK;
  int rc;
  const int nb = 2 * n_ops + 1;
  if ((rc = ensure_pt_capacity(e, nb, e->S, e->K)) != PHYLO_OK) return rc;
  for (int o = 0; o < n_ops; ++o)
    if ((rc = lk_ensure_node(e, ops[o].parent)) != PHYLO_OK) return rc;
  const size_t frag = (e->S == 20) ? TreeMGeom<20>::FRAG : TreeMGeom<61>::FRAG;
  const size_t need_frag = (size_t)(nb + 1) * e->K * frag;
  if (need_frag > e->capFrag) {
    CK(cudaStreamSynchronize(e->stream));
    dfree(e->dFrag);
    e->capFrag = 0;
    CK(cudaMalloc(&e->dFrag, sizeof(double) * need_frag * 2));
    e->capFrag = need_frag * 2;
  }
  const size_t pbytes = sizeof(TreeMInstr) * pl.steps.size();
  if (pbytes > e->capProg) {
    CK(cudaStreamSynchronize(e->stream));
    dfree(e->dProg);
    if (e->hProg) { cudaFreeHost(e->hProg); e->hProg = nullptr; }
    e->capProg = 0;
    CK(cudaMalloc(&e->dProg, pbytes * 2));
    CK(cudaMallocHost(&e->hProg, pbytes * 2));
    e->capProg = pbytes * 2;
  }
  if (!e->dWSite) CK(cudaMalloc(&e->dWSite, sizeof(double) * (size_t)e->N));
  CK(cudaStreamSynchronize(e->stream));  // pinned staging (hT, hProg) is about to be rewritten
  if (e->nodeTabDirty) {
    std::vector<double *> hc(e->cap);
    std::vector<int32_t *> hs(e->cap);
    for (int s = 0; s < e->cap; ++s) { hc[s] = e->nodes[s].clv; hs[s] = e->nodes[s].scale; }
    CK(cudaMemcpy(e->dNodeClv, hc.data(), sizeof(double *) * e->cap, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->dNodeSc, hs.data(), sizeof(int32_t *) * e->cap, cudaMemcpyHostToDevice));
    e->nodeTabDirty = false;
  }
  if (e->S == 20 && e->tmapDirty && (rc = lk_build_tmaps(e)) != PHYLO_OK) return rc;
  TreeMInstr *hp = (TreeMInstr *)e->hProg;
  for (size_t i = 0; i < pl.steps.size(); ++i) {
    const PlanStep &st = pl.steps[i];
    auto mode = [](int kind) { return kind == OPK_TIP ? (int)TM_TIP : (kind == OPK_CUR ? (int)TM_CUR : (int)TM_GLB); };
    int lm = mode(st.lkind), rm = mode(st.rkind), li = st.lidx, ri = st.ridx;
    double tl = st.t_left, tr = st.t_right;
    // medians: the step body is compiled per (left mode <= right mode); x * y == y * x bit for bit
    if (i + 1 < pl.steps.size() && lm > rm) { std::swap(lm, rm); std::swap(li, ri); std::swap(tl, tr); }
    hp[i] = TreeMInstr{lm | (rm << 2), li, ri, st.out_slot};
    if (i + 1 < pl.steps.size()) { e->hT[2 * i] = tl; e->hT[2 * i + 1] = tr; }
    else e->hT[2 * i] = tl;
  }
  CK(cudaMemcpyAsync(e->dProg, e->hProg, pbytes, cudaMemcpyHostToDevice, e->stream));
  if ((rc = build_pt(e, nb, 0)) != PHYLO_OK) return rc;
  {
    ProfScope prof(e, KC_PT_BUILD);
    const TreeMInstr *dprog = (const TreeMInstr *)e->dProg;
    if (e->S == 20) pt_frag_kernel<20><<<nb * e->K, 256, 0, e->stream>>>(e->dP, e->dFrag, dprog, n_ops, e->K);
    else pt_frag_kernel<61><<<nb * e->K, 256, 0, e->stream>>>(e->dP, e->dFrag, dprog, n_ops, e->K);
    LAUNCH_CHECK();
  }
  TreeMArgs a;
  a.prog = (const TreeMInstr *)e->dProg;
  a.n_steps = n_ops;
  a.K = e->K;
  a.frags = e->dFrag;
  a.tips = e->dTips;
  a.tip_stride = e->tipStride;
  a.N = e->N;
  a.node_clv = e->dNodeClv;
  a.node_sc = e->dNodeSc;
  a.pi = e->dPi;
  a.probs = e->dProbs;
  a.weights = e->dWeights;
  a.inv = e->dInv;
  a.pinvar = e->pinvar;
  a.site_lnl = e->dSite;
  a.wsite = e->dWSite;
  a.tmaps = (const char *)e->dTmaps;
  a.timing = nullptr;
  static const bool want_timing = [] { const char *v = getenv("PHYLO_TREEM_TIMING"); return v && v[0] == '1'; }();
  unsigned long long *dTiming = nullptr;
  if (want_timing) {
    CK(cudaMalloc(&dTiming, 24 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(dTiming, 0, 24 * sizeof(unsigned long long), e->stream));
    a.timing = dTiming;
  }
  // Slabs (host_masks): whole rounds of the kernel's CTA chunks (NW x R groups of 8 patterns on every SM), so that no
  // slab ends on a partly filled round: q = blocks of 1024 patterns in the smallest such slab. The link is ~10 times
  // faster than the scoring (config 4: 4.8 MB against 1.1 ms per round), so a short first slab is all that stays
  // exposed: slabs of q, 2 q and the rest.
  const int64_t ng_all = (e->N + 7) / 8, blocks = e->nPart;
  int64_t slab_end[3] = {blocks, blocks, blocks};
  int nslab = 1;
  if (host_masks) {
    const int64_t round_pats = (int64_t)e->sm_count * NW * R * 8;
    const int64_t q = round_pats / std::gcd(round_pats, (int64_t)kLnlBlock);
    if (blocks >= 4 * q) {
      nslab = 3;
      slab_end[0] = q;
      slab_end[1] = 3 * q;
    }
    if (!e->copyStream) CK(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
    while ((int)e->slabEvents.size() < nslab) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->slabEvents.push_back(ev);
    }
  }
  for (int sidx = 0; sidx < nslab; ++sidx) {
    const int64_t p_lo = sidx ? slab_end[sidx - 1] * kLnlBlock : 0, p_hi = std::min<int64_t>(e->N, slab_end[sidx] * kLnlBlock);
    if (host_masks &&
        (rc = lk_upload_slab(e, host_masks, mask_bytes, p_lo, p_hi, e->copyStream, e->slabEvents[sidx], e->stream)) != PHYLO_OK)
      return rc;
    a.g_begin = p_lo / 8;
    a.g_end = sidx + 1 == nslab ? ng_all : p_hi / 8;
    ProfScope prof(e, KC_TREE_FUSED);
    cudaError_t st;
#define TREEM(S_, M_, R_, NW_, KT_) launch_treem<S_, M_, R_, NW_, KT_>(e, a, treem_smem_bytes<S_>(e->K, R_, NW_))
    if (e->S == 20) {
      if (e->K == 4) st = R == 2 ? TREEM(20, uint32_t, 2, 16, 4) : TREEM(20, uint32_t, 1, 16, 4);
      else st = R == 2 ? TREEM(20, uint32_t, 2, 16, 0) : TREEM(20, uint32_t, 1, 16, 0);
    } else {
      if (e->K == 1) st = R == 2 ? TREEM(61, uint64_t, 2, 8, 1) : TREEM(61, uint64_t, 1, 8, 1);
      else st = R == 2 ? TREEM(61, uint64_t, 2, 8, 0) : TREEM(61, uint64_t, 1, 8, 0);
    }
#undef TREEM
    if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "tree-fused (DMMA) launch: %s", cudaGetErrorString(st));
    ++e->launches;
  }
  if (dTiming) {
    unsigned long long h[24];
    CK(cudaMemcpy(h, dTiming, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(dTiming);
    static const char *names[6] = {"tip+tip", "tip+cur", "tip+slot", "cur+slot", "slot+slot", "root"};
    for (int v = 0; v < 6; ++v)
      if (h[4 * v + 2])
        fprintf(stderr, "[treem timing] %-9s steps/CTA-chunk %8llu  body %9.0f cyc (of which table wait %7.0f)  barrier %8.0f cyc\n",
                names[v], h[4 * v + 2], (double)h[4 * v] / h[4 * v + 2], (double)h[4 * v + 3] / h[4 * v + 2],
                (double)h[4 * v + 1] / h[4 * v + 2]);
  }
  for (int o = 0; o < n_ops; ++o) e->nodes[ops[o].parent].valid = true;
  e->fused_result_ready = false;
  {
    ProfScope prof(e, KC_REDUCE);
    reduce1024_kernel<<<(int)e->nPart, 256, 0, e->stream>>>(e->dWSite, e->N, e->dPart);
    LAUNCH_CHECK();
  }
  if ((rc = lk_finish_reduce(e, e->hScalar)) != PHYLO_OK) return rc;
  *done = true;
  return PHYLO_OK;
}

extern "C" int phylo_lk_median_2(phylo_engine *e, int parent, int left, double t_left, int right,
                                 double t_right) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_median_2: no tips loaded");
  if (parent < e->T || parent >= e->cap)
    return fail(e, PHYLO_ERR_ARG, "lk_median_2: parent slot %d must be in [%d,%d)", parent, e->T, e->cap);
  if (parent == left || parent == right) return fail(e, PHYLO_ERR_ARG, "lk_median_2: parent aliases a child");
  CK(cudaSetDevice(e->device));
  Operand l, r;
  int rc;
  if ((rc = lk_operand(e, left, &l, "lk_median_2")) != PHYLO_OK) return rc;
  if ((rc = lk_operand(e, right, &r, "lk_median_2")) != PHYLO_OK) return rc;
  if ((rc = ensure_pt_capacity(e, 2, e->S, e->K)) != PHYLO_OK) return rc;
  if ((rc = lk_ensure_node(e, parent)) != PHYLO_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));  // hT is about to be rewritten
  e->hT[0] = t_left;
  e->hT[1] = t_right;
  if ((rc = build_pt(e, 2)) != PHYLO_OK) return rc;
  const size_t pk = (size_t)e->K * e->S * e->S;
  rc = lk_launch_prune(e, e->dP, e->dP + pk, l, r, e->nodes[parent].clv, e->nodes[parent].scale);
  if (rc != PHYLO_OK) return rc;
  e->nodes[parent].valid = true;
  return PHYLO_OK;
}

// Likelihood.median_3 (lib/nodeData.ml:22; `failwith "TODO"` in lib/likelihood_c.ml:16): the node's CLV
// conditioned on all THREE neighbours, L_p = (P_a L_a) o (P_b L_b) o (P_c L_c), with the same per-site
// rescaling. Two ordinary pruning updates: (a, b) into the engine's scratch CLV, then (scratch over a
// zero-length branch -- compose's t < 1e-10 -> identity, lib/mlmodel.c:339-341 -- , c) into `parent`.
extern "C" int phylo_lk_median_3(phylo_engine *e, int parent, int a, double t_a, int b, double t_b, int c, double t_c) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_median_3: no tips loaded");
  if (parent < e->T || parent >= e->cap)
    return fail(e, PHYLO_ERR_ARG, "lk_median_3: parent slot %d must be in [%d,%d)", parent, e->T, e->cap);
  if (parent == a || parent == b || parent == c) return fail(e, PHYLO_ERR_ARG, "lk_median_3: parent aliases a neighbour");
  CK(cudaSetDevice(e->device));
  Operand oa, ob, oc;
  int rc;
  if ((rc = lk_operand(e, a, &oa, "lk_median_3")) != PHYLO_OK) return rc;
  if ((rc = lk_operand(e, b, &ob, "lk_median_3")) != PHYLO_OK) return rc;
  if ((rc = lk_operand(e, c, &oc, "lk_median_3")) != PHYLO_OK) return rc;
  if ((rc = ensure_pt_capacity(e, 4, e->S, e->K)) != PHYLO_OK) return rc;
  if ((rc = lk_ensure_node(e, parent)) != PHYLO_OK) return rc;
  if (!e->dSum) {
    CK(cudaMalloc(&e->dSum, sizeof(double) * (size_t)e->N * e->K * e->S));
    CK(cudaMalloc(&e->dSumSc, sizeof(int32_t) * (size_t)e->N));
  }
  e->edge_ready = false;  // the scratch CLV is the edge sum table's buffer
  CK(cudaStreamSynchronize(e->stream));  // hT is about to be rewritten
  e->hT[0] = t_a; e->hT[1] = t_b; e->hT[2] = 0.0; e->hT[3] = t_c;
  if ((rc = build_pt(e, 4)) != PHYLO_OK) return rc;
  const size_t pk = (size_t)e->K * e->S * e->S;
  if ((rc = lk_launch_prune(e, e->dP, e->dP + pk, oa, ob, e->dSum, e->dSumSc)) != PHYLO_OK) return rc;
  Operand tmp{e->dSum, e->dSumSc, false};
  if ((rc = lk_launch_prune(e, e->dP + 2 * pk, e->dP + 3 * pk, tmp, oc, e->nodes[parent].clv, e->nodes[parent].scale)) != PHYLO_OK)
    return rc;
  e->nodes[parent].valid = true;
  return PHYLO_OK;
}

// children must be tips or CLVs that exist (resident or produced earlier in the schedule)
static int lk_validate_schedule(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                const char *who) {
  std::vector<char> ready(e->cap, 0);
  for (int s = 0; s < e->cap; ++s) ready[s] = (s < e->T) || e->nodes[s].valid;
  for (int o = 0; o < n_ops; ++o) {
    const phylo_op &op = ops[o];
    if (op.parent < e->T || op.parent >= e->cap)
      return fail(e, PHYLO_ERR_ARG, "%s: op %d parent slot %d must be in [%d,%d)", who, o, op.parent, e->T, e->cap);
    if (op.left < 0 || op.left >= e->cap || op.right < 0 || op.right >= e->cap || op.left == op.parent ||
        op.right == op.parent)
      return fail(e, PHYLO_ERR_ARG, "%s: op %d has bad child slots (%d,%d)", who, o, op.left, op.right);
    if (!ready[op.left] || !ready[op.right])
      return fail(e, PHYLO_ERR_ARG, "%s: op %d uses a child that is not computed yet (not post-order)", who, o);
    ready[op.parent] = 1;
  }
  if (root_a < 0 || root_a >= e->cap || root_b < 0 || root_b >= e->cap || !ready[root_a] || !ready[root_b])
    return fail(e, PHYLO_ERR_ARG, "%s: bad root edge (%d,%d)", who, root_a, root_b);
  return PHYLO_OK;
}

extern "C" int phylo_lk_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                   double root_t, double *lnl_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_score_tree: no tips loaded");
  if (n_ops < 0 || (n_ops > 0 && !ops) || !lnl_out) return fail(e, PHYLO_ERR_ARG, "lk_score_tree: bad arguments");
  CK(cudaSetDevice(e->device));
  int rc;
  if ((rc = lk_validate_schedule(e, ops, n_ops, root_a, root_b, "lk_score_tree")) != PHYLO_OK) return rc;
  {
    bool done = false;
    if ((rc = lk_score_tree_fused(e, ops, n_ops, root_a, root_b, root_t, &done)) != PHYLO_OK) return rc;
    if (!done && (rc = lk_score_tree_fusedm(e, ops, n_ops, root_a, root_b, root_t, &done)) != PHYLO_OK) return rc;
    if (done && e->defer_scalar) {
      *lnl_out = std::nan("");
      e->lk_evaluated = true;
      return PHYLO_OK;
    }
    if (done) {
      if (!e->fused_result_ready) CK(cudaStreamSynchronize(e->stream));
      *lnl_out = e->hScalar[0];
      e->lk_evaluated = true;
      if (e->prof_on) prof_resolve_lazy(e);
      return PHYLO_OK;
    }
  }
  const int nb = 2 * n_ops + 1;
  if ((rc = ensure_pt_capacity(e, nb, e->S, e->K)) != PHYLO_OK) return rc;
  for (int o = 0; o < n_ops; ++o)
    if ((rc = lk_ensure_node(e, ops[o].parent)) != PHYLO_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  for (int o = 0; o < n_ops; ++o) {
    e->hT[2 * o] = ops[o].t_left;
    e->hT[2 * o + 1] = ops[o].t_right;
  }
  e->hT[2 * n_ops] = root_t;
  if ((rc = build_pt(e, nb)) != PHYLO_OK) return rc;
  const size_t pk = (size_t)e->K * e->S * e->S;
  for (int o = 0; o < n_ops; ++o) {
    Operand l, r;
    if ((rc = lk_operand(e, ops[o].left, &l, "lk_score_tree")) != PHYLO_OK) return rc;
    if ((rc = lk_operand(e, ops[o].right, &r, "lk_score_tree")) != PHYLO_OK) return rc;
    LkNode &p = e->nodes[ops[o].parent];
    rc = lk_launch_prune(e, e->dP + (size_t)(2 * o) * pk, e->dP + (size_t)(2 * o + 1) * pk, l, r, p.clv, p.scale);
    if (rc != PHYLO_OK) return rc;
    p.valid = true;
  }
  Operand a, b;
  if ((rc = lk_operand(e, root_a, &a, "lk_score_tree")) != PHYLO_OK) return rc;
  if ((rc = lk_operand(e, root_b, &b, "lk_score_tree")) != PHYLO_OK) return rc;
  if ((rc = lk_root_eval(e, e->dP + (size_t)(2 * n_ops) * pk, a, b, e->hScalar)) != PHYLO_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  *lnl_out = e->hScalar[0];
  e->lk_evaluated = true;
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

// 3-directional CLVs (Node.Make3D, lib/node.ml:363-477: a node keeps one value per excluded neighbour;
// readjust_3, lib/node.ml:239-256). After a down-pass, up[v] -- the CLV of "the rest of the tree" seen from
// the far end of the branch above v -- is one more pruning update per node, run parent-before-child:
//   up[v] = (P(t_s) down[s]) o (P(t_p) up[p]),  s = v's sibling, p = their parent,
// with up[a] = down[b] (and vice versa) across the root edge. The pair (down[v], up[v]) joined over the
// branch above v is the tree's likelihood seen from that edge, so every edge becomes a root edge.
extern "C" int phylo_lk_uppass(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b, double root_t,
                               const int32_t *up_slot) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_uppass: no tips loaded");
  if (n_ops < 0 || (n_ops > 0 && !ops) || !up_slot) return fail(e, PHYLO_ERR_ARG, "lk_uppass: bad arguments");
  CK(cudaSetDevice(e->device));
  int rc;
  const int cap = e->cap;
  auto in_range = [&](int s) { return s >= 0 && s < cap; };
  if (!in_range(root_a) || !in_range(root_b) || root_a == root_b) return fail(e, PHYLO_ERR_ARG, "lk_uppass: bad root edge (%d,%d)", root_a, root_b);
  std::vector<char> is_down(cap, 0), is_up(cap, 0);
  for (int s = 0; s < e->T; ++s) is_down[s] = 1;
  for (int o = 0; o < n_ops; ++o) {
    const phylo_op &op = ops[o];
    if (!in_range(op.parent) || op.parent < e->T || !in_range(op.left) || !in_range(op.right))
      return fail(e, PHYLO_ERR_ARG, "lk_uppass: op %d has bad slots", o);
    is_down[op.parent] = 1;
  }
  for (int o = 0; o < n_ops; ++o)
    for (int c : {ops[o].left, ops[o].right}) {
      const int u = up_slot[c];
      if (u < 0) continue;
      if (!in_range(u) || u < e->T || is_down[u] || is_up[u])
        return fail(e, PHYLO_ERR_ARG, "lk_uppass: up_slot[%d] = %d must be a free interior slot of its own", c, u);
      is_up[u] = 1;
    }
  // sources of every update must exist: down CLVs of the siblings, and the parent's up value
  std::vector<int> upsrc(cap, -1);      // slot holding up[p]
  std::vector<double> tabove(cap, 0.0); // length of the branch above p
  upsrc[root_a] = root_b; tabove[root_a] = root_t;
  upsrc[root_b] = root_a; tabove[root_b] = root_t;
  struct Upd { int dst, sib, src; double t_sib, t_src; };
  std::vector<Upd> upd;
  for (int o = n_ops - 1; o >= 0; --o) {  // reverse post-order = parents first
    const phylo_op &op = ops[o];
    const int p = op.parent;
    const int kids[2] = {op.left, op.right};
    const double tk[2] = {op.t_left, op.t_right};
    for (int c = 0; c < 2; ++c) {
      const int v = kids[c], sib = kids[1 - c];
      tabove[v] = tk[c];
      if (up_slot[v] < 0) continue;
      if (upsrc[p] < 0) return fail(e, PHYLO_ERR_ARG, "lk_uppass: up_slot[%d] is set but its parent %d has no up value (give it a slot, or root the pass there)", v, p);
      upd.push_back(Upd{up_slot[v], sib, upsrc[p], tk[1 - c], tabove[p]});
      upsrc[v] = up_slot[v];
    }
  }
  if (upd.empty()) return PHYLO_OK;
  const int nb = 2 * (int)upd.size();
  if ((rc = ensure_pt_capacity(e, nb, e->S, e->K)) != PHYLO_OK) return rc;
  for (const Upd &u : upd)
    if ((rc = lk_ensure_node(e, u.dst)) != PHYLO_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));  // hT is about to be rewritten
  for (size_t i = 0; i < upd.size(); ++i) {
    e->hT[2 * i] = upd[i].t_sib;
    e->hT[2 * i + 1] = upd[i].t_src;
  }
  if ((rc = build_pt(e, nb)) != PHYLO_OK) return rc;
  const size_t pk = (size_t)e->K * e->S * e->S;
  // 4 states: the updates of one tree level are independent (each needs only its parent's up value), so a level
  // is ONE launch over (update, pattern range) items -- 2T - 4 launches become one per level. The values are the
  // per-node kernels' bit for bit (same body). PHYLO_UPPASS_BATCH=0 keeps one launch per update (cross-check).
  const char *sw = std::getenv("PHYLO_UPPASS_BATCH");
  if (e->S == 4 && (e->K == 1 || e->K == 2 || e->K == 4 || e->K == 8 || e->K == 16) && !(sw && sw[0] == '0')) {
    std::vector<int> lvl_of_slot(cap, 0), lvl(upd.size());
    int n_lvl = 0;
    for (size_t i = 0; i < upd.size(); ++i) {  // parents first: the level of the source is known
      lvl[i] = lvl_of_slot[upd[i].src] + (upd[i].src == root_a || upd[i].src == root_b ? 0 : 1);
      lvl_of_slot[upd[i].dst] = lvl[i];
      n_lvl = std::max(n_lvl, lvl[i] + 1);
    }
    auto undo = [&](int code) {  // nothing was launched for these slots, or the launch failed
      for (const Upd &u : upd) e->nodes[u.dst].valid = false;
      return code;
    };
    std::vector<std::vector<int>> by_lvl(n_lvl);
    for (size_t i = 0; i < upd.size(); ++i) by_lvl[lvl[i]].push_back((int)i);
    std::vector<PruneItem> items;
    items.reserve(upd.size());
    struct Launch { int first, count, tt_update; };  // tt_update >= 0: a tip + tip update through its own kernel
    std::vector<Launch> launches;
    for (int L = 0; L < n_lvl; ++L) {
      const int first = (int)items.size();
      for (int i : by_lvl[L]) {
        Operand l, r;
        if ((rc = lk_operand(e, upd[i].sib, &l, "lk_uppass")) != PHYLO_OK) return undo(rc);
        if ((rc = lk_operand(e, upd[i].src, &r, "lk_uppass")) != PHYLO_OK) return undo(rc);
        LkNode &dst = e->nodes[upd[i].dst];
        dst.valid = true;  // (a source of the next level; cleared again below if the call fails)
        if (l.tip && r.tip) { launches.push_back(Launch{0, 0, i}); continue; }
        items.push_back(PruneItem{e->dP + (size_t)(2 * i) * pk, e->dP + (size_t)(2 * i + 1) * pk, l.src, r.src, l.scale, r.scale,
                                  dst.clv, dst.scale, l.tip ? 1 : 0, r.tip ? 1 : 0});
      }
      if ((int)items.size() > first) launches.push_back(Launch{first, (int)items.size() - first, -1});
    }
    const size_t bytes = (sizeof(PruneItem) * std::max<size_t>(1, items.size()) + 255) & ~(size_t)255;
    if (bytes > e->capEdgeArena) {
      dfree(e->dEdgeArena);
      e->capEdgeArena = 0;
      if (cudaMalloc(&e->dEdgeArena, bytes) != cudaSuccess) {
        cudaGetLastError();
        return undo(fail(e, PHYLO_ERR_CUDA, "lk_uppass: cannot allocate %zu bytes of device memory", bytes));
      }
      e->capEdgeArena = bytes;
    }
    PruneItem *dItems = (PruneItem *)e->dEdgeArena;
    if (!items.empty()) CK(cudaMemcpyAsync(dItems, items.data(), sizeof(PruneItem) * items.size(), cudaMemcpyHostToDevice, e->stream));
    const int64_t per_node = (e->N * e->K + 256 * 2 - 1) / (256 * 2);  // CTAs that cover a node at U = 2
    for (const Launch &la : launches) {
      if (la.tt_update >= 0) {
        const int i = la.tt_update;
        Operand l, r;
        if ((rc = lk_operand(e, upd[i].sib, &l, "lk_uppass")) != PHYLO_OK) return rc;
        if ((rc = lk_operand(e, upd[i].src, &r, "lk_uppass")) != PHYLO_OK) return rc;
        LkNode &dst = e->nodes[upd[i].dst];
        if ((rc = lk_launch_prune(e, e->dP + (size_t)(2 * i) * pk, e->dP + (size_t)(2 * i + 1) * pk, l, r, dst.clv, dst.scale)) != PHYLO_OK) return rc;
        continue;
      }
      // about 8 CTAs per SM in flight over the whole level, at least one and at most `per_node` per update
      const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(per_node, ((int64_t)e->sm_count * 8 + la.count - 1) / la.count));
      for (int y0 = 0; y0 < la.count; y0 += 65535) {
        const int ny = std::min(65535, la.count - y0);
        ProfScope prof(e, KC_PRUNE_II);
        const dim3 grid((unsigned)gx, (unsigned)ny);
        switch (e->K) {
          case 1: prune4_level_kernel<1><<<grid, 256, 0, e->stream>>>(dItems + la.first + y0, e->N); break;
          case 2: prune4_level_kernel<2><<<grid, 256, 0, e->stream>>>(dItems + la.first + y0, e->N); break;
          case 4: prune4_level_kernel<4><<<grid, 256, 0, e->stream>>>(dItems + la.first + y0, e->N); break;
          case 8: prune4_level_kernel<8><<<grid, 256, 0, e->stream>>>(dItems + la.first + y0, e->N); break;
          default: prune4_level_kernel<16><<<grid, 256, 0, e->stream>>>(dItems + la.first + y0, e->N);
        }
        LAUNCH_CHECK();
      }
    }
    e->edge_ready = false;
    if (e->prof_on) prof_resolve_lazy(e);
    return PHYLO_OK;
  }
  for (size_t i = 0; i < upd.size(); ++i) {
    Operand l, r;
    if ((rc = lk_operand(e, upd[i].sib, &l, "lk_uppass")) != PHYLO_OK) return rc;
    if ((rc = lk_operand(e, upd[i].src, &r, "lk_uppass")) != PHYLO_OK) return rc;
    LkNode &dst = e->nodes[upd[i].dst];
    rc = lk_launch_prune(e, e->dP + (size_t)(2 * i) * pk, e->dP + (size_t)(2 * i + 1) * pk, l, r, dst.clv, dst.scale);
    if (rc != PHYLO_OK) return rc;
    dst.valid = true;
  }
  e->edge_ready = false;
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

// ---- Sankoff: cost-vector parsimony under a general cost matrix (sankoff_kernels.cuh)
extern "C" int phylo_sankoff_set_matrix(phylo_engine *e, int n_states, const int32_t *M) {
  if (!e) return PHYLO_ERR_ARG;
  if (n_states < 2 || n_states > 32 || !M) return fail(e, PHYLO_ERR_ARG, "sankoff_set_matrix: 2 <= n_states <= 32 and a matrix");
  for (int i = 0; i < n_states * n_states; ++i)
    if (M[i] < 0 || M[i] > 1000000) return fail(e, PHYLO_ERR_DATA, "sankoff_set_matrix: costs must be in [0, 1000000]");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  if (e->skT && n_states != e->skS) sankoff_free_data(e);  // another alphabet: the loaded characters go
  dfree(e->dSkM);
  CK(cudaMalloc(&e->dSkM, sizeof(int) * n_states * n_states));
  CK(cudaMemcpy(e->dSkM, M, sizeof(int) * n_states * n_states, cudaMemcpyHostToDevice));
  e->skS = n_states;
  e->skHasM = true;
  for (auto &v : e->skValid) v = 0;
  return PHYLO_OK;
}

extern "C" int phylo_sankoff_set_tips(phylo_engine *e, int T, int64_t N, int elt_bytes, int n_states, const void *codes,
                                      const double *weights, int capacity) {
  if (!e) return PHYLO_ERR_ARG;
  if (T < 2 || N < 1 || !codes || capacity < T || n_states < 2 || n_states > 32 ||
      !(elt_bytes == 1 || elt_bytes == 2 || elt_bytes == 4) || n_states > 8 * elt_bytes)
    return fail(e, PHYLO_ERR_ARG, "sankoff_set_tips: bad shape (2 <= n_states <= min(32, 8 elt_bytes), elt_bytes in {1,2,4})");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  sankoff_free_data(e);
  if (e->skHasM && n_states != e->skS) { dfree(e->dSkM); e->skHasM = false; }  // another alphabet: the matrix goes
  const uint32_t keep = n_states >= 32 ? 0xffffffffu : ((1u << n_states) - 1u);
  std::vector<uint32_t> h((size_t)T * N);
  for (size_t i = 0; i < h.size(); ++i) {
    const uint32_t v = elt_bytes == 1 ? ((const uint8_t *)codes)[i] : (elt_bytes == 2 ? ((const uint16_t *)codes)[i] : ((const uint32_t *)codes)[i]);
    if ((v & keep) == 0) return fail(e, PHYLO_ERR_DATA, "sankoff_set_tips: taxon %lld character %lld has an empty state set", (long long)(i / N), (long long)(i % N));
    h[i] = v & keep;
  }
  CK(cudaMalloc(&e->dSkTips, sizeof(uint32_t) * h.size()));
  CK(cudaMemcpy(e->dSkTips, h.data(), sizeof(uint32_t) * h.size(), cudaMemcpyHostToDevice));
  if (weights) {
    std::vector<uint32_t> hw(N);
    for (int64_t i = 0; i < N; ++i) {
      if (!(weights[i] >= 0.0) || weights[i] != std::floor(weights[i]) || weights[i] > 4e9)
        return fail(e, PHYLO_ERR_DATA, "sankoff_set_tips: weights must be non-negative integers");
      hw[i] = (uint32_t)weights[i];
    }
    CK(cudaMalloc(&e->dSkW, sizeof(uint32_t) * N));
    CK(cudaMemcpy(e->dSkW, hw.data(), sizeof(uint32_t) * N, cudaMemcpyHostToDevice));
  }
  if (!e->dSkTotal) CK(cudaMalloc(&e->dSkTotal, sizeof(unsigned long long)));
  e->skVec.assign(capacity, nullptr);
  e->skValid.assign(capacity, 0);
  CK(cudaMalloc(&e->dSkTab, sizeof(int *) * capacity));
  e->skTabDirty = true;
  e->skT = T; e->skN = N; e->skCap = capacity; e->skS = n_states;
  return PHYLO_OK;
}

static int sk_ready(phylo_engine *e, const char *who) {
  if (!e->skT) return fail(e, PHYLO_ERR_STATE, "%s: no characters loaded (phylo_sankoff_set_tips)", who);
  if (!e->skHasM) return fail(e, PHYLO_ERR_STATE, "%s: no cost matrix (phylo_sankoff_set_matrix)", who);
  return PHYLO_OK;
}
static int sk_ensure(phylo_engine *e, int slot) {
  if (e->skVec[slot]) return PHYLO_OK;
  CK(cudaMalloc(&e->skVec[slot], sizeof(int) * (size_t)e->skS * e->skN));
  e->skTabDirty = true;
  return PHYLO_OK;
}
static int sk_st(int S) { return S <= 4 ? 4 : (S <= 8 ? 8 : (S <= 16 ? 16 : (S <= 20 ? 20 : 32))); }

#define SK_DISPATCH(st, EXPR)                        \
  switch (st) {                                      \
    case 4: { constexpr int ST = 4; EXPR; } break;   \
    case 8: { constexpr int ST = 8; EXPR; } break;   \
    case 16: { constexpr int ST = 16; EXPR; } break; \
    case 20: { constexpr int ST = 20; EXPR; } break; \
    default: { constexpr int ST = 32; EXPR; } break; \
  }

// parent >= 0: median into the slot; parent < 0: root-edge join. *total += the characters' weighted minimum / join
static int sk_launch_node(phylo_engine *e, int parent, int left, int right, const char *who) {
  auto operand = [&](int s, const uint32_t **tip, const int **vec) -> int {
    *tip = nullptr; *vec = nullptr;
    if (s < 0 || s >= e->skCap) return fail(e, PHYLO_ERR_ARG, "%s: node slot %d out of range", who, s);
    if (s < e->skT) { *tip = e->dSkTips + (size_t)s * e->skN; return PHYLO_OK; }
    if (!e->skValid[s]) return fail(e, PHYLO_ERR_STATE, "%s: node slot %d has no cost vectors yet", who, s);
    *vec = e->skVec[s];
    return PHYLO_OK;
  };
  const uint32_t *lt, *rt;
  const int *lv, *rv;
  int rc;
  if ((rc = operand(left, &lt, &lv)) != PHYLO_OK) return rc;
  if ((rc = operand(right, &rt, &rv)) != PHYLO_OK) return rc;
  int *out = nullptr;
  if (parent >= 0) {
    if (parent < e->skT || parent >= e->skCap || parent == left || parent == right)
      return fail(e, PHYLO_ERR_ARG, "%s: bad parent slot %d", who, parent);
    if ((rc = sk_ensure(e, parent)) != PHYLO_OK) return rc;
    out = e->skVec[parent];
  }
  const int g = grid_for(e->skN, 128, e->sm_count * 8);
  ProfScope prof(e, KC_FITCH_NODE);
  SK_DISPATCH(sk_st(e->skS), (sankoff_node_kernel<ST><<<g, 128, 0, e->stream>>>(e->dSkM, e->skS, e->skN, lt, lv, rt, rv, out, e->dSkW, e->dSkTotal)));
  LAUNCH_CHECK();
  if (parent >= 0) e->skValid[parent] = 1;
  return PHYLO_OK;
}

static int sk_read_total(phylo_engine *e, uint64_t *out) {
  CK(cudaMemcpyAsync(e->hScalar + HS_BAD, e->dSkTotal, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  std::memcpy(out, e->hScalar + HS_BAD, 8);
  return PHYLO_OK;
}

extern "C" int phylo_sankoff_median_2(phylo_engine *e, int parent, int left, int right, uint64_t *subtree_cost_out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = sk_ready(e, "sankoff_median_2")) != PHYLO_OK) return rc;
  if (parent < 0) return fail(e, PHYLO_ERR_ARG, "sankoff_median_2: bad parent slot %d", parent);
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(e->dSkTotal, 0, sizeof(unsigned long long), e->stream));
  if ((rc = sk_launch_node(e, parent, left, right, "sankoff_median_2")) != PHYLO_OK) return rc;
  if (subtree_cost_out) return sk_read_total(e, subtree_cost_out);
  return PHYLO_OK;
}

extern "C" int phylo_sankoff_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                        uint64_t *length_out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = sk_ready(e, "sankoff_score_tree")) != PHYLO_OK) return rc;
  if (n_ops < 0 || (n_ops > 0 && !ops) || !length_out) return fail(e, PHYLO_ERR_ARG, "sankoff_score_tree: bad arguments");
  CK(cudaSetDevice(e->device));
  std::vector<char> ready(e->skCap, 0);
  for (int s = 0; s < e->skCap; ++s) ready[s] = s < e->skT || e->skValid[s];
  for (int o = 0; o < n_ops; ++o) {
    const phylo_op &op = ops[o];
    if (op.parent < e->skT || op.parent >= e->skCap || op.left < 0 || op.left >= e->skCap || op.right < 0 || op.right >= e->skCap ||
        op.left == op.parent || op.right == op.parent)
      return fail(e, PHYLO_ERR_ARG, "sankoff_score_tree: op %d has bad slots", o);
    if (!ready[op.left] || !ready[op.right]) return fail(e, PHYLO_ERR_ARG, "sankoff_score_tree: op %d uses a child that is not computed yet", o);
    ready[op.parent] = 1;
  }
  if (root_a < 0 || root_a >= e->skCap || root_b < 0 || root_b >= e->skCap || !ready[root_a] || !ready[root_b])
    return fail(e, PHYLO_ERR_ARG, "sankoff_score_tree: bad root edge (%d,%d)", root_a, root_b);
  CK(cudaMemsetAsync(e->dSkTotal, 0, sizeof(unsigned long long), e->stream));
  FusedPlan pl;
  const bool fused = e->opt_fused && n_ops > 0 && build_fused_plan(e->skCap, e->skT, ops, n_ops, root_a, root_b, 0.0, pl) &&
                     pl.depth <= kSkMaxDepth;
  if (fused) {
    if (e->opt_retain)
      for (int o = 0; o < n_ops; ++o)
        if ((rc = sk_ensure(e, ops[o].parent)) != PHYLO_OK) return rc;
    const size_t pbytes = sizeof(SkInstr) * pl.steps.size();
    if (pbytes > e->capProg) {
      CK(cudaStreamSynchronize(e->stream));
      dfree(e->dProg);
      if (e->hProg) { cudaFreeHost(e->hProg); e->hProg = nullptr; }
      e->capProg = 0;
      CK(cudaMalloc(&e->dProg, pbytes * 2));
      CK(cudaMallocHost(&e->hProg, pbytes * 2));
      e->capProg = pbytes * 2;
    }
    CK(cudaStreamSynchronize(e->stream));
    if (e->skTabDirty) {
      CK(cudaMemcpy(e->dSkTab, e->skVec.data(), sizeof(int *) * e->skCap, cudaMemcpyHostToDevice));
      e->skTabDirty = false;
    }
    SkInstr *hp = (SkInstr *)e->hProg;
    for (size_t i = 0; i < pl.steps.size(); ++i) {
      const PlanStep &st = pl.steps[i];
      hp[i] = SkInstr{st.lkind | (st.rkind << 2) | (st.push_first << 4), st.lidx, st.ridx, e->opt_retain ? st.out_slot : -1};
    }
    CK(cudaMemcpyAsync(e->dProg, e->hProg, pbytes, cudaMemcpyHostToDevice, e->stream));
    const int g = grid_for(e->skN, 128, e->sm_count * 8);
    {
      ProfScope prof(e, KC_FITCH_TREE);
      if (e->opt_retain) {
        SK_DISPATCH(sk_st(e->skS), (sankoff_tree_kernel<ST, true><<<g, 128, 0, e->stream>>>(
                                       e->dSkM, e->skS, e->skN, e->dSkTips, e->skN, (const SkInstr *)e->dProg, n_ops, e->dSkTab, e->dSkW, e->dSkTotal)));
      } else {
        SK_DISPATCH(sk_st(e->skS), (sankoff_tree_kernel<ST, false><<<g, 128, 0, e->stream>>>(
                                       e->dSkM, e->skS, e->skN, e->dSkTips, e->skN, (const SkInstr *)e->dProg, n_ops, e->dSkTab, e->dSkW, e->dSkTotal)));
      }
      LAUNCH_CHECK();
    }
    for (int o = 0; o < n_ops; ++o) e->skValid[ops[o].parent] = e->opt_retain ? 1 : 0;
  } else {
    for (int o = 0; o < n_ops; ++o) {
      // the subtree minima of interior medians are not part of the length: only the root join is summed
      if ((rc = sk_launch_node(e, ops[o].parent, ops[o].left, ops[o].right, "sankoff_score_tree")) != PHYLO_OK) return rc;
    }
    CK(cudaMemsetAsync(e->dSkTotal, 0, sizeof(unsigned long long), e->stream));
    if ((rc = sk_launch_node(e, -1, root_a, root_b, "sankoff_score_tree")) != PHYLO_OK) return rc;
  }
  rc = sk_read_total(e, length_out);
  if (e->prof_on) prof_resolve_lazy(e);
  return rc;
}

extern "C" int phylo_sankoff_get_costs(phylo_engine *e, int node, int32_t *out) {
  if (!e || !out) return PHYLO_ERR_ARG;
  if (!e->skT) return fail(e, PHYLO_ERR_STATE, "sankoff_get_costs: no characters loaded");
  if (node < e->skT || node >= e->skCap || !e->skValid[node]) return fail(e, PHYLO_ERR_STATE, "sankoff_get_costs: node slot %d has no cost vectors", node);
  CK(cudaSetDevice(e->device));
  std::vector<int32_t> planes((size_t)e->skS * e->skN);
  CK(cudaMemcpyAsync(planes.data(), e->skVec[node], sizeof(int) * planes.size(), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (int64_t n = 0; n < e->skN; ++n)
    for (int s = 0; s < e->skS; ++s) out[(size_t)n * e->skS + s] = planes[(size_t)s * e->skN + n];
  return PHYLO_OK;
}

// ---- model-parameter derivatives (lk_grad_kernels.cuh)
// matrices of one branch for the gradient kernel: P_k = exp(Q t r_k), then for every parameter p
// d/d theta_p exp(Q t r_k) = dexp_X[D], X = Q t r_k, D = t (r_k dQ_p + drates_p[k] Q)
struct GradHost {  // per-call constants of grad_branch_matrices: G_p = Vinv dQ_p V and scratch
  std::vector<double> Gp, W, mid, x, ex;
  GradHost(const phylo_engine *e, int np, const double *dQ) {
    const int S = e->S;
    const size_t ss = (size_t)S * S;
    const double *V = e->hV.data(), *Vi = e->hVinv.data();
    Gp.assign((size_t)np * ss, 0.0);
    W.resize(ss); mid.resize(ss); x.resize(S); ex.resize(S);
    for (int p = 0; p < np && dQ; ++p) {
      const double *dq = dQ + (size_t)p * ss;
      for (int m = 0; m < S; ++m)
        for (int j = 0; j < S; ++j) {
          double a = 0.0;
          for (int i = 0; i < S; ++i) a += Vi[(size_t)m * S + i] * dq[(size_t)i * S + j];
          W[(size_t)m * S + j] = a;
        }
      for (int m = 0; m < S; ++m)
        for (int n = 0; n < S; ++n) {
          double a = 0.0;
          for (int j = 0; j < S; ++j) a += W[(size_t)m * S + j] * V[(size_t)j * S + n];
          Gp[(size_t)p * ss + (size_t)m * S + n] = a;
        }
    }
  }
};

template <int ST>  // ST = 4: loops unrolled for nucleotides (510 branches x 28 products were 2 ms of the call); 0 = run-time S
static void grad_branch_matrices_t(const phylo_engine *e, GradHost &h, double t, int np, const double *drates,
                                   double *out /* [(np+1)][K][S][S] */) {
  const int S = ST ? ST : e->S, K = e->K;
  const size_t ss = (size_t)S * S;
  const double *V = e->hV.data(), *Vi = e->hVinv.data(), *lam = e->hLam.data();
  double *W = h.W.data(), *mid = h.mid.data(), *x = h.x.data(), *ex = h.ex.data();
  const double *Gp = h.Gp.data();
  auto sandwich = [&](const double *md, double *dst) {  // dst = V md Vinv
    for (int i = 0; i < S; ++i)
      for (int n = 0; n < S; ++n) {
        double a = 0.0;
        for (int m = 0; m < S; ++m) a += V[(size_t)i * S + m] * md[(size_t)m * S + n];
        W[(size_t)i * S + n] = a;
      }
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        double a = 0.0;
        for (int n = 0; n < S; ++n) a += W[(size_t)i * S + n] * Vi[(size_t)n * S + j];
        dst[(size_t)i * S + j] = a;
      }
  };
  for (int k = 0; k < K; ++k) {
    const double tau = t * e->hRates[k];
    for (int m = 0; m < S; ++m) { x[m] = lam[m] * tau; ex[m] = std::exp(x[m]); }
    std::fill(mid, mid + ss, 0.0);
    for (int m = 0; m < S; ++m) mid[(size_t)m * S + m] = ex[m];
    sandwich(mid, out + (size_t)k * ss);
    for (int p = 0; p < np; ++p) {
      const double dr = drates ? drates[(size_t)p * K + k] : 0.0;
      for (int m = 0; m < S; ++m)
        for (int n = 0; n < S; ++n) {
          // direction in eigen-space: t (r_k G_p + dr Lambda); divided difference Phi_mn of exp
          double d = tau * Gp[(size_t)p * ss + (size_t)m * S + n];
          if (m == n) d += t * dr * lam[m];
          const double dx = x[m] - x[n];
          const double phi = (m == n || std::fabs(dx) < 1e-9) ? 0.5 * (ex[m] + ex[n]) : (ex[m] - ex[n]) / dx;
          mid[(size_t)m * S + n] = phi * d;
        }
      sandwich(mid, out + ((size_t)(1 + p) * K + k) * ss);
    }
  }
}

static void grad_branch_matrices(const phylo_engine *e, GradHost &h, double t, int np, const double *drates, double *out) {
  if (e->S == 4) grad_branch_matrices_t<4>(e, h, t, np, drates, out);
  else grad_branch_matrices_t<0>(e, h, t, np, drates, out);
}

// A fragments of param_grad4_mma_kernel for parameters q0 .. q0 + nq - 1 of every branch: [branch][K][4][32], lane
// (row = lane / 4, c = lane % 4) of k-step (k, s) holds row's coefficient of a_k[s] b_k[c]; row 0 = the likelihood,
// row 1 + q = parameter q0 + q (the prior's derivative enters at branch 0 = the root edge), other rows zero.
static void grad_fill_afrag(const phylo_engine *e, const std::vector<double> &hm, size_t per_branch, int n_edges, int q0,
                            int nq, const double *dpi, double *out) {
  const int K = e->K;
  for (int i = 0; i < n_edges; ++i) {
    const double *m = hm.data() + (size_t)i * per_branch;
    for (int k = 0; k < K; ++k)
      for (int s = 0; s < 4; ++s)
        for (int lane = 0; lane < 32; ++lane) {
          const int row = lane >> 2, c = lane & 3;
          const double P = m[(size_t)k * 16 + s * 4 + c];
          double v = 0.0;
          if (row == 0) {
            v = e->hProbs[k] * (e->hPi[s] * P);
          } else if (row <= nq) {
            const int q = q0 + row - 1;
            v = e->hPi[s] * m[((size_t)(1 + q) * K + k) * 16 + s * 4 + c];
            if (i == 0 && dpi) v += dpi[(size_t)q * 4 + s] * P;
            v *= e->hProbs[k];
          }
          out[(((size_t)i * K + k) * 4 + s) * 32 + lane] = v;
        }
  }
}

template <typename GetOperands>
static int grad4_mma_path(phylo_engine *e, const std::vector<double> &hm, size_t per_branch, int n_edges, int n_params,
                          const double *dpi, double *lnl_out, double *grad_out, GetOperands operands) {
  const int K = e->K;
  const int64_t nb = e->nPart;
  std::vector<EdgeJoin> he(n_edges);
  for (int i = 0; i < n_edges; ++i) {
    Operand a, b;
    const int rc = operands(i, &a, &b);
    if (rc != PHYLO_OK) return rc;
    he[i] = EdgeJoin{a.src, b.src, a.scale, b.scale, a.tip ? 1 : 0, b.tip ? 1 : 0};
  }
  constexpr int kRows = 7;  // parameters per pass (row 0 of the 8-row tile is the likelihood)
  const int nq_max = std::min(n_params, kRows);
  // branches per launch: the per-branch block results stay below 256 MB (PHYLO_GRAD_CHUNK_BYTES: another budget, so
  // that a test can drive the several-launches path on a small tree)
  int64_t budget = (int64_t)256 << 20;
  if (const char *v = std::getenv("PHYLO_GRAD_CHUNK_BYTES")) budget = std::max<int64_t>(1, std::atoll(v));
  const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(n_edges, budget / (8 * nq_max * nb)));
  auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t descBytes = up256(sizeof(EdgeJoin) * n_edges), fragBytes = up256(sizeof(double) * (size_t)n_edges * K * 128),
               gpBytes = up256(sizeof(double) * (size_t)chunk * nq_max * nb), outBytes = up256(sizeof(double) * (size_t)n_params * nb);
  const size_t bytes = descBytes + fragBytes + gpBytes + outBytes;
  if (bytes > e->capEdgeArena) {
    dfree(e->dEdgeArena);
    e->capEdgeArena = 0;
    if (cudaMalloc(&e->dEdgeArena, bytes) != cudaSuccess) {
      cudaGetLastError();
      return fail(e, PHYLO_ERR_CUDA, "lk_param_gradient: cannot allocate %zu bytes of device memory", bytes);
    }
    e->capEdgeArena = bytes;
  }
  char *arena = e->dEdgeArena;
  EdgeJoin *dE = (EdgeJoin *)arena;
  double *dA = (double *)(arena + descBytes), *dGp = (double *)(arena + descBytes + fragBytes),
         *dOut = (double *)(arena + descBytes + fragBytes + gpBytes);
  cudaMemcpyAsync(dE, he.data(), sizeof(EdgeJoin) * n_edges, cudaMemcpyHostToDevice, e->stream);
  cudaMemsetAsync(dOut, 0, sizeof(double) * (size_t)n_params * nb, e->stream);
  std::vector<double> ha((size_t)n_edges * K * 128);
  for (int q0 = 0; q0 < n_params; q0 += kRows) {
    const int nq = std::min(kRows, n_params - q0);
    if (q0 > 0) CK(cudaStreamSynchronize(e->stream));  // `ha` is reused by the next pass
    grad_fill_afrag(e, hm, per_branch, n_edges, q0, nq, dpi, ha.data());
    cudaMemcpyAsync(dA, ha.data(), sizeof(double) * ha.size(), cudaMemcpyHostToDevice, e->stream);
    for (int e0 = 0; e0 < n_edges; e0 += chunk) {
      const int ne = std::min(chunk, n_edges - e0);
      {
        ProfScope prof(e, KC_EDGE);
        const int g = (int)std::min<int64_t>(nb * ne, (int64_t)e->sm_count * 16);
#define GM(KV)                                                                                                            \
  param_grad4_mma_kernel<KV><<<g, 256, 0, e->stream>>>(dE + e0, ne, dA + (size_t)e0 * KV * 128, nq, e->dPi, e->pinvar,    \
                                                       (const uint8_t *)e->dInv, e->dWeights, dGp, e->N);
        switch (K) {
          case 1: GM(1) break;
          case 2: GM(2) break;
          case 4: GM(4) break;
          default: GM(8)
        }
#undef GM
        ++e->launches;
      }
      {
        ProfScope prof(e, KC_REDUCE);
        const int64_t cols = (int64_t)nq * nb;
        grad_sum_branches_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, e->stream>>>(dGp, ne, cols, dOut + (size_t)q0 * nb);
        ++e->launches;
      }
    }
  }
  std::vector<double> hg((size_t)n_params * nb);
  const cudaError_t st = cudaMemcpyAsync(hg.data(), dOut, sizeof(double) * hg.size(), cudaMemcpyDeviceToHost, e->stream);
  const cudaError_t st2 = cudaStreamSynchronize(e->stream);
  const cudaError_t le = cudaGetLastError();
  if (st != cudaSuccess || st2 != cudaSuccess || le != cudaSuccess)
    return fail(e, PHYLO_ERR_CUDA, "lk_param_gradient: %s", cudaGetErrorString(st != cudaSuccess ? st : st2 != cudaSuccess ? st2 : le));
  for (int p = 0; p < n_params; ++p) grad_out[p] = phylo_reduce_partials(hg.data() + (size_t)p * nb, nb);
  if (lnl_out) *lnl_out = e->hScalar[0];
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

extern "C" int phylo_lk_param_gradient(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                       double root_t, const int32_t *up_slot, int n_params, const double *dQ,
                                       const double *drates, const double *dpi, double *lnl_out, double *grad_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0 || !e->has_model) return fail(e, PHYLO_ERR_STATE, "lk_param_gradient: model and tips first");
  if (n_ops < 0 || (n_ops > 0 && !ops) || !up_slot || n_params < 1 || n_params > 64 || !grad_out || (!dQ && !drates && !dpi))
    return fail(e, PHYLO_ERR_ARG, "lk_param_gradient: bad arguments (1..64 parameters, at least one of dQ / drates / dpi)");
  if (e->S > 64) return fail(e, PHYLO_ERR_UNSUPPORTED, "lk_param_gradient: at most 64 states");
  if (dpi && e->pinvar >= 0.0)
    return fail(e, PHYLO_ERR_UNSUPPORTED, "lk_param_gradient: prior derivatives together with an invariant-sites class");
  CK(cudaSetDevice(e->device));
  int rc;
  const int S = e->S, K = e->K;
  const size_t ss = (size_t)S * S, per_branch = (size_t)(n_params + 1) * K * ss;
  struct Edge { int a, b; double t; bool root; };
  std::vector<Edge> edges;
  edges.push_back(Edge{root_a, root_b, root_t, true});
  for (int o = 0; o < n_ops; ++o) {
    const int kids[2] = {ops[o].left, ops[o].right};
    const double tk[2] = {ops[o].t_left, ops[o].t_right};
    for (int c = 0; c < 2; ++c) {
      const int v = kids[c];
      if (v < 0 || v >= e->cap) return fail(e, PHYLO_ERR_ARG, "lk_param_gradient: op %d has bad slots", o);
      if (up_slot[v] < 0) return fail(e, PHYLO_ERR_ARG, "lk_param_gradient: node %d has no up value (run phylo_lk_uppass with every slot set)", v);
      edges.push_back(Edge{up_slot[v], v, tk[c], false});  // a = rest of the tree, b = the subtree below the branch
    }
  }
  std::vector<double> hm(per_branch * edges.size());
  {
    GradHost gh(e, n_params, dQ);  // (G_p once per call, not per branch)
    for (size_t i = 0; i < edges.size(); ++i) grad_branch_matrices(e, gh, edges[i].t, n_params, drates, hm.data() + i * per_branch);
  }
  // 4 states: all branches in one launch on the fp64 tensor cores (param_grad4_mma_kernel), seven parameters per
  // pass; PHYLO_GRAD_MMA=0 keeps the scalar per-branch kernel below (measurement / cross-check only)
  {
    const char *sw = std::getenv("PHYLO_GRAD_MMA");
    if (S == 4 && (K == 1 || K == 2 || K == 4 || K == 8) && e->mask_dev_bytes == 1 && !(sw && sw[0] == '0'))
      return grad4_mma_path(e, hm, per_branch, (int)edges.size(), n_params, dpi, lnl_out, grad_out,
                            [&](int i, Operand *a, Operand *b) {
                              int r = lk_operand(e, edges[i].a, a, "lk_param_gradient");
                              return r != PHYLO_OK ? r : lk_operand(e, edges[i].b, b, "lk_param_gradient");
                            });
  }
  double *dM = nullptr, *dG = nullptr, *dDpi = nullptr;
  const size_t gdoubles = (size_t)n_params * e->nPart;
  auto cleanup = [&] { cudaFree(dM); cudaFree(dG); cudaFree(dDpi); };
  if (cudaMalloc(&dM, sizeof(double) * hm.size()) != cudaSuccess || cudaMalloc(&dG, sizeof(double) * gdoubles) != cudaSuccess ||
      (dpi && cudaMalloc(&dDpi, sizeof(double) * n_params * S) != cudaSuccess)) {
    cleanup();
    return fail(e, PHYLO_ERR_CUDA, "lk_param_gradient: out of device memory");
  }
  cudaMemcpyAsync(dM, hm.data(), sizeof(double) * hm.size(), cudaMemcpyHostToDevice, e->stream);
  cudaMemsetAsync(dG, 0, sizeof(double) * gdoubles, e->stream);
  if (dpi) cudaMemcpyAsync(dDpi, dpi, sizeof(double) * n_params * S, cudaMemcpyHostToDevice, e->stream);
  // parameters per pass: P + nq derivative matrices must fit shared memory, nq <= 8
  const size_t fixed = sizeof(double) * ((size_t)K * ss + S + kLnlBlock + 32);
  int nq_max = 8;
  while (nq_max > 1 && fixed + sizeof(double) * (size_t)nq_max * (K * ss + S) > 200 * 1024) --nq_max;
  if (fixed + sizeof(double) * (K * ss + S) > 200 * 1024) { cleanup(); return fail(e, PHYLO_ERR_UNSUPPORTED, "lk_param_gradient: matrices exceed shared memory"); }
  for (size_t i = 0; i < edges.size(); ++i) {
    Operand a, b;
    if ((rc = lk_operand(e, edges[i].a, &a, "lk_param_gradient")) != PHYLO_OK) { cleanup(); return rc; }
    if ((rc = lk_operand(e, edges[i].b, &b, "lk_param_gradient")) != PHYLO_OK) { cleanup(); return rc; }
    for (int q0 = 0; q0 < n_params; q0 += nq_max) {
      const int nq = std::min(nq_max, n_params - q0);
      const size_t smem = fixed + sizeof(double) * (size_t)nq * (K * ss + S);
      const double *dpi_edge = edges[i].root ? dDpi : nullptr;  // the prior enters once, at the root edge
      ProfScope prof(e, KC_EDGE);
#define GRAD_LAUNCH(ST_, M_)                                                                                              \
  {                                                                                                                       \
    auto kern = param_grad_kernel<ST_, M_>;                                                                               \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                   \
    kern<<<(int)e->nPart, 256, smem, e->stream>>>(dM + i * per_branch, n_params, q0, nq, e->dPi, dpi_edge, e->dProbs,     \
                                                  e->pinvar, (const M_ *)e->dInv, a.src, a.scale, a.tip, b.src, b.scale,   \
                                                  b.tip, e->dWeights, dG, e->nPart, e->N, S, K);                          \
  }
      if (S == 4) GRAD_LAUNCH(4, uint8_t)
      else if (e->mask_dev_bytes == 1) GRAD_LAUNCH(0, uint8_t)
      else if (e->mask_dev_bytes == 4) GRAD_LAUNCH(0, uint32_t)
      else GRAD_LAUNCH(0, uint64_t)
#undef GRAD_LAUNCH
      ++e->launches;
      const cudaError_t st = cudaGetLastError();
      if (st != cudaSuccess) { cleanup(); return fail(e, PHYLO_ERR_CUDA, "param_grad launch: %s", cudaGetErrorString(st)); }
    }
  }
  std::vector<double> hg(gdoubles);
  const cudaError_t st = cudaMemcpyAsync(hg.data(), dG, sizeof(double) * gdoubles, cudaMemcpyDeviceToHost, e->stream);
  const cudaError_t st2 = cudaStreamSynchronize(e->stream);
  cleanup();
  if (st != cudaSuccess || st2 != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "lk_param_gradient: %s", cudaGetErrorString(st != cudaSuccess ? st : st2));
  for (int p = 0; p < n_params; ++p) grad_out[p] = phylo_reduce_partials(hg.data() + (size_t)p * e->nPart, e->nPart);
  if (lnl_out) *lnl_out = e->hScalar[0];
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

// ---- device-side scalar exchange (exchange_kernels.cuh)
extern "C" int phylo_exchange_alloc(phylo_engine *e, void **mailbox_out, unsigned char *ipc_handle64) {
  if (!e || !mailbox_out) return PHYLO_ERR_ARG;
  CK(cudaSetDevice(e->device));
  if (!e->xMailbox) {
    CK(cudaMalloc(&e->xMailbox, kXMailboxBytes));
    CK(cudaMemsetAsync(e->xMailbox, 0, kXMailboxBytes, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  *mailbox_out = e->xMailbox;
  if (ipc_handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries IPC handles as 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->xMailbox));
    std::memcpy(ipc_handle64, &h, 64);
  }
  return PHYLO_OK;
}

extern "C" int phylo_exchange_open(phylo_engine *e, const unsigned char *ipc_handle64, void **mailbox_out) {
  if (!e || !ipc_handle64 || !mailbox_out) return PHYLO_ERR_ARG;
  CK(cudaSetDevice(e->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handle64, 64);
  void *p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  e->xOpened.push_back(p);
  *mailbox_out = p;
  return PHYLO_OK;
}

extern "C" int phylo_exchange_set(phylo_engine *e, int world, int rank, void *const *mailboxes) {
  if (!e) return PHYLO_ERR_ARG;
  if (world < 1 || world > kXMaxWorld || rank < 0 || rank >= world || !mailboxes)
    return fail(e, PHYLO_ERR_ARG, "exchange_set: need 1 <= world <= %d, 0 <= rank < world", kXMaxWorld);
  if (!e->xMailbox || mailboxes[rank] != e->xMailbox)
    return fail(e, PHYLO_ERR_ARG, "exchange_set: mailboxes[rank] must be this engine's own mailbox (phylo_exchange_alloc)");
  for (int q = 0; q < world; ++q) {
    if (!mailboxes[q]) return fail(e, PHYLO_ERR_ARG, "exchange_set: mailbox %d is NULL", q);
    e->xPeers.box[q] = (double *)mailboxes[q];
  }
  e->xWorld = world;
  e->xRank = rank;
  return PHYLO_OK;
}

// launches the exchange kernel for `n` doubles at `src` (device) and waits for its result in mapped host memory
static int exchange_run(phylo_engine *e, const double *src, int n, int mode, double *out_bits, const char *who) {
  if (e->xWorld < 1) return fail(e, PHYLO_ERR_STATE, "%s: phylo_exchange_set has not been called", who);
  if (n > kXMaxPartials) return fail(e, PHYLO_ERR_UNSUPPORTED, "%s: %d block partials per rank exceed the mailbox slot (%d)", who, n, kXMaxPartials);
  double *dev_out = nullptr;
  CK(cudaHostGetDevicePointer((void **)&dev_out, e->hScalar + HS_XCHG_OUT, 0));
  const unsigned long long seq = ++e->xSeq;
  volatile double *out = e->hScalar + HS_XCHG_OUT;
  out[1] = 0.0;
  {
    ProfScope prof(e, KC_REDUCE);
    exchange_kernel<<<1, 1024, 0, e->stream>>>(src, n, e->xRank, e->xWorld, e->xPeers, seq, mode, dev_out);
    LAUNCH_CHECK();
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (int spin = 0;; ++spin) {
    double f = out[1];
    unsigned long long bits;
    std::memcpy(&bits, &f, 8);
    if (bits == seq) break;
    if (bits == ~0ull) return fail(e, PHYLO_ERR_CUDA, "%s: a peer did not arrive within the exchange time-out", who);
    if ((spin & 4095) == 4095 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) {
      CK(cudaStreamSynchronize(e->stream));
      return fail(e, PHYLO_ERR_CUDA, "%s: the exchange kernel did not publish a result", who);
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  *out_bits = out[0];
  return PHYLO_OK;
}

extern "C" int phylo_lk_exchange_reduce(phylo_engine *e, double *lnl_out) {
  if (!e || !lnl_out) return PHYLO_ERR_ARG;
  if (!e->lk_evaluated) return fail(e, PHYLO_ERR_STATE, "lk_exchange_reduce: no likelihood evaluation to combine");
  CK(cudaSetDevice(e->device));
  return exchange_run(e, e->dPart, (int)e->nPart, 0, lnl_out, "lk_exchange_reduce");
}

extern "C" int phylo_exchange_sum_u64(phylo_engine *e, uint64_t value, uint64_t *sum_out) {
  if (!e || !sum_out) return PHYLO_ERR_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));  // the mapped input word is about to be rewritten
  std::memcpy(e->hScalar + HS_XCHG_IN, &value, 8);
  double *dev_in = nullptr;
  CK(cudaHostGetDevicePointer((void **)&dev_in, e->hScalar + HS_XCHG_IN, 0));
  double bits = 0.0;
  const int rc = exchange_run(e, dev_in, 1, 1, &bits, "exchange_sum_u64");
  if (rc != PHYLO_OK) return rc;
  std::memcpy(sum_out, &bits, 8);
  return PHYLO_OK;
}

extern "C" int phylo_lk_score_alignment(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                                        const double *weights, int capacity, const phylo_op *ops, int n_ops,
                                        int root_a, int root_b, double root_t, double *lnl_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (n_ops < 0 || (n_ops > 0 && !ops) || !lnl_out) return fail(e, PHYLO_ERR_ARG, "lk_score_alignment: bad arguments");
  int rc;
  if ((rc = lk_prepare_shape(e, T, N, masks, mask_bytes, weights, capacity, "lk_score_alignment")) != PHYLO_OK) return rc;
  if ((rc = lk_validate_schedule(e, ops, n_ops, root_a, root_b, "lk_score_alignment")) != PHYLO_OK) return rc;
  bool done = false;
  if ((rc = lk_score_tree_fused(e, ops, n_ops, root_a, root_b, root_t, &done, masks, mask_bytes)) != PHYLO_OK) {
    lk_free_data(e);
    return rc;
  }
  if (!done && (rc = lk_score_tree_fusedm(e, ops, n_ops, root_a, root_b, root_t, &done, masks, mask_bytes)) != PHYLO_OK) {
    lk_free_data(e);  // (20 / 61 states: the tree-fused DMMA kernel, slabs uploaded while the ones before are scored)
    return rc;
  }
  if (!done) {  // not eligible for the fused kernels: plain upload, then the per-node path
    if ((rc = lk_upload_slab(e, masks, mask_bytes, 0, N, nullptr, nullptr, e->stream)) != PHYLO_OK) { lk_free_data(e); return rc; }
    if ((rc = lk_check_bad(e, "lk_score_alignment")) != PHYLO_OK) return rc;
    return phylo_lk_score_tree(e, ops, n_ops, root_a, root_b, root_t, lnl_out);
  }
  CK(cudaStreamSynchronize(e->stream));
  if ((rc = lk_check_bad(e, "lk_score_alignment")) != PHYLO_OK) return rc;
  *lnl_out = e->hScalar[0];
  e->lk_evaluated = true;
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

extern "C" int phylo_lk_edge_lnl(phylo_engine *e, int a_slot, int b_slot, const double *t, int n_t,
                                 double *lnl_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_edge_lnl: no tips loaded");
  if (!t || n_t < 1 || !lnl_out) return fail(e, PHYLO_ERR_ARG, "lk_edge_lnl: bad arguments");
  CK(cudaSetDevice(e->device));
  Operand a, b;
  int rc;
  if ((rc = lk_operand(e, a_slot, &a, "lk_edge_lnl")) != PHYLO_OK) return rc;
  if ((rc = lk_operand(e, b_slot, &b, "lk_edge_lnl")) != PHYLO_OK) return rc;
  if ((rc = ensure_pt_capacity(e, n_t, e->S, e->K)) != PHYLO_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n_t; ++i) e->hT[i] = t[i];
  if ((rc = build_pt(e, n_t)) != PHYLO_OK) return rc;
  const size_t pk = (size_t)e->K * e->S * e->S;
  for (int i = 0; i < n_t; ++i) {
    if ((rc = lk_root_eval(e, e->dP + (size_t)i * pk, a, b, e->hScalar)) != PHYLO_OK) return rc;
    CK(cudaStreamSynchronize(e->stream));
    lnl_out[i] = e->hScalar[0];
  }
  e->lk_evaluated = true;
  return PHYLO_OK;
}

// lnL across many edges in one call: one P(t) build, one root join + fold per edge, results gathered in
// pinned memory, ONE synchronisation -- the cost of an SPR / TBR candidate once the directional CLVs exist
extern "C" int phylo_lk_edge_lnl_batch(phylo_engine *e, int n_edges, const int32_t *a_slots, const int32_t *b_slots,
                                       const double *t, double *lnl_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_edge_lnl_batch: no tips loaded");
  if (n_edges < 1 || !a_slots || !b_slots || !t || !lnl_out) return fail(e, PHYLO_ERR_ARG, "lk_edge_lnl_batch: bad arguments");
  CK(cudaSetDevice(e->device));
  int rc;
  if ((rc = ensure_pt_capacity(e, n_edges, e->S, e->K)) != PHYLO_OK) return rc;
  if (n_edges > e->capEdgeRes) {
    if (e->hEdgeRes) { cudaFreeHost(e->hEdgeRes); e->hEdgeRes = nullptr; }
    e->capEdgeRes = 0;
    CK(cudaMallocHost(&e->hEdgeRes, sizeof(double) * std::max(n_edges, 1024)));
    e->capEdgeRes = std::max(n_edges, 1024);
  }
  double *hres = e->hEdgeRes;
  CK(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n_edges; ++i) e->hT[i] = t[i];
  rc = build_pt(e, n_edges);
  const size_t pk = (size_t)e->K * e->S * e->S;
  // 4 states: every (edge, 1024-pattern block) pair is a work item of ONE launch, then the
  // remaining fold levels for all edges at once (y grid dimension <= 65535 edges per launch)
  if (rc == PHYLO_OK && e->S == 4 && (e->K == 1 || e->K == 2 || e->K == 4 || e->K == 8 || e->K == 16) && n_edges <= 65535) {
    std::vector<EdgeJoin> he(n_edges);
    for (int i = 0; i < n_edges; ++i) {
      Operand a, b;
      if ((rc = lk_operand(e, a_slots[i], &a, "lk_edge_lnl_batch")) != PHYLO_OK) break;
      if ((rc = lk_operand(e, b_slots[i], &b, "lk_edge_lnl_batch")) != PHYLO_OK) break;
      he[i] = EdgeJoin{a.src, b.src, a.scale, b.scale, a.tip ? 1 : 0, b.tip ? 1 : 0};
    }
    const int64_t nb1 = e->nPart, nb2 = (nb1 + kLnlBlock - 1) / kLnlBlock, nb3 = (nb2 + kLnlBlock - 1) / kLnlBlock;
    const size_t descBytes = (sizeof(EdgeJoin) * n_edges + 255) & ~(size_t)255;
    const size_t bytes = descBytes + sizeof(double) * (size_t)n_edges * (size_t)(nb1 + nb2 + nb3 + 1);
    if (rc == PHYLO_OK && bytes > e->capEdgeArena) {  // grow-only scratch kept across calls (a search calls this per candidate set)
      dfree(e->dEdgeArena);
      e->capEdgeArena = 0;
      if (cudaMalloc(&e->dEdgeArena, bytes) != cudaSuccess) {
        cudaGetLastError();
        rc = fail(e, PHYLO_ERR_CUDA, "lk_edge_lnl_batch: cannot allocate %zu bytes of device memory", bytes);
      } else {
        e->capEdgeArena = bytes;
      }
    }
    char *arena = e->dEdgeArena;
    if (rc == PHYLO_OK) {
      EdgeJoin *dE = (EdgeJoin *)arena;
      double *p1 = (double *)(arena + descBytes), *p2 = p1 + (size_t)n_edges * nb1, *p3 = p2 + (size_t)n_edges * nb2;
      cudaMemcpyAsync(dE, he.data(), sizeof(EdgeJoin) * n_edges, cudaMemcpyHostToDevice, e->stream);
      {
        ProfScope prof(e, KC_ROOT);
        const int g = (int)std::min<int64_t>(nb1 * n_edges, (int64_t)e->sm_count * 8);
#define RB(KV) root4_batch_kernel<KV><<<g, 256, 0, e->stream>>>(dE, n_edges, e->dP, e->dPi, e->dProbs, e->pinvar, \
                                                                (const uint8_t *)e->dInv, e->dWeights, p1, e->N)
        switch (e->K) {
          case 1: RB(1); break;
          case 2: RB(2); break;
          case 4: RB(4); break;
          case 8: RB(8); break;
          default: RB(16);
        }
#undef RB
        ++e->launches;
      }
      {
        ProfScope prof(e, KC_REDUCE);
        const double *cur = p1;
        int64_t n = nb1;
        double *outs[2] = {p2, p3};
        int lvl = 0;
        do {
          const int64_t nb = (n + kLnlBlock - 1) / kLnlBlock;
          reduce1024_rows_kernel<<<dim3((unsigned)nb, (unsigned)n_edges), 256, 0, e->stream>>>(cur, n, outs[lvl & 1]);
          ++e->launches;
          cur = outs[lvl & 1];
          ++lvl;
          n = nb;
        } while (n > 1);
        cudaMemcpyAsync(hres, cur, sizeof(double) * n_edges, cudaMemcpyDeviceToHost, e->stream);
      }
    }
    const cudaError_t st = cudaStreamSynchronize(e->stream);
    const cudaError_t le = cudaGetLastError();
    if (rc == PHYLO_OK && st == cudaSuccess && le == cudaSuccess)
      for (int i = 0; i < n_edges; ++i) lnl_out[i] = hres[i];
    if (rc != PHYLO_OK) return rc;
    if (st != cudaSuccess || le != cudaSuccess)
      return fail(e, PHYLO_ERR_CUDA, "lk_edge_lnl_batch: %s", cudaGetErrorString(st != cudaSuccess ? st : le));
    // (site lnL / level-1 partials of the last whole evaluation are left as they were)
    if (e->prof_on) prof_resolve_lazy(e);
    return PHYLO_OK;
  }
  for (int i = 0; i < n_edges && rc == PHYLO_OK; ++i) {
    Operand a, b;
    if ((rc = lk_operand(e, a_slots[i], &a, "lk_edge_lnl_batch")) != PHYLO_OK) break;
    if ((rc = lk_operand(e, b_slots[i], &b, "lk_edge_lnl_batch")) != PHYLO_OK) break;
    rc = lk_root_eval(e, e->dP + (size_t)i * pk, a, b, hres + i);
  }
  const cudaError_t st = cudaStreamSynchronize(e->stream);
  if (rc == PHYLO_OK && st == cudaSuccess)
    for (int i = 0; i < n_edges; ++i) lnl_out[i] = hres[i];
  if (rc != PHYLO_OK) return rc;
  if (st != cudaSuccess) return fail(e, PHYLO_ERR_CUDA, "lk_edge_lnl_batch: %s", cudaGetErrorString(st));
  e->lk_evaluated = true;
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

// ---------------------------------------------------------- site-pattern compression ----
template <int EB>
static void launch_cmp_transpose(const uint8_t *in, uint8_t *rec, int T, int64_t N, int TP, cudaStream_t st) {
  dim3 grid((unsigned)((N + 31) / 32), (unsigned)((T + 31) / 32));
  cmp_transpose_kernel<EB><<<grid, 256, 0, st>>>(in, rec, T, N, TP);
}
template <int EB>
static void launch_cmp_gather(const uint8_t *in, int T, int64_t N, const int *rep_site, int64_t P, uint8_t *out, int g,
                              cudaStream_t st) {
  cmp_gather_kernel<EB><<<g, 256, 0, st>>>(in, T, N, rep_site, P, out);
}

extern "C" int phylo_compress_patterns(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                                       const double *weights_in, void *patterns_out, double *weights_out,
                                       int32_t *site_to_pattern, int64_t *n_patterns) {
  return phylo_compress_patterns_pitched(e, T, N, masks, mask_bytes, 0, weights_in, patterns_out, weights_out,
                                         site_to_pattern, n_patterns);
}

// host_pitch_bytes != 0: row t of the host alignment starts at masks + t * host_pitch_bytes (a column slab of a
// wider matrix: what phylo_group_compress_patterns hands each device)
extern "C" int phylo_compress_patterns_pitched(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                                               uint64_t host_pitch_bytes, const double *weights_in, void *patterns_out,
                                               double *weights_out, int32_t *site_to_pattern, int64_t *n_patterns) {
  if (!e) return PHYLO_ERR_ARG;
  if (host_pitch_bytes != 0 && N > 0 && mask_bytes > 0 && host_pitch_bytes < (uint64_t)N * (uint64_t)mask_bytes)
    return fail(e, PHYLO_ERR_ARG, "compress_patterns: host pitch %llu is shorter than a row of %lld cells",
                (unsigned long long)host_pitch_bytes, (long long)N);
  if (T < 1 || N < 1 || N > 2000000000ll || !masks || !patterns_out || !weights_out || !n_patterns ||
      !(mask_bytes == 1 || mask_bytes == 2 || mask_bytes == 4 || mask_bytes == 8))
    return fail(e, PHYLO_ERR_ARG, "compress_patterns: bad arguments (T=%d N=%lld mask_bytes=%d)", T, (long long)N, mask_bytes);
  if (e->dSymTab && (mask_bytes != 1 || !e->symtab_fits_byte))
    return fail(e, PHYLO_ERR_UNSUPPORTED,
                "compress_patterns: with a symbol table the alignment must be 1 byte per cell and every table entry < 256");
  CK(cudaSetDevice(e->device));
  const int EB = mask_bytes, TP = (int)(((size_t)T * EB + 15) / 16 * 16);
  uint64_t M = 1;
  while (M < 2 * (uint64_t)N) M <<= 1;
  const int64_t nb = (N + 1023) / 1024;
  // one arena for all temporaries (freed on return)
  struct Part { size_t off, bytes; };
  size_t total = 0;
  auto reserve = [&](size_t bytes) { Part p{total, bytes}; total += (bytes + 255) & ~(size_t)255; return p; };
  const Part pIn = reserve((size_t)T * N * EB), pRec = reserve((size_t)N * TP), pKey = reserve(8 * (size_t)N),
             pTKey = reserve(8 * (size_t)M), pTRep = reserve(4 * (size_t)M), pSlot = reserve(4 * (size_t)N),
             pFlag = reserve(4 * (size_t)N), pBsum = reserve(4 * (size_t)(nb + 1)), pPid = reserve(4 * (size_t)N),
             pRepSite = reserve(4 * (size_t)N), pSitePat = reserve(4 * (size_t)N), pWout = reserve(8 * (size_t)N),
             pWin = reserve(weights_in ? 8 * (size_t)N : 0), pScal = reserve(64), pOut = reserve((size_t)T * N * EB);
  char *arena = nullptr;
  if (cudaMalloc(&arena, total) != cudaSuccess) {
    cudaGetLastError();
    return fail(e, PHYLO_ERR_CUDA, "compress_patterns: cannot allocate %zu bytes of device memory", total);
  }
  struct Free { char *p; ~Free() { cudaFree(p); } } guard{arena};
  uint8_t *dIn = (uint8_t *)(arena + pIn.off), *dRec = (uint8_t *)(arena + pRec.off), *dOut = (uint8_t *)(arena + pOut.off);
  uint64_t *dKey = (uint64_t *)(arena + pKey.off);
  unsigned long long *dTKey = (unsigned long long *)(arena + pTKey.off), *dColl = (unsigned long long *)(arena + pScal.off);
  long long *dTotal = (long long *)(arena + pScal.off + 8);
  int *dTRep = (int *)(arena + pTRep.off), *dFlag = (int *)(arena + pFlag.off), *dBsum = (int *)(arena + pBsum.off),
      *dPid = (int *)(arena + pPid.off), *dRepSite = (int *)(arena + pRepSite.off), *dSitePat = (int *)(arena + pSitePat.off);
  uint32_t *dSlot = (uint32_t *)(arena + pSlot.off);
  double *dWout = (double *)(arena + pWout.off), *dWin = weights_in ? (double *)(arena + pWin.off) : nullptr;
  cudaStream_t st = e->stream;
  if (host_pitch_bytes == 0 || host_pitch_bytes == (uint64_t)N * EB)
    CK(cudaMemcpyAsync(dIn, masks, (size_t)T * N * EB, cudaMemcpyHostToDevice, st));
  else
    CK(cudaMemcpy2DAsync(dIn, (size_t)N * EB, masks, (size_t)host_pitch_bytes, (size_t)N * EB, (size_t)T, cudaMemcpyHostToDevice, st));
  if (weights_in) CK(cudaMemcpyAsync(dWin, weights_in, 8 * (size_t)N, cudaMemcpyHostToDevice, st));
  if (e->dSymTab) {  // the cells are alphabet symbols: translate in place, patterns come out as state masks
    ProfScope prof(e, KC_COMPRESS);
    CK(cudaMemsetAsync(dColl, 0, 8, st));
    cmp_symbols_kernel<<<grid_for((int64_t)(((size_t)T * N + 15) / 16), 256, e->sm_count * 16), 256, 0, st>>>(
        dIn, (size_t)T * N, e->dSymTab, dColl);
    LAUNCH_CHECK();
    unsigned long long unknown = 0;
    CK(cudaMemcpyAsync(&unknown, dColl, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (unknown)
      return fail(e, PHYLO_ERR_DATA, "compress_patterns: %llu cells hold a symbol the table does not know", unknown);
  }
  const int g = grid_for(N, 256, e->sm_count * 16);
  long long hTotal = 0;
  unsigned long long hColl = 0;
  {
    ProfScope prof(e, KC_COMPRESS);
    const uint64_t kCmpSeed0 = 0x243f6a8885a308d3ull;
    bool fused_hash = false;
    const int chunks = TP / 16;
    const int G = chunks >= 32 ? 32 : (chunks >= 16 ? 16 : (chunks >= 8 ? 8 : (chunks >= 4 ? 4 : (chunks >= 2 ? 2 : 1))));
    const int gG = grid_for(N, 256 / G, e->sm_count * 16);
    CK(cudaMemsetAsync(dRec, 0, (size_t)N * TP, st));  // record padding must compare equal
    switch (EB) {
      case 1:
        if (N % 4 == 0) {
          dim3 grid((unsigned)((N + kCmpTileSites - 1) / kCmpTileSites), (unsigned)((T + kCmpTileTaxa - 1) / kCmpTileTaxa));
          fused_hash = grid.y == 1;  // whole records pass through shared memory: hash them on the way out
          cmp_transpose4_kernel<<<grid, 256, kCmpTileSites * kCmpPitch, st>>>(dIn, dRec, T, N, TP, kCmpSeed0,
                                                                             fused_hash ? dKey : nullptr);
        } else {
          launch_cmp_transpose<1>(dIn, dRec, T, N, TP, st);
        }
        break;
      case 2: launch_cmp_transpose<2>(dIn, dRec, T, N, TP, st); break;
      case 4: launch_cmp_transpose<4>(dIn, dRec, T, N, TP, st); break;
      default: launch_cmp_transpose<8>(dIn, dRec, T, N, TP, st);
    }
    LAUNCH_CHECK();
    for (int attempt = 0;; ++attempt) {
      CK(cudaMemsetAsync(dTKey, 0, 8 * (size_t)M, st));
      CK(cudaMemsetAsync(dTRep, 0x7f, 4 * (size_t)M, st));
      CK(cudaMemsetAsync(dColl, 0, 16, st));
      if (!(fused_hash && attempt == 0)) {
        const uint64_t seed = kCmpSeed0 + 0x9e3779b97f4a7c15ull * (uint64_t)attempt;
        switch (G) {
          case 32: cmp_hash_kernel<32><<<gG, 256, 0, st>>>(dRec, N, TP, seed, dKey); break;
          case 16: cmp_hash_kernel<16><<<gG, 256, 0, st>>>(dRec, N, TP, seed, dKey); break;
          case 8: cmp_hash_kernel<8><<<gG, 256, 0, st>>>(dRec, N, TP, seed, dKey); break;
          case 4: cmp_hash_kernel<4><<<gG, 256, 0, st>>>(dRec, N, TP, seed, dKey); break;
          case 2: cmp_hash_kernel<2><<<gG, 256, 0, st>>>(dRec, N, TP, seed, dKey); break;
          default: cmp_hash_kernel<1><<<gG, 256, 0, st>>>(dRec, N, TP, seed, dKey);
        }
        LAUNCH_CHECK();
      }
      cmp_insert_kernel<<<g, 256, 0, st>>>(dKey, N, dTKey, dTRep, M - 1, dSlot);
      LAUNCH_CHECK();
      switch (G) {
        case 32: cmp_verify_kernel<32><<<gG, 256, 0, st>>>(dRec, N, TP, dTRep, dSlot, dFlag, dColl); break;
        case 16: cmp_verify_kernel<16><<<gG, 256, 0, st>>>(dRec, N, TP, dTRep, dSlot, dFlag, dColl); break;
        case 8: cmp_verify_kernel<8><<<gG, 256, 0, st>>>(dRec, N, TP, dTRep, dSlot, dFlag, dColl); break;
        case 4: cmp_verify_kernel<4><<<gG, 256, 0, st>>>(dRec, N, TP, dTRep, dSlot, dFlag, dColl); break;
        case 2: cmp_verify_kernel<2><<<gG, 256, 0, st>>>(dRec, N, TP, dTRep, dSlot, dFlag, dColl); break;
        default: cmp_verify_kernel<1><<<gG, 256, 0, st>>>(dRec, N, TP, dTRep, dSlot, dFlag, dColl);
      }
      LAUNCH_CHECK();
      CK(cudaMemcpyAsync(&hColl, dColl, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (hColl == 0) break;
      if (attempt == 3) return fail(e, PHYLO_ERR_NUMERIC, "compress_patterns: hash collisions persist after 4 seeds");
    }
    cmp_scan_block_sums<<<(int)nb, 256, 0, st>>>(dFlag, N, dBsum);
    LAUNCH_CHECK();
    cmp_scan_sums<<<1, 1024, 0, st>>>(dBsum, nb, dTotal);
    LAUNCH_CHECK();
    cmp_scan_apply<<<(int)nb, 256, 0, st>>>(dFlag, N, dBsum, dPid, dRepSite);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(&hTotal, dTotal, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(dWout, 0, 8 * (size_t)N, st));
    cmp_assign_kernel<<<g, 256, 0, st>>>(N, dTRep, dSlot, dPid, dWin, dSitePat, dWout);
    LAUNCH_CHECK();
    CK(cudaStreamSynchronize(st));
    const int64_t P = hTotal;
    const int gg = grid_for((int64_t)T * P, 256, e->sm_count * 16);
    switch (EB) {
      case 1: launch_cmp_gather<1>(dIn, T, N, dRepSite, P, dOut, gg, st); break;
      case 2: launch_cmp_gather<2>(dIn, T, N, dRepSite, P, dOut, gg, st); break;
      case 4: launch_cmp_gather<4>(dIn, T, N, dRepSite, P, dOut, gg, st); break;
      default: launch_cmp_gather<8>(dIn, T, N, dRepSite, P, dOut, gg, st);
    }
    LAUNCH_CHECK();
  }
  const int64_t P = hTotal;
  CK(cudaMemcpyAsync(patterns_out, dOut, (size_t)T * P * EB, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(weights_out, dWout, 8 * (size_t)P, cudaMemcpyDeviceToHost, st));
  if (site_to_pattern) CK(cudaMemcpyAsync(site_to_pattern, dSitePat, 4 * (size_t)N, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *n_patterns = P;
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

// ------------------------------------------------------------ branch-length loop ----
static const int kEdgeMaxT = 16;  // branch lengths per edge_eval pass

extern "C" int phylo_lk_edge_prepare(phylo_engine *e, int a_slot, int b_slot) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_edge_prepare: no tips loaded");
  CK(cudaSetDevice(e->device));
  Operand a, b;
  int rc;
  if ((rc = lk_operand(e, a_slot, &a, "lk_edge_prepare")) != PHYLO_OK) return rc;
  if ((rc = lk_operand(e, b_slot, &b, "lk_edge_prepare")) != PHYLO_OK) return rc;
  const size_t KS = (size_t)e->K * e->S;
  if (!e->dSum) {
    CK(cudaMalloc(&e->dSum, sizeof(double) * (size_t)e->N * KS));
    CK(cudaMalloc(&e->dSumSc, sizeof(int32_t) * (size_t)e->N));
  }
  if (!e->dEdgePart) CK(cudaMalloc(&e->dEdgePart, sizeof(double) * (size_t)e->nPart * 3 * kEdgeMaxT));
  if (!e->dEdgeOut) CK(cudaMalloc(&e->dEdgeOut, sizeof(double) * 3 * kEdgeMaxT));
  if (!e->dEdgeT) CK(cudaMalloc(&e->dEdgeT, sizeof(double) * kEdgeMaxT));
  // one pruning update with M1 / M2 in place of P_left / P_right (see phylo_lk_set_model): the
  // DNA, DMMA (20 / 61 states) and any-S kernels, their tip paths and their rescaling all apply;
  // the "scale counters" that come out are exactly the summed counters the evaluation needs
  if ((rc = lk_launch_prune(e, e->dUL, e->dUR, a, b, e->dSum, e->dSumSc)) != PHYLO_OK) return rc;
  e->edge_ready = true;
  e->edge_a = a_slot;
  e->edge_b = b_slot;
  return PHYLO_OK;
}

extern "C" int phylo_lk_edge_eval(phylo_engine *e, const double *t, int n_t, double *lnl_out, double *d1_out,
                                  double *d2_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (!e->edge_ready) return fail(e, PHYLO_ERR_STATE, "lk_edge_eval: call phylo_lk_edge_prepare first");
  if (!t || n_t < 1 || !lnl_out) return fail(e, PHYLO_ERR_ARG, "lk_edge_eval: bad arguments");
  if (e->nPart > (int64_t)kLnlBlock * kLnlBlock) return fail(e, PHYLO_ERR_UNSUPPORTED, "lk_edge_eval: more than 2^30 patterns");
  CK(cudaSetDevice(e->device));
  const int KS = e->K * e->S;
  for (int t0 = 0; t0 < n_t; t0 += kEdgeMaxT) {
    // as many lengths per pass as the coefficient table allows (<= 48 KB of shared memory)
    const int per = std::max(1, std::min(kEdgeMaxT, (int)(22 * 1024 / (sizeof(double) * 3 * KS))));  // + 24 KB static
    for (int c0 = t0; c0 < std::min(n_t, t0 + kEdgeMaxT); c0 += per) {
      const int nc = std::min(per, std::min(n_t, t0 + kEdgeMaxT) - c0);
      EdgeLengths tl;
      for (int i = 0; i < 16; ++i) tl.t[i] = i < nc ? t[c0 + i] : 0.0;
      const size_t smem = sizeof(double) * 3 * (size_t)KS * nc;
      const int g = (int)std::min<int64_t>(e->nPart, (int64_t)e->sm_count * 4);
      {
        ProfScope prof(e, KC_EDGE);
#define EDGE_EVAL(MT, GV, PV)                                                                                      \
  edge_eval_kernel<MT, GV, PV><<<g, 256, smem, e->stream>>>(e->dSum, e->dSumSc, e->dLam, e->dRates, e->dProbs, e->dPi, \
                                                             e->pinvar, (const MT *)e->dInv, e->dWeights, tl, nc, \
                                                             e->sym, e->S, e->K, e->N, e->dEdgePart)
#define EDGE_EVAL_G(MT)                                                      \
  {                                                                          \
    if (KS == 4) EDGE_EVAL(MT, 1, 4);                                        \
    else if (KS < 4) EDGE_EVAL(MT, 1, 0);                                    \
    else if (KS == 16) EDGE_EVAL(MT, 4, 4);                                  \
    else if (KS <= 16) EDGE_EVAL(MT, 4, 0);                                  \
    else if (KS <= 32) EDGE_EVAL(MT, 8, 0);                                  \
    else if (KS <= 64) EDGE_EVAL(MT, 32, 2);                                 \
    else if (KS <= 80) EDGE_EVAL(MT, 16, 5);                                 \
    else if (KS <= 128) EDGE_EVAL(MT, 16, 0);                                \
    else EDGE_EVAL(MT, 32, 0);                                               \
  }
        if (e->mask_dev_bytes == 1) EDGE_EVAL_G(uint8_t)
        else if (e->mask_dev_bytes == 4) EDGE_EVAL_G(uint32_t)
        else EDGE_EVAL_G(uint64_t)
#undef EDGE_EVAL_G
#undef EDGE_EVAL
        LAUNCH_CHECK();
        // the folded sums go straight into mapped host memory: no D2H copy
        double *host_out = nullptr;
        CK(cudaHostGetDevicePointer((void **)&host_out, e->hScalar + HS_EDGE_OUT, 0));
        fold_rows_kernel<<<3 * nc, 256, 0, e->stream>>>(e->dEdgePart, e->nPart, host_out);
        LAUNCH_CHECK();
      }
      CK(cudaStreamSynchronize(e->stream));
      for (int i = 0; i < nc; ++i) {
        lnl_out[c0 + i] = e->hScalar[HS_EDGE_OUT + 3 * i];
        if (d1_out) d1_out[c0 + i] = e->hScalar[HS_EDGE_OUT + 3 * i + 1];
        if (d2_out) d2_out[c0 + i] = e->hScalar[HS_EDGE_OUT + 3 * i + 2];
      }
    }
  }
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

// Safeguarded Newton on t -> lnL(t) over [t_min, t_max]: the bracket shrinks with the sign of
// the first derivative; a Newton step that leaves the bracket (or a non-concave point) is
// replaced by the bracket's geometric mean. Every iteration is one pass over the sum table.
extern "C" int phylo_lk_optimize_branch(phylo_engine *e, int a_slot, int b_slot, double t0, double t_min, double t_max,
                                        double tol, int max_iter, double *t_opt, double *lnl_opt, int *iters_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (!(t_min > 0.0) || !(t_max > t_min) || !t_opt) return fail(e, PHYLO_ERR_ARG, "lk_optimize_branch: need 0 < t_min < t_max");
  int rc;
  if ((rc = phylo_lk_edge_prepare(e, a_slot, b_slot)) != PHYLO_OK) return rc;
  double lo = t_min, hi = t_max, t = std::min(std::max(t0, t_min), t_max);
  double glo = 0.0, ghi = 0.0;  // dlnL/dt at the bracket ends once they have been visited
  bool has_lo = false, has_hi = false;
  double lnl = 0.0, d1 = 0.0, d2 = 0.0, best_t = t, best_lnl = -INFINITY;
  int it = 0;
  if (tol <= 0.0) tol = 1e-8;
  if (max_iter < 1) max_iter = 50;
  for (; it < max_iter; ++it) {
    if ((rc = phylo_lk_edge_eval(e, &t, 1, &lnl, &d1, &d2)) != PHYLO_OK) return rc;
    if (lnl > best_lnl) { best_lnl = lnl; best_t = t; }
    if (d1 > 0.0) { lo = t; glo = d1; has_lo = true; } else { hi = t; ghi = d1; has_hi = true; }
    double next = (d2 < 0.0) ? t - d1 / d2 : -1.0;  // Newton
    if (!(next > lo && next < hi)) {
      if (has_lo && has_hi) {  // regula falsi on the derivative, kept off the ends (Illinois-style damping)
        next = lo - glo * (hi - lo) / (ghi - glo);
        const double margin = 0.05 * (hi - lo);
        next = std::min(std::max(next, lo + margin), hi - margin);
      } else {
        next = std::sqrt(lo * hi);  // no sign change seen yet: walk the bracket geometrically
      }
    }
    const bool at_bound = (t <= t_min && d1 <= 0.0) || (t >= t_max && d1 >= 0.0);
    if (at_bound || std::fabs(next - t) <= tol * std::max(t, 1e-8) || hi - lo <= tol * std::max(lo, 1e-8)) { ++it; break; }
    t = next;
  }
  *t_opt = best_t;
  if (lnl_opt) *lnl_opt = best_lnl;
  if (iters_out) *iters_out = it;
  return PHYLO_OK;
}

extern "C" int phylo_lk_get_clv(phylo_engine *e, int node, double *clv_out, int32_t *scale_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->T == 0) return fail(e, PHYLO_ERR_STATE, "lk_get_clv: no tips loaded");
  if (node < e->T || node >= e->cap || !e->nodes[node].valid)
    return fail(e, PHYLO_ERR_ARG, "lk_get_clv: slot %d holds no interior CLV", node);
  if (!clv_out) return fail(e, PHYLO_ERR_ARG, "lk_get_clv: clv_out is NULL");
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(clv_out, e->nodes[node].clv, sizeof(double) * (size_t)e->N * e->K * e->S,
                     cudaMemcpyDeviceToHost, e->stream));
  if (scale_out)
    CK(cudaMemcpyAsync(scale_out, e->nodes[node].scale, sizeof(int32_t) * (size_t)e->N, cudaMemcpyDeviceToHost,
                       e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return PHYLO_OK;
}

extern "C" int phylo_lk_get_site_lnl(phylo_engine *e, double *out) {
  if (!e) return PHYLO_ERR_ARG;
  if (!e->lk_evaluated || !out) return fail(e, PHYLO_ERR_STATE, "lk_get_site_lnl: nothing evaluated yet");
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(out, e->dSite, sizeof(double) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return PHYLO_OK;
}

extern "C" int phylo_lk_get_block_partials(phylo_engine *e, double *out, int64_t *n_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (!e->lk_evaluated) return fail(e, PHYLO_ERR_STATE, "lk_get_block_partials: nothing evaluated yet");
  CK(cudaSetDevice(e->device));
  if (n_out) *n_out = e->nPart;
  if (out) {
    CK(cudaMemcpyAsync(out, e->dPart, sizeof(double) * (size_t)e->nPart, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return PHYLO_OK;
}

// Host restatement of levels >= 2 of the canonical reduction, used to combine the level-1
// block partials gathered from several ranks (bit-identical for any rank count).
static double fold1024_host(const double *v, int64_t n) {
  double w[32];
  for (int g = 0; g < 32; ++g) {
    double y[32];
    for (int j = 0; j < 32; ++j) {
      const int64_t idx = (int64_t)g * 32 + j;
      y[j] = idx < n ? v[idx] : 0.0;
    }
    for (int off = 16; off >= 1; off >>= 1)
      for (int j = 0; j < off; ++j) y[j] = y[j] + y[j + off];
    w[g] = y[0];
  }
  for (int off = 16; off >= 1; off >>= 1)
    for (int j = 0; j < off; ++j) w[j] = w[j] + w[j + off];
  return w[0];
}

extern "C" double phylo_reduce_partials(const double *partials, int64_t n) {
  if (!partials || n <= 0) return 0.0;
  std::vector<double> cur(partials, partials + n);
  while (cur.size() > 1) {
    const int64_t m = (int64_t)cur.size(), nb = (m + kLnlBlock - 1) / kLnlBlock;
    std::vector<double> nxt(nb);
    for (int64_t b = 0; b < nb; ++b)
      nxt[b] = fold1024_host(cur.data() + b * kLnlBlock, std::min<int64_t>(kLnlBlock, m - b * kLnlBlock));
    cur.swap(nxt);
  }
  return cur[0];
}

// ------------------------------------------------------------------------ Fitch ----
static int np_device(int n_states) {
  static const int sizes[] = {1, 2, 3, 4, 5, 6, 8, 12, 16, 24, 32, 64};
  for (int s : sizes)
    if (n_states <= s) return s;
  return 64;
}

#define NP_DISPATCH(np, EXPR)                                   \
  switch (np) {                                                 \
    case 1: { constexpr int NP = 1; EXPR; } break;              \
    case 2: { constexpr int NP = 2; EXPR; } break;              \
    case 3: { constexpr int NP = 3; EXPR; } break;              \
    case 4: { constexpr int NP = 4; EXPR; } break;              \
    case 5: { constexpr int NP = 5; EXPR; } break;              \
    case 6: { constexpr int NP = 6; EXPR; } break;              \
    case 8: { constexpr int NP = 8; EXPR; } break;              \
    case 12: { constexpr int NP = 12; EXPR; } break;            \
    case 16: { constexpr int NP = 16; EXPR; } break;            \
    case 24: { constexpr int NP = 24; EXPR; } break;            \
    case 32: { constexpr int NP = 32; EXPR; } break;            \
    default: { constexpr int NP = 64; EXPR; } break;            \
  }

static int fitch_ensure(phylo_engine *e, int slot, bool fin) {
  std::vector<uint32_t *> &v = fin ? e->fFin : e->fPre;
  if (v[slot]) return PHYLO_OK;
  // padded to whole 32-word tiles so the TMA-staged tree kernel can bulk-copy full tile rows
  const size_t words_pad = (size_t)((e->fWords + 31) / 32) * 32;
  CK(cudaMalloc(&v[slot], sizeof(uint32_t) * words_pad * e->fNPdev));
  CK(cudaMemsetAsync(v[slot], 0, sizeof(uint32_t) * words_pad * e->fNPdev, e->stream));
  e->tabDirty = true;
  return PHYLO_OK;
}

static int fitch_sync_tables(phylo_engine *e) {
  if (!e->tabDirty) return PHYLO_OK;
  CK(cudaMemcpyAsync(e->dPreTab, e->fPre.data(), sizeof(uint32_t *) * e->fcap, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->dFinTab, e->fFin.data(), sizeof(uint32_t *) * e->fcap, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));  // the std::vector storage is pageable
  e->tabDirty = false;
  return PHYLO_OK;
}

static int fitch_stage(phylo_engine *e, size_t bytes) {
  if (bytes <= e->capStage) return PHYLO_OK;
  CK(cudaStreamSynchronize(e->stream));
  dfree(e->dStage);
  e->capStage = 0;
  CK(cudaMalloc(&e->dStage, bytes));
  e->capStage = bytes;
  return PHYLO_OK;
}

static int fitch_cost_capacity(phylo_engine *e, size_t n) {
  if (n <= e->capCost) return PHYLO_OK;
  CK(cudaStreamSynchronize(e->stream));
  dfree(e->dCost);
  if (e->hCost) { cudaFreeHost(e->hCost); e->hCost = nullptr; }
  e->hCostDev = nullptr;
  e->capCost = 0;
  CK(cudaMalloc(&e->dCost, sizeof(unsigned long long) * n));
  CK(cudaMallocHost(&e->hCost, sizeof(unsigned long long) * n));
  e->capCost = n;
  return PHYLO_OK;
}

static int fitch_sched_capacity(phylo_engine *e, size_t bytes) {
  if (bytes <= e->capSched) return PHYLO_OK;
  CK(cudaStreamSynchronize(e->stream));
  dfree(e->dSched);
  if (e->hSched) { cudaFreeHost(e->hSched); e->hSched = nullptr; }
  e->capSched = 0;
  CK(cudaMalloc(&e->dSched, bytes));
  CK(cudaMallocHost(&e->hSched, bytes));
  e->capSched = bytes;
  return PHYLO_OK;
}

// encode `count` nodes worth of reference-layout codes (already on device in dStage) into
// plane buffers; returns the number of empty elements through *bad
static int fitch_encode(phylo_engine *e, const void *dcodes, uint32_t *dst, unsigned long long *dBad,
                        bool symbols = true) {
  const int g = grid_for(e->fWords * 32, 256, e->sm_count * 8);
  const uint64_t *lut = symbols ? e->dSymTab : nullptr;
  ProfScope prof(e, KC_FITCH_TRANSCODE);
  switch (e->felt) {
    case 1: fitch_encode_kernel<uint8_t><<<g, 256, 0, e->stream>>>((const uint8_t *)dcodes, dst, e->fN, e->fWords, e->fNPdev, dBad, lut); break;
    case 2: fitch_encode_kernel<uint16_t><<<g, 256, 0, e->stream>>>((const uint16_t *)dcodes, dst, e->fN, e->fWords, e->fNPdev, dBad, nullptr); break;
    case 4: fitch_encode_kernel<uint32_t><<<g, 256, 0, e->stream>>>((const uint32_t *)dcodes, dst, e->fN, e->fWords, e->fNPdev, dBad, nullptr); break;
    default: fitch_encode_kernel<uint64_t><<<g, 256, 0, e->stream>>>((const uint64_t *)dcodes, dst, e->fN, e->fWords, e->fNPdev, dBad, nullptr);
  }
  LAUNCH_CHECK();
  return PHYLO_OK;
}

extern "C" int phylo_fitch_set_tips(phylo_engine *e, int T, int64_t N, int elt_bytes, int n_states,
                                    const void *codes, const double *weights, int capacity) {
  return phylo_fitch_set_tips_pitched(e, T, N, elt_bytes, n_states, codes, 0, weights, capacity);
}

extern "C" int phylo_fitch_set_tips_pitched(phylo_engine *e, int T, int64_t N, int elt_bytes, int n_states,
                                            const void *codes, uint64_t host_pitch_bytes, const double *weights,
                                            int capacity) {
  if (!e) return PHYLO_ERR_ARG;
  // elt_bytes == 0: the characters arrive in the device's own bit-sliced layout (phylo_fitch_pack_planes);
  // sets read back with phylo_fitch_get_states then use the narrowest element that holds n_states bits
  const bool planes_in = elt_bytes == 0;
  if (planes_in) {
    if (n_states < 1 || n_states > 64 || e->dSymTab)
      return fail(e, PHYLO_ERR_ARG, "fitch_set_tips: plane input needs 1 <= n_states <= 64 and no symbol table");
    elt_bytes = n_states <= 8 ? 1 : (n_states <= 16 ? 2 : (n_states <= 32 ? 4 : 8));
  }
  const size_t plane_row = (size_t)((N + 31) / 32) * np_device(n_states) * sizeof(uint32_t);
  if (host_pitch_bytes != 0 && N > 0 && host_pitch_bytes < (planes_in ? (uint64_t)plane_row : (uint64_t)N * (uint64_t)std::max(elt_bytes, 1)))
    return fail(e, PHYLO_ERR_ARG, "fitch_set_tips: host pitch %llu is shorter than a row of %lld codes",
                (unsigned long long)host_pitch_bytes, (long long)N);
  if (e->dSymTab && elt_bytes != 1)
    return fail(e, PHYLO_ERR_ARG, "fitch_set_tips: a symbol table is set, the characters must be 1 byte each (got %d)", elt_bytes);
  if (T < 1 || N < 1 || !codes || capacity < T ||
      !(elt_bytes == 1 || elt_bytes == 2 || elt_bytes == 4 || elt_bytes == 8) || n_states < 1 ||
      n_states > elt_bytes * 8)
    return fail(e, PHYLO_ERR_ARG, "fitch_set_tips: bad arguments (T=%d N=%lld elt_bytes=%d n_states=%d capacity=%d)",
                T, (long long)N, elt_bytes, n_states, capacity);
  std::vector<uint32_t> hw;
  if (weights) {
    hw.assign((size_t)((N + 31) / 32) * 32, 0u);
    for (int64_t i = 0; i < N; ++i) {
      const double w = weights[i];
      if (!(w >= 0.0) || w > 4294967295.0 || w != std::floor(w))
        return fail(e, PHYLO_ERR_UNSUPPORTED,
                    "fitch_set_tips: weight[%lld]=%g is not a non-negative integer < 2^32 (needed for exact lengths)",
                    (long long)i, w);
      hw[i] = (uint32_t)w;
    }
  }
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  // same shape as what is loaded: keep the plane buffers and tables, refresh contents only
  const bool reuse = e->fT == T && e->fN == N && e->fcap0 == capacity && e->felt == elt_bytes &&
                     e->fNP == n_states && e->dPreTab && (weights != nullptr) == (e->dFW != nullptr);
  if (reuse) {
    // the plane buffers stay allocated but what they hold belongs to the previous alignment
    std::fill(e->nodeCost.begin(), e->nodeCost.end(), 0);
    std::fill(e->fValid.begin(), e->fValid.end(), 0);
    std::fill(e->fFinValid.begin(), e->fFinValid.end(), 0);
  } else {
    fitch_free_data(e);
    e->fT = T; e->fN = N; e->fcap = capacity; e->fcap0 = capacity; e->felt = elt_bytes; e->fNP = n_states;
    e->fValid.assign(capacity, 0);
    e->fFinValid.assign(capacity, 0);
    e->fOwned.assign(capacity, 0);
    e->fFreeList.clear();
    e->fNext = T;
    e->fNPdev = np_device(n_states);
    e->fWords = (N + 31) / 32;
    e->fPre.assign(capacity, nullptr);
    e->fFin.assign(capacity, nullptr);
    e->nodeCost.assign(capacity, 0);
    CK(cudaMalloc(&e->dPreTab, sizeof(uint32_t *) * capacity));
    CK(cudaMalloc(&e->dFinTab, sizeof(uint32_t *) * capacity));
    // the tips' buffers: rows of one slab (padded to whole 32-word tiles like fitch_ensure's), so that an upload in
    // the device layout is two 2-D copies instead of T small ones
    const size_t tip_row_words = (size_t)((e->fWords + 31) / 32) * 32 * e->fNPdev;
    CK(cudaMalloc(&e->dTipSlab, sizeof(uint32_t) * tip_row_words * (size_t)T));
    CK(cudaMemsetAsync(e->dTipSlab, 0, sizeof(uint32_t) * tip_row_words * (size_t)T, e->stream));
    for (int t = 0; t < T; ++t) e->fPre[t] = e->dTipSlab + (size_t)t * tip_row_words;
    e->tabDirty = true;
  }
  int rc;
  if ((rc = fitch_cost_capacity(e, (size_t)capacity + 4)) != PHYLO_OK) return rc;
  // upload in chunks of whole taxa through the staging buffer, transcoding on device
  const size_t row = (size_t)N * elt_bytes;
  const size_t host_row = host_pitch_bytes ? (size_t)host_pitch_bytes : row;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, ((size_t)256 << 20) / row));
  if (!planes_in && (rc = fitch_stage(e, row * chunk)) != PHYLO_OK) return rc;
  CK(cudaMemsetAsync(e->dCost, 0, sizeof(unsigned long long), e->stream));
  if (planes_in) {
    // straight into the node buffers from two DMA queues, then one checking pass; nothing to transcode
    if (!e->copyStream) CK(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
    if (!e->copyStream2) CK(cudaStreamCreateWithFlags(&e->copyStream2, cudaStreamNonBlocking));
    for (int t = 0; t < T; ++t)
      if ((rc = fitch_ensure(e, t, false)) != PHYLO_OK) return rc;
    if ((rc = fitch_sync_tables(e)) != PHYLO_OK) return rc;  // also orders the copies after the allocation memsets
    while (e->slabEvents.size() < 2) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->slabEvents.push_back(ev);
    }
    const size_t hp = host_pitch_bytes ? (size_t)host_pitch_bytes : plane_row;
    const size_t dev_pitch = sizeof(uint32_t) * (size_t)((e->fWords + 31) / 32) * 32 * e->fNPdev;
    if (e->dTipSlab && e->fPre[0] == e->dTipSlab && plane_row * (size_t)T <= ((size_t)64 << 20)) {
      // small alignments: T copies of a few hundred KB each pay ~3 us of DMA set-up apiece; two 2-D copies of
      // T/2 rows (one per queue) do not
      const int h = (T + 1) / 2;
      CK(cudaMemcpy2DAsync(e->dTipSlab, dev_pitch, codes, hp, plane_row, (size_t)h, cudaMemcpyHostToDevice, e->copyStream));
      if (T > h)
        CK(cudaMemcpy2DAsync((char *)e->dTipSlab + (size_t)h * dev_pitch, dev_pitch, (const char *)codes + (size_t)h * hp, hp,
                             plane_row, (size_t)(T - h), cudaMemcpyHostToDevice, e->copyStream2));
    } else {
      for (int t = 0; t < T; ++t)
        CK(cudaMemcpyAsync(e->fPre[t], (const char *)codes + (size_t)t * hp, plane_row, cudaMemcpyHostToDevice,
                           (t & 1) ? e->copyStream2 : e->copyStream));
    }
    CK(cudaEventRecord(e->slabEvents[0], e->copyStream));
    CK(cudaEventRecord(e->slabEvents[1], e->copyStream2));
    CK(cudaStreamWaitEvent(e->stream, e->slabEvents[0], 0));
    CK(cudaStreamWaitEvent(e->stream, e->slabEvents[1], 0));
    {
      ProfScope prof(e, KC_FITCH_TRANSCODE);
      dim3 grid((unsigned)grid_for(e->fWords, 256, e->sm_count * 2), (unsigned)T);
      fitch_planes_check_kernel<<<grid, 256, 0, e->stream>>>(e->dPreTab, e->fWords, e->fN, e->fNPdev, e->dCost);
      LAUNCH_CHECK();
    }
  } else {
  // Within a chunk the H2D copy runs in pieces of ~8 MB on the copy stream while the pieces
  // that have landed are already being transcoded on the engine's stream.
  if (!e->copyStream) CK(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
  const int piece = (int)std::max<size_t>(1, ((size_t)8 << 20) / row);
  for (int t = 0; t < T; ++t)
    if ((rc = fitch_ensure(e, t, false)) != PHYLO_OK) return rc;  // allocations (and their memsets) up front
  {
    // the copy stream must not overtake work already queued on the engine's stream (memsets above)
    while (e->slabEvents.empty()) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->slabEvents.push_back(ev);
    }
    CK(cudaEventRecord(e->slabEvents[0], e->stream));
    CK(cudaStreamWaitEvent(e->copyStream, e->slabEvents[0], 0));
  }
  for (int t0 = 0; t0 < T; t0 += chunk) {
    const int nt = std::min(chunk, T - t0);
    const int npieces = (nt + piece - 1) / piece;
    while ((int)e->slabEvents.size() < npieces) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->slabEvents.push_back(ev);
    }
    for (int pc = 0; pc < npieces; ++pc) {
      const int a0 = pc * piece, a1 = std::min(nt, a0 + piece);
      if (host_row == row)
        CK(cudaMemcpyAsync((char *)e->dStage + (size_t)a0 * row, (const char *)codes + (size_t)(t0 + a0) * row,
                           row * (size_t)(a1 - a0), cudaMemcpyHostToDevice, e->copyStream));
      else  // a column slab of a wider host matrix (phylo_group shards)
        CK(cudaMemcpy2DAsync((char *)e->dStage + (size_t)a0 * row, row, (const char *)codes + (size_t)(t0 + a0) * host_row,
                             host_row, row, (size_t)(a1 - a0), cudaMemcpyHostToDevice, e->copyStream));
      CK(cudaEventRecord(e->slabEvents[pc], e->copyStream));
      CK(cudaStreamWaitEvent(e->stream, e->slabEvents[pc], 0));
      for (int t = a0; t < a1; ++t)
        if ((rc = fitch_encode(e, (const char *)e->dStage + (size_t)t * row, e->fPre[t0 + t], e->dCost)) != PHYLO_OK)
          return rc;
    }
    if (t0 + chunk < T) {  // staging buffer is reused by the next chunk
      CK(cudaStreamSynchronize(e->stream));
    }
  }
  }
  CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (e->hCost[0]) {
    const unsigned long long bad = e->hCost[0];
    fitch_free_data(e);
    return fail(e, PHYLO_ERR_DATA, "fitch_set_tips: %llu characters have an empty state set", bad);
  }
  if (weights) {
    if (!e->dFW) CK(cudaMalloc(&e->dFW, sizeof(uint32_t) * hw.size()));
    CK(cudaMemcpy(e->dFW, hw.data(), sizeof(uint32_t) * hw.size(), cudaMemcpyHostToDevice));
  }
  for (int t = 0; t < T; ++t) e->fValid[t] = 1;
  return PHYLO_OK;
}

// node-slot lifetime for the Fitch / bitvector sets: same contract as phylo_lk_node_alloc / _release
static int fitch_grow_slots(phylo_engine *e, int new_cap) {
  CK(cudaStreamSynchronize(e->stream));
  uint32_t **np = nullptr, **nf = nullptr;
  if (cudaMalloc(&np, sizeof(uint32_t *) * new_cap) != cudaSuccess || cudaMalloc(&nf, sizeof(uint32_t *) * new_cap) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(np); cudaFree(nf);
    return fail(e, PHYLO_ERR_CUDA, "fitch_node_alloc: cannot grow the slot tables to %d slots", new_cap);
  }
  dfree(e->dPreTab); dfree(e->dFinTab);
  e->dPreTab = np; e->dFinTab = nf;
  e->fPre.resize(new_cap, nullptr);
  e->fFin.resize(new_cap, nullptr);
  e->fValid.resize(new_cap, 0);
  e->fFinValid.resize(new_cap, 0);
  e->fOwned.resize(new_cap, 0);
  e->nodeCost.resize(new_cap, 0);
  e->fcap = new_cap;
  e->tabDirty = true;
  return PHYLO_OK;
}

extern "C" int phylo_fitch_node_alloc(phylo_engine *e, int *slot_out, uint64_t *generation_out) {
  if (!e) return PHYLO_ERR_ARG;
  if (!slot_out) return fail(e, PHYLO_ERR_ARG, "fitch_node_alloc: slot_out is NULL");
  if (e->fT == 0) return fail(e, PHYLO_ERR_STATE, "fitch_node_alloc: no Fitch data loaded");
  CK(cudaSetDevice(e->device));
  int slot = -1;
  if (!e->fFreeList.empty()) {
    slot = e->fFreeList.back();
    e->fFreeList.pop_back();
  } else {
    while (e->fNext < e->fcap && (e->fOwned[e->fNext] || e->fValid[e->fNext])) ++e->fNext;  // slots a schedule named directly
    if (e->fNext >= e->fcap) {
      const int rc = fitch_grow_slots(e, std::max(e->fcap * 2, e->fcap + 16));
      if (rc != PHYLO_OK) return rc;
    }
    slot = e->fNext++;
  }
  e->fOwned[slot] = 1;
  e->fValid[slot] = 0;
  e->fFinValid[slot] = 0;
  e->nodeCost[slot] = 0;
  *slot_out = slot;
  if (generation_out) *generation_out = e->fGen;
  return PHYLO_OK;
}

extern "C" int phylo_fitch_node_release(phylo_engine *e, int slot, uint64_t generation) {
  if (!e) return PHYLO_ERR_ARG;
  if (generation != e->fGen) return PHYLO_OK;  // the alignment this slot belonged to is gone already
  if (slot < e->fT || slot >= e->fcap || !e->fOwned[slot])
    return fail(e, PHYLO_ERR_ARG, "fitch_node_release: slot %d was not handed out by phylo_fitch_node_alloc", slot);
  e->fOwned[slot] = 0;
  e->fValid[slot] = 0;
  e->fFinValid[slot] = 0;
  e->fFreeList.push_back(slot);
  return PHYLO_OK;
}

extern "C" int phylo_fitch_node_stats(phylo_engine *e, int *capacity, int *in_use, int *with_buffers) {
  if (!e) return PHYLO_ERR_ARG;
  int used = 0, buf = 0;
  for (int s = e->fT; s < e->fcap; ++s) { used += e->fOwned[s] != 0; buf += e->fPre[s] != nullptr; }
  if (capacity) *capacity = e->fcap - e->fT;
  if (in_use) *in_use = used;
  if (with_buffers) *with_buffers = buf;
  return PHYLO_OK;
}

static int fitch_slot_ok(phylo_engine *e, int slot, bool need_data, const char *who) {
  if (e->fT == 0) return fail(e, PHYLO_ERR_STATE, "%s: no Fitch data loaded", who);
  if (slot < 0 || slot >= e->fcap) return fail(e, PHYLO_ERR_ARG, "%s: node slot %d out of range [0,%d)", who, slot, e->fcap);
  if (need_data && (!e->fPre[slot] || !e->fValid[slot]))
    return fail(e, PHYLO_ERR_STATE, "%s: node slot %d holds no state sets", who, slot);
  return PHYLO_OK;
}

static int fitch_pair(phylo_engine *e, int parent, int left, int right, bool store, uint64_t *out) {
  int rc;
  const char *who = store ? "fitch_median_2" : "fitch_distance";
  if ((rc = fitch_slot_ok(e, left, true, who)) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, right, true, who)) != PHYLO_OK) return rc;
  if (store) {
    if ((rc = fitch_slot_ok(e, parent, false, who)) != PHYLO_OK) return rc;
    if (parent < e->fT) return fail(e, PHYLO_ERR_ARG, "%s: parent slot %d is a tip", who, parent);
    if ((rc = fitch_ensure(e, parent, false)) != PHYLO_OK) return rc;
  }
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(e->dCost, 0, sizeof(unsigned long long), e->stream));
  const int g = grid_for(e->fWords, 256, e->sm_count * 8);
  uint32_t *c = store ? e->fPre[parent] : nullptr;
  const uint32_t *a = e->fPre[left], *b = e->fPre[right];
  {
  ProfScope prof(e, KC_FITCH_NODE);
  if (store) {
    NP_DISPATCH(e->fNPdev, (fitch_median2_kernel<NP, true><<<g, 256, 0, e->stream>>>(a, b, c, e->fWords, e->fN, e->dFW, e->dCost)));
  } else {
    NP_DISPATCH(e->fNPdev, (fitch_median2_kernel<NP, false><<<g, 256, 0, e->stream>>>(a, b, c, e->fWords, e->fN, e->dFW, e->dCost)));
  }
  LAUNCH_CHECK();
  }
  CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (out) *out = e->hCost[0];
  if (store) { e->nodeCost[parent] = e->hCost[0]; e->fValid[parent] = 1; e->fFinValid[parent] = 0; }
  return PHYLO_OK;
}

extern "C" int phylo_fitch_median_2(phylo_engine *e, int parent, int left, int right, uint64_t *cost_out) {
  if (!e) return PHYLO_ERR_ARG;
  return fitch_pair(e, parent, left, right, true, cost_out);
}

extern "C" int phylo_fitch_distance(phylo_engine *e, int a, int b, uint64_t *dist_out) {
  if (!e) return PHYLO_ERR_ARG;
  return fitch_pair(e, -1, a, b, false, dist_out);
}

// NonAdditive.median_3 (lib/nodeData.ml:22): final state set of a node from its parent's final set and
// the preliminary sets of the node itself and its two children (rule: fitch_final_rule). The result is
// written into dst's preliminary buffer, i.e. it is a node value like any other.
extern "C" int phylo_fitch_median_3(phylo_engine *e, int dst, int prelim, int parent_final, int left, int right) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  for (int s : {prelim, parent_final, left, right})
    if ((rc = fitch_slot_ok(e, s, true, "fitch_median_3")) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, dst, false, "fitch_median_3")) != PHYLO_OK) return rc;
  if (dst < e->fT) return fail(e, PHYLO_ERR_ARG, "fitch_median_3: dst slot %d is a tip", dst);
  if (dst == prelim || dst == parent_final || dst == left || dst == right)
    return fail(e, PHYLO_ERR_ARG, "fitch_median_3: dst aliases an operand");
  CK(cudaSetDevice(e->device));
  if ((rc = fitch_ensure(e, dst, false)) != PHYLO_OK) return rc;
  const int g = grid_for(e->fWords, 256, e->sm_count * 8);
  {
    ProfScope prof(e, KC_FITCH_UPPASS);
    NP_DISPATCH(e->fNPdev, (fitch_final1_kernel<NP><<<g, 256, 0, e->stream>>>(e->fPre[prelim], e->fPre[parent_final], e->fPre[left],
                                                                              e->fPre[right], e->fPre[dst], e->fWords)));
    LAUNCH_CHECK();
  }
  e->fValid[dst] = 1;
  e->fFinValid[dst] = 0;
  e->nodeCost[dst] = 0;
  return PHYLO_OK;
}

static int fitch_check_schedule(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                const char *who) {
  if (e->fT == 0) return fail(e, PHYLO_ERR_STATE, "%s: no Fitch data loaded", who);
  if (n_ops < 0 || (n_ops > 0 && !ops)) return fail(e, PHYLO_ERR_ARG, "%s: bad arguments", who);
  std::vector<char> ready(e->fcap, 0);
  for (int s = 0; s < e->fcap; ++s) ready[s] = e->fPre[s] != nullptr && e->fValid[s];  // earlier medians stay usable
  for (int o = 0; o < n_ops; ++o) {
    const phylo_op &op = ops[o];
    if (op.parent < e->fT || op.parent >= e->fcap)
      return fail(e, PHYLO_ERR_ARG, "%s: op %d parent slot %d must be in [%d,%d)", who, o, op.parent, e->fT, e->fcap);
    if (op.left < 0 || op.left >= e->fcap || op.right < 0 || op.right >= e->fcap || op.left == op.parent ||
        op.right == op.parent)
      return fail(e, PHYLO_ERR_ARG, "%s: op %d has bad child slots (%d,%d)", who, o, op.left, op.right);
    if (!ready[op.left] || !ready[op.right])
      return fail(e, PHYLO_ERR_ARG, "%s: op %d uses a child that is not computed yet (not post-order)", who, o);
    ready[op.parent] = 1;
  }
  if (root_a < 0 || root_a >= e->fcap || root_b < 0 || root_b >= e->fcap || !ready[root_a] || !ready[root_b])
    return fail(e, PHYLO_ERR_ARG, "%s: bad root edge (%d,%d)", who, root_a, root_b);
  return PHYLO_OK;
}

// Words (32 characters) per tree evaluation below which the latency-optimised register walk is used for
// 5..8 planes in auto mode, and below which the host spins on the tile kernel's mapped results instead of
// blocking in a stream sync. 4 planes: the tile kernel at every size (profiles/README.md: it also beats the
// L2 walk at 16 M and 64 M characters).
static const int64_t kFitchTileMaxWords = 1 << 18;
static const int64_t kFitchSpinMaxWords = 1 << 18;

// measurement only: PHYLO_FITCH_TIMING=1 makes the tile / warp-tile kernels leave %globaltimer stamps per CTA;
// the call prints where the time of the LAST call went (host and device clocks are separate: durations only).
static bool fitch_timing_on() {
  static const bool on = [] { const char *v = getenv("PHYLO_FITCH_TIMING"); return v && v[0] == '1'; }();
  return on;
}
static void fitch_timing_report(phylo_engine *e, unsigned long long *dStamps, int grid, const char *who, double host_plan_us,
                                double host_launch_us, double host_wait_us) {
  cudaStreamSynchronize(e->stream);
  std::vector<unsigned long long> h((size_t)grid * 8);
  cudaMemcpy(h.data(), dStamps, h.size() * 8, cudaMemcpyDeviceToHost);
  unsigned long long t0 = ~0ull;
  for (int b = 0; b < grid; ++b) t0 = std::min(t0, h[(size_t)b * 8]);
  static int calls = 0;
  if ((++calls % 8) != 0) return;
  fprintf(stderr, "[fitch timing %s] grid %d host: plan %.1f us, launch call %.1f us, wait %.1f us | device (us after first CTA start; min..max over CTAs):",
          who, grid, host_plan_us, host_launch_us, host_wait_us);
  const char *names[8] = {"start", "inputs", "phase1/steps", "phase2", "tiles done", "fold", "counter", "published"};
  for (int i = 0; i < 8; ++i) {
    unsigned long long lo = ~0ull, hi = 0;
    for (int b = 0; b < grid; ++b) {
      const unsigned long long t = h[(size_t)b * 8 + i];
      if (t == 0) continue;
      lo = std::min(lo, t); hi = std::max(hi, t);
    }
    if (hi) fprintf(stderr, " %s %.1f..%.1f", names[i], (lo - t0) / 1e3, (hi - t0) / 1e3);
  }
  fprintf(stderr, "\n");
}

// Fitch length is independent of where the tree is rooted (a median's cost is symmetric and the total is the
// minimum number of changes). For a length-only evaluation (PHYLO_OPT_RETAIN_CLV = 0) the schedule is re-rooted on
// the edge that minimises the height of the two halves: the chain of dependent medians -- what bounds a small
// alignment -- shrinks from the tree's height under the caller's root (25 for the 64-taxon bench tree) to about
// half its diameter. The nodes on the path between the new and the old root edge are re-used for the "rest of the
// tree" sets seen from below; every interior node still gets exactly one median. Returns false (schedule left
// alone) when a result is never consumed or consumed twice (not one tree).
static bool fitch_reroot_center(const phylo_op *ops, int n_ops, int ra, int rb, int cap, std::vector<phylo_op> &out,
                                int *nra, int *nrb) {
  if (n_ops < 2 || ra == rb) return false;
  std::vector<int> prod(cap, -1), up(cap, -1), hd(cap, 0), hu(cap, 0);
  for (int o = 0; o < n_ops; ++o) {
    const phylo_op &op = ops[o];
    if (prod[op.parent] >= 0 || up[op.left] >= 0 || up[op.right] >= 0 || op.left == op.right) return false;
    if (prod[op.left] > o || prod[op.right] > o) return false;
    prod[op.parent] = o;
    up[op.left] = op.parent;
    up[op.right] = op.parent;
    hd[op.parent] = 1 + std::max(hd[op.left], hd[op.right]);
  }
  if (up[ra] >= 0 || up[rb] >= 0) return false;
  for (int o = 0; o < n_ops; ++o)
    if (up[ops[o].parent] < 0 && ops[o].parent != ra && ops[o].parent != rb) return false;  // a result nobody consumes
  up[ra] = rb;
  up[rb] = ra;
  hu[ra] = hd[rb];
  hu[rb] = hd[ra];
  int best = ra, best_h = std::max(hd[ra], hu[ra]);
  for (int o = n_ops - 1; o >= 0; --o) {  // parents before children
    const phylo_op &op = ops[o];
    hu[op.left] = 1 + std::max(hu[op.parent], hd[op.right]);
    hu[op.right] = 1 + std::max(hu[op.parent], hd[op.left]);
    for (int c : {op.left, op.right}) {
      const int h = std::max(hd[c], hu[c]);
      if (h < best_h) { best_h = h; best = c; }
    }
  }
  if (best == ra || best == rb) return false;  // already rooted as well as it can be
  // path best -> p1 -> ... -> pk (pk = ra or rb)
  std::vector<int> path;
  std::vector<char> on_path(cap, 0);
  for (int v = up[best]; ; v = up[v]) {
    path.push_back(v);
    on_path[v] = 1;
    if (v == ra || v == rb) break;
  }
  out.clear();
  out.reserve(n_ops);
  for (int o = 0; o < n_ops; ++o)
    if (!on_path[ops[o].parent]) out.push_back(ops[o]);
  const int top = path.back(), other = top == ra ? rb : ra;
  int below = path.size() >= 2 ? path[path.size() - 2] : best;  // the path's child of `top`
  {
    const phylo_op &op = ops[prod[top]];
    phylo_op n = op;
    n.parent = top;
    n.left = other;
    n.right = op.left == below ? op.right : op.left;
    out.push_back(n);
  }
  for (int i = (int)path.size() - 2; i >= 0; --i) {
    const int p = path[i];
    below = i >= 1 ? path[i - 1] : best;
    const phylo_op &op = ops[prod[p]];
    phylo_op n = op;
    n.parent = p;
    n.left = path[i + 1];  // holds the set of everything above p
    n.right = op.left == below ? op.right : op.left;
    out.push_back(n);
  }
  *nra = best;
  *nrb = path[0];
  return (int)out.size() == n_ops;
}

// Host-only view of the re-rooting for tests (no CUDA call).
extern "C" int phylo_fitch_reroot(const phylo_op *ops, int n_ops, int capacity, int root_a, int root_b, phylo_op *ops_out,
                                  int *root_a_out, int *root_b_out) {
  if (n_ops < 0 || (n_ops > 0 && !ops) || capacity < 2 || !ops_out || !root_a_out || !root_b_out) return PHYLO_ERR_ARG;
  if (root_a < 0 || root_a >= capacity || root_b < 0 || root_b >= capacity) return PHYLO_ERR_ARG;
  for (int o = 0; o < n_ops; ++o)
    if (ops[o].parent < 0 || ops[o].parent >= capacity || ops[o].left < 0 || ops[o].left >= capacity || ops[o].right < 0 ||
        ops[o].right >= capacity)
      return PHYLO_ERR_ARG;
  std::vector<phylo_op> out;
  int a = root_a, b = root_b;
  if (!fitch_reroot_center(ops, n_ops, root_a, root_b, capacity, out, &a, &b)) {
    for (int o = 0; o < n_ops; ++o) ops_out[o] = ops[o];
    *root_a_out = root_a;
    *root_b_out = root_b;
    return PHYLO_OK;
  }
  for (int o = 0; o < n_ops; ++o) ops_out[o] = out[o];
  *root_a_out = a;
  *root_b_out = b;
  return PHYLO_OK;
}

// On-chip tile kernel (fitch_tile_kernel): per-warp subtree lists + the medians above the cut,
// results straight into mapped host memory. Returns PHYLO_OK with *done = false when the
// schedule does not fit. The compiled program of the last call is kept: scoring the same
// schedule again (another character set on the same tree, new weights, a benchmark loop) skips
// the compilation after a memcmp of the ops and a check of every buffer pointer it refers to.
static int fitch_tile_compile(phylo_engine *e, const phylo_op *ops_in, int n_ops, int root_a_in, int root_b_in, bool retain,
                              bool *ok) {
  *ok = false;
  FitchTileCache &tc = e->tileCache;
  tc.valid = false;
  const int n_tot = n_ops + 1;  // + root-edge join
  // length only: no set leaves the chip, so the tree may be evaluated from the edge that makes it shallowest
  const phylo_op *ops = ops_in;
  int root_a = root_a_in, root_b = root_b_in;
  std::vector<phylo_op> rerooted;
  if (!retain && fitch_reroot_center(ops_in, n_ops, root_a_in, root_b_in, e->fcap, rerooted, &root_a, &root_b)) ops = rerooted.data();
  // operands: >= 0 = index into the tile's input rows; < 0 = -1 - (op that produces it)
  std::vector<int> produced(e->fcap, -1), inputs;
  struct Raw { int l, r, out_slot; };
  std::vector<Raw> raw(n_tot);
  auto operand = [&](int slot) {  // every use of a resident set gets its own table row (results overwrite rows)
    if (produced[slot] >= 0) return -1 - produced[slot];
    inputs.push_back(slot);
    return (int)inputs.size() - 1;
  };
  std::vector<int> size(n_tot, 1), consumer(n_tot, -1);
  for (int o = 0; o < n_tot; ++o) {
    raw[o].l = operand(o < n_ops ? ops[o].left : root_a);
    raw[o].r = operand(o < n_ops ? ops[o].right : root_b);
    raw[o].out_slot = o < n_ops ? ops[o].parent : -1;
    for (int c : {raw[o].l, raw[o].r})
      if (c < 0) {
        if (consumer[-1 - c] >= 0) return PHYLO_OK;  // a result used twice: not a forest, other kernels handle it
        consumer[-1 - c] = o;
        size[o] += size[-1 - c];
      }
    if (o < n_ops) produced[ops[o].parent] = o;
  }
  const int n_in = (int)inputs.size();
  const bool weighted = e->dFW != nullptr;
  const size_t smem = (size_t)n_in * 512 + (size_t)(n_tot + 4) * (sizeof(FitchTileOp) + (weighted ? 256 : 64)) +
                      (size_t)n_in * 8;
  if (smem > 200 * 1024) return PHYLO_OK;
  // Cut the forest into whole subtrees of at most `cap` medians (phase 1, dealt to the warps,
  // largest first to the least loaded warp); what is above the cut runs on warp 0 (phase 2).
  const int cap = std::max(4, (n_tot + kFitchTileWarps - 1) / kFitchTileWarps);
  std::vector<int> task_of(n_tot, -1);  // -1: phase 2; else the phase-1 task (= its top op)
  for (int o = n_tot - 1; o >= 0; --o) {  // consumers have larger indices: parents are visited first
    if (consumer[o] >= 0 && task_of[consumer[o]] >= 0) task_of[o] = task_of[consumer[o]];
    else if (size[o] <= cap) task_of[o] = o;
  }
  std::vector<int> tasks;
  for (int o = 0; o < n_tot; ++o)
    if (task_of[o] == o) tasks.push_back(o);
  std::sort(tasks.begin(), tasks.end(), [&](int x, int y) { return size[x] != size[y] ? size[x] > size[y] : x < y; });
  std::vector<int> load(kFitchTileWarps, 0), warp_of_task(n_tot, 0);
  for (int t : tasks) {
    const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    warp_of_task[t] = w;
    load[w] += size[t];
  }
  // program order: phase-1 ops of warp 0, 1, ... (each in schedule order), then phase 2
  std::vector<int> order, tstart(kFitchTileWarps + 2, 0);
  tc.pos.assign(n_tot, 0);
  order.reserve(n_tot);
  for (int w = 0; w <= kFitchTileWarps; ++w) {
    tstart[w] = (int)order.size();
    for (int o = 0; o < n_tot; ++o) {
      const bool mine = w < kFitchTileWarps ? (task_of[o] >= 0 && warp_of_task[task_of[o]] == w) : task_of[o] < 0;
      if (mine) { tc.pos[o] = (int)order.size(); order.push_back(o); }
    }
  }
  tstart[kFitchTileWarps + 1] = (int)order.size();
  // An operand produced by the median the same warp evaluated just before stays in registers: it
  // becomes the left operand (the rule is symmetric) and is flagged in bit 0 of l_off.
  std::vector<char> from_regs(n_tot, 0);
  for (int w = 0; w <= kFitchTileWarps; ++w)
    for (int i = tstart[w] + 1; i < tstart[w + 1]; ++i) {
      Raw &rw = raw[order[i]];
      const int prev = -1 - order[i - 1];
      if (rw.r == prev) std::swap(rw.l, rw.r);
      if (rw.l == prev) from_regs[order[i]] = 1;
    }
  int rc;
  const size_t blob = sizeof(FitchTileOp) * n_tot + 8 * (size_t)n_in + 4 * (size_t)(kFitchTileWarps + 2) + 64;
  if ((rc = fitch_sched_capacity(e, blob)) != PHYLO_OK) return rc;
  if ((rc = fitch_cost_capacity(e, (size_t)n_tot + 8)) != PHYLO_OK) return rc;
  if ((size_t)n_tot + 2 > e->capAcc) {
    CK(cudaStreamSynchronize(e->stream));
    dfree(e->dAcc);
    e->capAcc = 0;
    const size_t cap2 = ((size_t)n_tot + 2) * 2;
    CK(cudaMalloc(&e->dAcc, sizeof(unsigned long long) * cap2 * kFitchAccCopies));
    CK(cudaMemsetAsync(e->dAcc, 0, sizeof(unsigned long long) * cap2 * kFitchAccCopies, e->stream));
    e->capAcc = cap2;
  }
  const bool inl = blob - 64 <= (size_t)kFitchInlineProg;
  FitchTileArgs &a = tc.a;
  if (!inl) CK(cudaStreamSynchronize(e->stream));  // pinned staging is about to be rewritten
  char *hb = inl ? (char *)a.prog : (char *)e->hSched;
  FitchTileOp *hops = (FitchTileOp *)hb;
  const uint32_t **hin = (const uint32_t **)(hb + sizeof(FitchTileOp) * n_tot);
  int *hlev = (int *)(hb + sizeof(FitchTileOp) * n_tot + 8 * (size_t)n_in);
  std::vector<int> res_row(n_tot, -1);  // table row holding op o's result (schedule order: producers first)
  auto row_of = [&](int c) { return c >= 0 ? c : res_row[-1 - c]; };
  for (int o = 0; o < n_tot; ++o) res_row[o] = row_of(raw[o].l);
  tc.slots.clear();
  tc.ptrs.clear();
  auto remember = [&](int slot) { tc.slots.push_back(slot); tc.ptrs.push_back(e->fPre[slot]); };
  for (int i = 0; i < n_tot; ++i) {
    const Raw &rw = raw[order[i]];
    // rows: an input use has its own row; a result lives in the row of its producer's left operand
    hops[i].l_off = 512u * (uint32_t)row_of(rw.l) | (from_regs[order[i]] ? 1u : 0u);
    hops[i].r_off = 512u * (uint32_t)row_of(rw.r);
    hops[i].out = (retain && rw.out_slot >= 0) ? e->fPre[rw.out_slot] : nullptr;
    if (retain && rw.out_slot >= 0) remember(rw.out_slot);
  }
  for (int i = 0; i < n_in; ++i) { hin[i] = e->fPre[inputs[i]]; remember(inputs[i]); }
  for (int w = 0; w < kFitchTileWarps + 2; ++w) hlev[w] = tstart[w];
  if (!inl) CK(cudaMemcpyAsync(e->dSched, hb, blob - 64, cudaMemcpyHostToDevice, e->stream));
  a.inline_prog = inl;
  char *db = (char *)e->dSched;
  a.ops = (const FitchTileOp *)db;
  a.in_ptr = (const uint32_t *const *)(db + sizeof(FitchTileOp) * n_tot);
  a.task_start = (const int *)(db + sizeof(FitchTileOp) * n_tot + 8 * (size_t)n_in);
  a.n_in = n_in; a.n_ops = n_tot;
  a.nwords = e->fWords; a.N = e->fN; a.wt = e->dFW;
  a.acc = e->dAcc;
  if (!e->hCostDev) CK(cudaHostGetDevicePointer((void **)&e->hCostDev, e->hCost, 0));
  a.host_out = e->hCostDev;
  a.stamps = nullptr;
  tc.smem = smem;
  tc.weighted = weighted;
  tc.ops.assign(ops_in, ops_in + n_ops);
  tc.root_a = root_a_in; tc.root_b = root_b_in;
  tc.retain = retain;
  tc.valid = inl;  // a program staged through dSched is overwritten by other calls: compiled again next time
  *ok = true;
  return PHYLO_OK;
}

static bool fitch_tile_cache_hit(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b, bool retain) {
  const FitchTileCache &tc = e->tileCache;
  if (!tc.valid || (int)tc.ops.size() != n_ops || tc.root_a != root_a || tc.root_b != root_b || tc.retain != retain) return false;
  if (tc.a.nwords != e->fWords || tc.a.N != e->fN || tc.a.wt != e->dFW || tc.a.acc != e->dAcc || tc.a.host_out != e->hCostDev ||
      e->hCostDev == nullptr)
    return false;
  for (int o = 0; o < n_ops; ++o)  // field by field: the struct has padding
    if (ops[o].parent != tc.ops[o].parent || ops[o].left != tc.ops[o].left || ops[o].right != tc.ops[o].right) return false;
  for (size_t i = 0; i < tc.slots.size(); ++i)
    if (e->fPre[tc.slots[i]] != tc.ptrs[i]) return false;
  return true;
}

static int fitch_score_tree_tile(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b, bool retain,
                                 uint64_t *length_out, bool *done) {
  *done = false;
  const auto ht0 = std::chrono::steady_clock::now();
  auto ht1 = ht0, ht2 = ht0;
  int grid_used = 0, rc;
  if (!fitch_tile_cache_hit(e, ops, n_ops, root_a, root_b, retain)) {
    bool ok = false;
    if ((rc = fitch_tile_compile(e, ops, n_ops, root_a, root_b, retain, &ok)) != PHYLO_OK) return rc;
    if (!ok) return PHYLO_OK;
  }
  FitchTileCache &tc = e->tileCache;
  FitchTileArgs &a = tc.a;
  const int n_tot = n_ops + 1;
  const bool weighted = tc.weighted;
  const size_t smem = tc.smem;
  // unweighted costs fit 48 bits: every published word carries the call's 16-bit tag, the host waits for the
  // tags (no ordering between words needed); weighted costs may need 64 bits: results, fence, sequence number
  a.seq = ++e->tileSeq;
  a.tag = weighted ? 0ull : (((a.seq & 0x7fffull) | 0x8000ull) << 48);
  if (a.tag) for (int i = 0; i < n_tot; ++i) e->hCost[i] = 0;
  e->hCost[n_tot + 1] = ~0ull;  // untagged protocol: the slot the last CTA overwrites with a.seq
  a.stamps = nullptr;
  if (fitch_timing_on()) {
    if (!e->dStamps) CK(cudaMalloc(&e->dStamps, 8 * 8 * 8192));
    CK(cudaMemsetAsync(e->dStamps, 0, 8 * 8 * 8192, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    a.stamps = e->dStamps;
  }
  ht1 = std::chrono::steady_clock::now();
  {
    ProfScope prof(e, KC_FITCH_TREE);
    auto kern = weighted ? fitch_tile_kernel<unsigned long long> : fitch_tile_kernel<uint16_t>;
    if (e->tileSmem != smem || e->tileWeighted != weighted) {  // attribute + occupancy are looked up once per program shape
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int q = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, 256, smem) != cudaSuccess || q < 1) q = 1;
      e->tileSmem = smem;
      e->tileWeighted = weighted;
      e->tileOcc = q;
    }
    // whole waves: every CTA gets the same number of tiles (+-1)
    const int64_t ntiles = (e->fWords + 31) / 32, maxg = (int64_t)e->sm_count * e->tileOcc;
    const int64_t waves = (ntiles + maxg - 1) / maxg;
    const int g = (int)((ntiles + waves - 1) / waves);
    if (!weighted && waves > 2000) return PHYLO_OK;  // 16-bit per-lane counters: the other kernels take it
    a.prefetch_next = (waves >= 2 && waves <= 4) ? 1 : 0;
    kern<<<g, 256, smem, e->stream>>>(a);
    LAUNCH_CHECK();
    grid_used = g;
    ht2 = std::chrono::steady_clock::now();
  }
  // small alignments: spin on what the last CTA writes into mapped host memory (a blocking stream
  // sync costs more than the kernel); otherwise, or after 2 ms, a stream sync
  {
    volatile unsigned long long *hc = (volatile unsigned long long *)e->hCost;
    const unsigned long long want = a.tag >> 48;
    auto published = [&]() {
      if (!a.tag) return hc[n_tot + 1] == a.seq;
      for (int i = 0; i < n_tot; ++i)
        if ((hc[i] >> 48) != want) return false;
      return true;
    };
    bool seen = false;
    if (e->fWords <= kFitchSpinMaxWords) {
      const auto t0 = std::chrono::steady_clock::now();
      for (int spin = 0;; ++spin) {
        if (published()) { seen = true; break; }
        if ((spin & 1023) == 1023 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
      }
    }
    if (!seen) {
      CK(cudaStreamSynchronize(e->stream));
      if (!published()) return fail(e, PHYLO_ERR_CUDA, "fitch_score_tree: the tile kernel did not publish its result");
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  const unsigned long long mask = a.tag ? 0x0000ffffffffffffull : ~0ull;
  uint64_t length = 0;
  for (int i = 0; i < n_tot; ++i) length += e->hCost[i] & mask;
  *length_out = length;
  if (retain) {
    for (int o = 0; o < n_ops; ++o) { e->nodeCost[ops[o].parent] = e->hCost[tc.pos[o]] & mask; e->fValid[ops[o].parent] = 1; e->fFinValid[ops[o].parent] = 0; }
  } else {  // length only: the parents' sets were not written (and the medians may belong to another rooting)
    for (int o = 0; o < n_ops; ++o) { e->nodeCost[ops[o].parent] = 0; e->fValid[ops[o].parent] = 0; e->fFinValid[ops[o].parent] = 0; }
  }
  if (a.stamps) {
    const auto ht3 = std::chrono::steady_clock::now();
    auto us = [](auto x, auto y) { return std::chrono::duration<double, std::micro>(y - x).count(); };
    fitch_timing_report(e, e->dStamps, std::min(grid_used, 8192), "tile", us(ht0, ht1), us(ht1, ht2), us(ht2, ht3));
  }
  if (e->prof_on) prof_resolve_lazy(e);
  *done = true;
  return PHYLO_OK;
}

extern "C" int phylo_fitch_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                      uint64_t *length_out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = fitch_check_schedule(e, ops, n_ops, root_a, root_b, "fitch_score_tree")) != PHYLO_OK) return rc;
  if (!length_out) return fail(e, PHYLO_ERR_ARG, "fitch_score_tree: length_out is NULL");
  CK(cudaSetDevice(e->device));
  const bool tile_ok = e->fNPdev == 4 && (e->opt_fitch_walk == 3 || e->opt_fitch_walk == 1);
  if (tile_ok && !e->opt_retain) {  // PHYLO_OPT_RETAIN_CLV = 0: the length only; no interior set is written
    bool done = false;
    if ((rc = fitch_score_tree_tile(e, ops, n_ops, root_a, root_b, false, length_out, &done)) != PHYLO_OK) return rc;
    if (done) return PHYLO_OK;
  }
  for (int o = 0; o < n_ops; ++o)
    if ((rc = fitch_ensure(e, ops[o].parent, false)) != PHYLO_OK) return rc;
  if (tile_ok) {
    bool done = false;
    if ((rc = fitch_score_tree_tile(e, ops, n_ops, root_a, root_b, true, length_out, &done)) != PHYLO_OK) return rc;
    if (done) return PHYLO_OK;
  }
  if ((rc = fitch_sync_tables(e)) != PHYLO_OK) return rc;
  if ((rc = fitch_cost_capacity(e, (size_t)n_ops + 4)) != PHYLO_OK) return rc;
  if ((rc = fitch_sched_capacity(e, sizeof(FitchInstr) * (size_t)(n_ops + 2))) != PHYLO_OK)
    return rc;
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemsetAsync(e->dCost, 0, sizeof(unsigned long long) * (size_t)(n_ops + 2), e->stream));

  // register-walk kernel: compiled depth-first plan, tips prefetched D steps ahead
  FusedPlan pl;
  // auto: the register walk only pays where the dependent L2 walk is latency-bound (small alignments)
  const bool walk = (e->opt_fitch_walk == 2 || (e->opt_fitch_walk == 1 && e->fWords <= kFitchTileMaxWords)) && e->fNPdev <= 8 &&
                    build_fused_plan(e->fcap, e->fT, ops, n_ops, root_a, root_b, 0.0, pl);
  size_t smem_walk = 0;
  if (walk) {
    const size_t ns = pl.steps.size();
    smem_walk = ns * sizeof(FitchInstr) + (ns + (ns & 1)) * sizeof(unsigned long long) +
                (size_t)pl.depth * e->fNPdev * 128 * sizeof(uint32_t);
  }
  if (walk && smem_walk <= 200 * 1024) {
    const int ns = (int)pl.steps.size();
    FitchInstr *hp = (FitchInstr *)e->hSched;  // capacity: sizeof(FitchInstr) * (n_ops + 2) ensured below
    for (int i = 0; i < ns; ++i) {
      const PlanStep &st = pl.steps[i];
      FitchInstr in{};
      const int lk = (st.lkind == OPK_STORED) ? OPK_TIP : st.lkind, rk = (st.rkind == OPK_STORED) ? OPK_TIP : st.rkind;
      in.kinds = lk | (rk << 2) | (st.push_first << 4);
      in.l = (lk == OPK_TIP) ? e->fPre[st.lidx] : nullptr;
      in.r = (rk == OPK_TIP) ? e->fPre[st.ridx] : nullptr;
      in.out = st.out_slot >= 0 ? e->fPre[st.out_slot] : nullptr;
      hp[i] = in;
    }
    CK(cudaMemcpyAsync(e->dSched, hp, sizeof(FitchInstr) * (size_t)ns, cudaMemcpyHostToDevice, e->stream));
    const int g = grid_for(e->fWords, 128, e->sm_count * 16);
    {
      ProfScope prof(e, KC_FITCH_TREE);
#define FITCH_WALK(NPV, DV)                                                                                         \
  {                                                                                                                 \
    auto kern = fitch_treep_kernel<NPV, DV>;                                                                        \
    if (smem_walk > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_walk)); \
    kern<<<g, 128, smem_walk, e->stream>>>((const FitchInstr *)e->dSched, ns, pl.depth, e->fWords, e->fN, e->dFW,   \
                                           e->dCost, e->dCost + n_ops + 1);                                        \
  }
      switch (e->fNPdev) {
        case 1: FITCH_WALK(1, 8) break;
        case 2: FITCH_WALK(2, 8) break;
        case 3: FITCH_WALK(3, 8) break;
        case 4: FITCH_WALK(4, 8) break;
        case 5: FITCH_WALK(5, 4) break;
        case 6: FITCH_WALK(6, 4) break;
        default: FITCH_WALK(8, 4) break;
      }
#undef FITCH_WALK
      ++e->launches;
      CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long) * (size_t)(n_ops + 2), cudaMemcpyDeviceToHost,
                       e->stream));
    CK(cudaStreamSynchronize(e->stream));
    *length_out = e->hCost[n_ops + 1];
    for (int i = 0; i < ns; ++i)
      if (pl.steps[i].out_slot >= 0) { e->nodeCost[pl.steps[i].out_slot] = e->hCost[i]; e->fValid[pl.steps[i].out_slot] = 1; e->fFinValid[pl.steps[i].out_slot] = 0; }
    if (e->prof_on) prof_resolve_lazy(e);
    return PHYLO_OK;
  }

  FitchStep *hs = (FitchStep *)e->hSched;
  for (int o = 0; o < n_ops; ++o) hs[o] = FitchStep{ops[o].parent, ops[o].left, ops[o].right};
  CK(cudaMemcpyAsync(e->dSched, hs, sizeof(FitchStep) * (size_t)n_ops, cudaMemcpyHostToDevice, e->stream));
  const int g = grid_for(e->fWords, 128, e->sm_count * 16);
  const size_t smem2 = sizeof(unsigned long long) * (size_t)(n_ops + 1);
  {
  ProfScope prof(e, KC_FITCH_TREE);
  NP_DISPATCH(e->fNPdev, (fitch_tree_kernel<NP><<<g, 128, smem2, e->stream>>>(
                             e->dPreTab, (const FitchStep *)e->dSched, n_ops, root_a, root_b, e->fT, e->fWords, e->fN,
                             e->dFW, e->dCost, e->dCost + n_ops + 1)));
  LAUNCH_CHECK();
  }
  CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long) * (size_t)(n_ops + 2), cudaMemcpyDeviceToHost,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *length_out = e->hCost[n_ops + 1];
  for (int o = 0; o < n_ops; ++o) { e->nodeCost[ops[o].parent] = e->hCost[o]; e->fValid[ops[o].parent] = 1; e->fFinValid[ops[o].parent] = 0; }
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

// ------------------------------------------- general-TCM median (CostMatrix, lib/costMatrix.ml) ----
extern "C" int phylo_tcm_set_matrix(phylo_engine *e, int n_states, const int32_t *M, int metric) {
  if (!e) return PHYLO_ERR_ARG;
  if (n_states < 2 || n_states > 6 || !M) return fail(e, PHYLO_ERR_ARG, "tcm_set_matrix: 2 <= n_states <= 6 and a matrix are required");
  const int S = n_states, sets = 1 << S;
  for (int i = 0; i < S * S; ++i)
    if (M[i] < 0 || M[i] > 30000) return fail(e, PHYLO_ERR_ARG, "tcm_set_matrix: costs must be in [0, 30000]");
  std::vector<uint32_t> tab((size_t)sets * sets, 0);
  for (int a = 1; a < sets; ++a)
    for (int b = 1; b < sets; ++b) {
      // find_median_general (lib/costMatrix.ml:68-86) / find_median_metric (:107-124): the best k over
      // all (istate, jstate) pairs is the best k for the cheapest istate and the cheapest jstate
      int best = INT32_MAX;
      uint32_t med = 0;
      for (int k = 0; k < S; ++k) {
        if (metric && !(((a | b) >> k) & 1)) continue;
        int ca = INT32_MAX, cb = INT32_MAX;
        for (int i = 0; i < S; ++i) {
          if ((a >> i) & 1) ca = std::min(ca, M[i * S + k]);
          if ((b >> i) & 1) cb = std::min(cb, M[i * S + k]);
        }
        const int c = ca + cb;
        if (c < best) { best = c; med = 1u << k; }
        else if (c == best) med |= 1u << k;
      }
      tab[(size_t)a * sets + b] = (uint32_t)best | (med << 16);
    }
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  dfree(e->dTcm);
  CK(cudaMalloc(&e->dTcm, sizeof(uint32_t) * tab.size()));
  CK(cudaMemcpy(e->dTcm, tab.data(), sizeof(uint32_t) * tab.size(), cudaMemcpyHostToDevice));
  e->tcmS = S;
  return PHYLO_OK;
}

// one median (store) or the cost only (parent < 0) on the engine's stream; adds into dCost[idx]
static int tcm_launch(phylo_engine *e, int parent, int left, int right, size_t idx) {
  const int g = grid_for(e->fWords, 256, e->sm_count * 8);
  const size_t smem = sizeof(uint32_t) << (2 * e->tcmS);
  uint32_t *c = parent >= 0 ? e->fPre[parent] : nullptr;
  const uint32_t *a = e->fPre[left], *b = e->fPre[right];
  ProfScope prof(e, KC_FITCH_NODE);
  if (parent >= 0) {
    NP_DISPATCH(e->fNPdev, (tcm_median2_kernel<NP, true><<<g, 256, smem, e->stream>>>(a, b, c, e->fWords, e->fN, e->tcmS, e->dTcm, e->dFW, e->dCost + idx)));
  } else {
    NP_DISPATCH(e->fNPdev, (tcm_median2_kernel<NP, false><<<g, 256, smem, e->stream>>>(a, b, c, e->fWords, e->fN, e->tcmS, e->dTcm, e->dFW, e->dCost + idx)));
  }
  LAUNCH_CHECK();
  return PHYLO_OK;
}

static int tcm_ready(phylo_engine *e, const char *who) {
  if (e->fT == 0) return fail(e, PHYLO_ERR_STATE, "%s: no Fitch data loaded", who);
  if (!e->dTcm) return fail(e, PHYLO_ERR_STATE, "%s: call phylo_tcm_set_matrix first", who);
  if (e->tcmS != e->fNP) return fail(e, PHYLO_ERR_STATE, "%s: the cost matrix has %d states, the characters %d", who, e->tcmS, e->fNP);
  return PHYLO_OK;
}

extern "C" int phylo_tcm_median_2(phylo_engine *e, int parent, int left, int right, uint64_t *cost_out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = tcm_ready(e, "tcm_median_2")) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, left, true, "tcm_median_2")) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, right, true, "tcm_median_2")) != PHYLO_OK) return rc;
  if (parent >= 0) {
    if ((rc = fitch_slot_ok(e, parent, false, "tcm_median_2")) != PHYLO_OK) return rc;
    if (parent < e->fT) return fail(e, PHYLO_ERR_ARG, "tcm_median_2: parent slot %d is a tip", parent);
    if ((rc = fitch_ensure(e, parent, false)) != PHYLO_OK) return rc;
  }
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(e->dCost, 0, sizeof(unsigned long long), e->stream));
  if ((rc = tcm_launch(e, parent, left, right, 0)) != PHYLO_OK) return rc;
  CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (cost_out) *cost_out = e->hCost[0];
  if (parent >= 0) { e->nodeCost[parent] = e->hCost[0]; e->fValid[parent] = 1; e->fFinValid[parent] = 0; }
  return PHYLO_OK;
}

extern "C" int phylo_tcm_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                    uint64_t *length_out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = tcm_ready(e, "tcm_score_tree")) != PHYLO_OK) return rc;
  if ((rc = fitch_check_schedule(e, ops, n_ops, root_a, root_b, "tcm_score_tree")) != PHYLO_OK) return rc;
  if (!length_out) return fail(e, PHYLO_ERR_ARG, "tcm_score_tree: length_out is NULL");
  CK(cudaSetDevice(e->device));
  for (int o = 0; o < n_ops; ++o)
    if ((rc = fitch_ensure(e, ops[o].parent, false)) != PHYLO_OK) return rc;
  if ((rc = fitch_cost_capacity(e, (size_t)n_ops + 8)) != PHYLO_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemsetAsync(e->dCost, 0, sizeof(unsigned long long) * (size_t)(n_ops + 2), e->stream));
  for (int o = 0; o < n_ops; ++o)
    if ((rc = tcm_launch(e, ops[o].parent, ops[o].left, ops[o].right, (size_t)o)) != PHYLO_OK) return rc;
  if ((rc = tcm_launch(e, -1, root_a, root_b, (size_t)n_ops)) != PHYLO_OK) return rc;
  CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long) * (size_t)(n_ops + 1), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  uint64_t total = 0;
  for (int o = 0; o <= n_ops; ++o) total += e->hCost[o];
  for (int o = 0; o < n_ops; ++o) { e->nodeCost[ops[o].parent] = e->hCost[o]; e->fValid[ops[o].parent] = 1; e->fFinValid[ops[o].parent] = 0; }
  *length_out = total;
  if (e->prof_on) prof_resolve_lazy(e);
  return PHYLO_OK;
}

extern "C" int phylo_fitch_get_node_costs(phylo_engine *e, uint64_t *out) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->fT == 0 || !out) return fail(e, PHYLO_ERR_STATE, "fitch_get_node_costs: no Fitch data loaded");
  for (int s = 0; s < e->fcap; ++s) out[s] = e->nodeCost[s];
  return PHYLO_OK;
}

extern "C" int phylo_fitch_uppass(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if (e->fT == 0) return fail(e, PHYLO_ERR_STATE, "fitch_uppass: no Fitch data loaded");
  if (n_ops < 0 || (n_ops > 0 && !ops)) return fail(e, PHYLO_ERR_ARG, "fitch_uppass: bad arguments");
  if ((rc = fitch_slot_ok(e, root_a, true, "fitch_uppass")) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, root_b, true, "fitch_uppass")) != PHYLO_OK) return rc;
  std::vector<int> parent_of(e->fcap, -1);
  for (int o = 0; o < n_ops; ++o) {
    const phylo_op &op = ops[o];
    if ((rc = fitch_slot_ok(e, op.parent, true, "fitch_uppass (run the down-pass first)")) != PHYLO_OK) return rc;
    if ((rc = fitch_slot_ok(e, op.left, true, "fitch_uppass")) != PHYLO_OK) return rc;
    if ((rc = fitch_slot_ok(e, op.right, true, "fitch_uppass")) != PHYLO_OK) return rc;
    parent_of[op.left] = op.parent;
    parent_of[op.right] = op.parent;
  }
  CK(cudaSetDevice(e->device));
  for (int o = 0; o < n_ops; ++o)
    if ((rc = fitch_ensure(e, ops[o].parent, true)) != PHYLO_OK) return rc;
  if ((rc = fitch_sync_tables(e)) != PHYLO_OK) return rc;
  if ((rc = fitch_sched_capacity(e, sizeof(FitchUpStep) * (size_t)(n_ops + 1))) != PHYLO_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  FitchUpStep *hs = (FitchUpStep *)e->hSched;
  for (int o = 0; o < n_ops; ++o) {  // reverse order: parents first
    const phylo_op &op = ops[n_ops - 1 - o];
    const bool at_root = (op.parent == root_a || op.parent == root_b);
    hs[o] = FitchUpStep{op.parent, at_root ? -1 : parent_of[op.parent], op.left, op.right};
    if (!at_root && parent_of[op.parent] < 0)
      return fail(e, PHYLO_ERR_ARG, "fitch_uppass: node %d has no parent in the schedule and is not on the root edge", op.parent);
  }
  if (n_ops > 0) {
    CK(cudaMemcpyAsync(e->dSched, hs, sizeof(FitchUpStep) * (size_t)n_ops, cudaMemcpyHostToDevice, e->stream));
    const int g = grid_for(e->fWords, 128, e->sm_count * 16);
    ProfScope prof(e, KC_FITCH_UPPASS);
    NP_DISPATCH(e->fNPdev, (fitch_uppass_kernel<NP><<<g, 128, 0, e->stream>>>(
                               e->dPreTab, e->dFinTab, (const FitchUpStep *)e->dSched, n_ops, root_a, root_b, e->fWords)));
    LAUNCH_CHECK();
  }
  CK(cudaStreamSynchronize(e->stream));
  for (int o = 0; o < n_ops; ++o) e->fFinValid[ops[o].parent] = 1;
  return PHYLO_OK;
}

extern "C" int phylo_fitch_get_states(phylo_engine *e, int node, int which, void *out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = fitch_slot_ok(e, node, true, "fitch_get_states")) != PHYLO_OK) return rc;
  if (!out) return fail(e, PHYLO_ERR_ARG, "fitch_get_states: out is NULL");
  // leaves keep their observed sets as final sets
  const bool have_fin = e->fFin[node] && e->fFinValid[node];
  const uint32_t *src = (which == 1 && have_fin) ? e->fFin[node] : e->fPre[node];
  if (which == 1 && !have_fin && node >= e->fT)
    return fail(e, PHYLO_ERR_STATE, "fitch_get_states: no final sets for node %d (run fitch_uppass)", node);
  CK(cudaSetDevice(e->device));
  const size_t row = (size_t)e->fN * e->felt;
  if ((rc = fitch_stage(e, row)) != PHYLO_OK) return rc;
  const int g = grid_for(e->fWords * 32, 256, e->sm_count * 8);
  {
  ProfScope prof(e, KC_FITCH_TRANSCODE);
  switch (e->felt) {
    case 1: fitch_decode_kernel<uint8_t><<<g, 256, 0, e->stream>>>(src, (uint8_t *)e->dStage, e->fN, e->fWords, e->fNPdev); break;
    case 2: fitch_decode_kernel<uint16_t><<<g, 256, 0, e->stream>>>(src, (uint16_t *)e->dStage, e->fN, e->fWords, e->fNPdev); break;
    case 4: fitch_decode_kernel<uint32_t><<<g, 256, 0, e->stream>>>(src, (uint32_t *)e->dStage, e->fN, e->fWords, e->fNPdev); break;
    default: fitch_decode_kernel<uint64_t><<<g, 256, 0, e->stream>>>(src, (uint64_t *)e->dStage, e->fN, e->fWords, e->fNPdev);
  }
  LAUNCH_CHECK();
  }
  CK(cudaMemcpyAsync(out, e->dStage, row, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return PHYLO_OK;
}

extern "C" int phylo_fitch_set_states(phylo_engine *e, int node, const void *codes) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = fitch_slot_ok(e, node, false, "fitch_set_states")) != PHYLO_OK) return rc;
  if (!codes) return fail(e, PHYLO_ERR_ARG, "fitch_set_states: codes is NULL");
  CK(cudaSetDevice(e->device));
  const size_t row = (size_t)e->fN * e->felt;
  if ((rc = fitch_stage(e, row)) != PHYLO_OK) return rc;
  if ((rc = fitch_ensure(e, node, false)) != PHYLO_OK) return rc;
  if ((rc = fitch_cost_capacity(e, 4)) != PHYLO_OK) return rc;
  CK(cudaMemsetAsync(e->dCost, 0, sizeof(unsigned long long), e->stream));
  CK(cudaMemcpyAsync(e->dStage, codes, row, cudaMemcpyHostToDevice, e->stream));
  if ((rc = fitch_encode(e, e->dStage, e->fPre[node], e->dCost, false)) != PHYLO_OK) return rc;  // codes, never symbols
  CK(cudaStreamSynchronize(e->stream));
  e->fValid[node] = 1;
  e->fFinValid[node] = 0;
  return PHYLO_OK;
}

// ------------------------------------------------------------ bitvector set algebra ----
static int bv_binop(phylo_engine *e, int dst, int a, int b, bool is_union) {
  int rc;
  const char *who = is_union ? "bv_union" : "bv_inter";
  if ((rc = fitch_slot_ok(e, a, true, who)) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, b, true, who)) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, dst, false, who)) != PHYLO_OK) return rc;
  CK(cudaSetDevice(e->device));
  if ((rc = fitch_ensure(e, dst, false)) != PHYLO_OK) return rc;
  const int64_t n = e->fWords * e->fNPdev;
  const int g = grid_for(n, 256, e->sm_count * 8);
  if (is_union) bv_binop_kernel<true><<<g, 256, 0, e->stream>>>(e->fPre[a], e->fPre[b], e->fPre[dst], n);
  else bv_binop_kernel<false><<<g, 256, 0, e->stream>>>(e->fPre[a], e->fPre[b], e->fPre[dst], n);
  LAUNCH_CHECK();
  e->fValid[dst] = 1;
  e->fFinValid[dst] = 0;
  return PHYLO_OK;
}

extern "C" int phylo_bv_union(phylo_engine *e, int dst, int a, int b) {
  if (!e) return PHYLO_ERR_ARG;
  return bv_binop(e, dst, a, b, true);
}
extern "C" int phylo_bv_inter(phylo_engine *e, int dst, int a, int b) {
  if (!e) return PHYLO_ERR_ARG;
  return bv_binop(e, dst, a, b, false);
}

static int bv_count(phylo_engine *e, int a, int mode, uint64_t mask, int n, uint64_t *out, const char *who) {
  int rc;
  if ((rc = fitch_slot_ok(e, a, true, who)) != PHYLO_OK) return rc;
  if (!out) return fail(e, PHYLO_ERR_ARG, "%s: out is NULL", who);
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(e->dCost, 0, sizeof(unsigned long long), e->stream));
  const int g = grid_for(e->fWords, 256, e->sm_count * 8);
  bv_count_kernel<<<g, 256, 0, e->stream>>>(e->fPre[a], e->fWords, e->fN, e->fNPdev, mode, mask, n, e->dCost);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *out = e->hCost[0];
  return PHYLO_OK;
}

extern "C" int phylo_bv_popcount(phylo_engine *e, int a, uint64_t *out) {
  if (!e) return PHYLO_ERR_ARG;
  return bv_count(e, a, 0, 0, 0, out, "bv_popcount");
}
extern "C" int phylo_bv_saturation(phylo_engine *e, int a, uint64_t state_mask, uint64_t *out) {
  if (!e) return PHYLO_ERR_ARG;
  return bv_count(e, a, 1, state_mask, 0, out, "bv_saturation");
}
extern "C" int phylo_bv_poly_saturation(phylo_engine *e, int a, int n, uint64_t *out) {
  if (!e) return PHYLO_ERR_ARG;
  if (e->fT && n > e->felt * 8) {  // lib/bitvector/bv.c:137: n > WIDTH counts nothing
    if (out) *out = 0;
    return PHYLO_OK;
  }
  return bv_count(e, a, 2, 0, n, out, "bv_poly_saturation");
}

// bv_eltcount (lib/bitvector/bv.c:59-69): number of states in the set of character i
extern "C" int phylo_bv_eltcount(phylo_engine *e, int a, int64_t i, int *out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = fitch_slot_ok(e, a, true, "bv_eltcount")) != PHYLO_OK) return rc;
  if (!out || i < 0 || i >= e->fN) return fail(e, PHYLO_ERR_ARG, "bv_eltcount: character %lld out of range", (long long)i);
  CK(cudaSetDevice(e->device));
  std::vector<uint32_t> pl(e->fNPdev);
  CK(cudaMemcpyAsync(pl.data(), e->fPre[a] + (i / 32) * e->fNPdev, sizeof(uint32_t) * e->fNPdev, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  int n = 0;
  for (int s = 0; s < e->fNPdev; ++s) n += (pl[s] >> (i % 32)) & 1u;
  *out = n;
  return PHYLO_OK;
}

extern "C" int phylo_bv_compare(phylo_engine *e, int a, int b, int *out) {
  if (!e) return PHYLO_ERR_ARG;
  int rc;
  if ((rc = fitch_slot_ok(e, a, true, "bv_compare")) != PHYLO_OK) return rc;
  if ((rc = fitch_slot_ok(e, b, true, "bv_compare")) != PHYLO_OK) return rc;
  if (!out) return fail(e, PHYLO_ERR_ARG, "bv_compare: out is NULL");
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(e->dCost, 0xff, sizeof(unsigned long long), e->stream));
  const int g = grid_for(e->fWords, 256, e->sm_count * 8);
  bv_firstdiff_kernel<<<g, 256, 0, e->stream>>>(e->fPre[a], e->fPre[b], e->fWords, e->fN, e->fNPdev, e->dCost);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(e->hCost, e->dCost, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const unsigned long long first = e->hCost[0];
  if (first == ~0ull) { *out = 0; return PHYLO_OK; }
  // fetch the two differing elements: planes of the word holding `first`
  const int64_t w = (int64_t)(first / 32);
  const int bit = (int)(first % 32);
  std::vector<uint32_t> pa(e->fNPdev), pb(e->fNPdev);
  CK(cudaMemcpy(pa.data(), e->fPre[a] + w * e->fNPdev, sizeof(uint32_t) * e->fNPdev, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pb.data(), e->fPre[b] + w * e->fNPdev, sizeof(uint32_t) * e->fNPdev, cudaMemcpyDeviceToHost));
  uint64_t ea = 0, eb = 0;
  for (int s = 0; s < e->fNPdev; ++s) {
    ea |= (uint64_t)((pa[s] >> bit) & 1u) << s;
    eb |= (uint64_t)((pb[s] >> bit) & 1u) << s;
  }
  *out = ea > eb ? 1 : -1;  // lib/bitvector/bv.c:82-84
  return PHYLO_OK;
}
