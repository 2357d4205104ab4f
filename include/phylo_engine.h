/* phylo_engine.h -- C ABI of the B200-native tree-scoring engine for phylocaml.
 *
 * This is the drop-in boundary: the entry points below are what phylocaml's OCaml
 * `external`s for the Likelihood / NonAdditive / Bitvector / MlModel hot path bind to
 * (through the thin `value` stubs in stubs/phylo_stubs.c). Plain pointers and sizes only.
 * All citations are file:line under the reference tree (amnh/phylocaml).
 *
 * Conventions
 *   - Every function returns PHYLO_OK (0) or a negative PHYLO_ERR_* code; the message is
 *     available from phylo_last_error(). Nothing aborts or exits. The OCaml stubs turn a
 *     non-zero code into `Failure msg`, the reference's own error convention
 *     (lib/mlmodel.c:109-121: failwith).
 *   - An engine handle is bound to ONE CUDA device and is not thread-safe (the reference is
 *     single-threaded: no caml_enter_blocking_section anywhere). Multi-GPU = one process
 *     (one handle) per GPU, patterns sharded contiguously (bench.py), or one
 *     phylo_group handle that owns an engine per GPU inside a single process (last section).
 *   - There is no CPU fallback: phylo_engine_create fails when no CUDA device is usable.
 *   - Matrices are row-major float64, exactly as OCaml Bigarray c_layout holds them
 *     (lib/mlModel.ml:50-51).
 *   - "Node slots" 0..capacity-1 name device-resident node data (CLV + scale counters for
 *     likelihood; state sets for Fitch). Slots 0..T-1 are the tips.
 */
#ifndef PHYLO_ENGINE_H
#define PHYLO_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHYLO_OK 0
#define PHYLO_ERR_CUDA (-1)     /* a CUDA runtime call failed */
#define PHYLO_ERR_ARG (-2)      /* bad argument (shape, slot id, NULL, ...) */
#define PHYLO_ERR_STATE (-3)    /* call order (e.g. score before set_model) */
#define PHYLO_ERR_DATA (-4)     /* invalid input data (all-zero state mask, ...) */
#define PHYLO_ERR_NUMERIC (-5)  /* eigen-decomposition failed / complex eigenvalues */
#define PHYLO_ERR_UNSUPPORTED (-6)

#define PHYLO_SCALE_EXP 256  /* CLVs are rescaled by 2^256 when a site maximum < 2^-256 */
#define PHYLO_LNL_BLOCK 1024 /* patterns per level-1 block of the site-sum reduction */

typedef struct phylo_engine phylo_engine;

/* One step of a post-order schedule: `parent` = median of `left` and `right`, reached by
 * branches t_left / t_right (ignored by Fitch). One entry per call Node.median_2 would make
 * while Tree.post_order_edges walks the tree (lib/tree.ml:171-187, lib/node.ml:183-198).
 * The reference stores no branch lengths (lib/tree.ml:19-22); they travel here. */
typedef struct {
  int32_t parent, left, right, pad_;
  double t_left, t_right;
} phylo_op;

/* ----------------------------------------------------------------------- engine ---- */
int phylo_engine_create(int device, phylo_engine **out);
void phylo_engine_destroy(phylo_engine *e);
/* message of the last failed call on `e`; e == NULL: last phylo_engine_create failure */
const char *phylo_last_error(const phylo_engine *e);
/* all work is issued on this CUDA stream (a cudaStream_t; NULL = legacy default stream) */
int phylo_engine_set_stream(phylo_engine *e, void *cuda_stream);
int phylo_engine_sync(phylo_engine *e);
/* kernels launched by this engine since creation (bench.py's gpu_launches evidence) */
uint64_t phylo_engine_launch_count(const phylo_engine *e);
/* Alphabet symbols instead of state masks (the data format one step before the path). With a
 * table set, every alignment handed to phylo_lk_set_tips(_pitched), phylo_lk_score_alignment,
 * phylo_fitch_set_tips(_pitched) and phylo_compress_patterns must be 1 byte per cell and is read
 * as symbols: cell c stands for the state set table256[c], translated on the device by the
 * kernels that already convert the upload (no extra pass, no host loop). The table is what
 * Alphabet's name -> code map holds for single-character names, e.g. Alphabet.nucleotides
 * (lib/alphabet.ml:309-326: A,C,G,T,- = 1,2,4,8,16, IUPAC and indel-polymorphism letters = the OR
 * of their states, ? = 31) or Alphabet.dna (:301-307); `case:false` alphabets list both cases.
 * A symbol whose entry is 0 is unknown: the call fails with PHYLO_ERR_DATA, like the reference's
 * `Illegal_Character` (lib/alphabet.ml:186). For likelihood the caller folds gap / missing into
 * "all states" in the table (lib/mlModel.mli:75-76). Fitch needs every entry < 256 (sets come
 * back one byte per character); phylo_compress_patterns likewise, and returns state masks.
 * table256 == NULL restores plain state-mask input. phylo_fitch_set_states always takes codes. */
int phylo_engine_set_symbol_table(phylo_engine *e, const uint64_t *table256);
/* Engine options. PHYLO_OPT_FUSED_TREE (default 1): phylo_lk_score_tree evaluates eligible
 * schedules (4 states, K in {1,2,4,8}, plain tree) with a single-launch tree-fused kernel
 * (warp-autonomous kernel for K <= 4, tile kernel otherwise); 2 = tile kernel only;
 * 0 = one kernel per node. PHYLO_OPT_RETAIN_CLV (default 1): every interior CLV of a
 * score_tree call is left in its node slot (for phylo_lk_get_clv / edge_lnl / incremental
 * re-scoring); 0 = lnL only, the tree-fused kernel then writes no CLV at all. The same option
 * governs phylo_fitch_score_tree (4 planes, tile kernel): 0 = the length only -- no interior set
 * is written (the parents' slots are left invalid, their node costs 0) and the tree is evaluated
 * from its centre edge (phylo_fitch_reroot). */
#define PHYLO_OPT_FUSED_TREE 1
#define PHYLO_OPT_RETAIN_CLV 2
/* PHYLO_OPT_FITCH_WALK selects the whole-tree Fitch kernel: 1 (default) = automatic (4 planes:
 * the on-chip tile kernel at every size; up to 8 planes below ~8 M characters: the register walk;
 * otherwise the L2 walk); 3 = on-chip tile kernel (4 planes: subtrees dealt to the warps of a CTA,
 * medians out of shared memory with the running set in registers, results published by the last
 * CTA); 2 = register walk (compiled depth-first plan, tips prefetched several medians ahead, up to
 * 8 planes); 0 = L2 walk (re-reads its own earlier writes through L2, any plane count). */
#define PHYLO_OPT_FITCH_WALK 3
/* PHYLO_OPT_DEFER_SCALAR (default 0): 1 = phylo_lk_score_tree leaves the evaluation's level-1 block partials
 * on the device and returns without waiting for (or computing) the local lnL (*lnl_out = NaN); the caller
 * combines them across ranks with phylo_lk_exchange_reduce. No host round trip between the two. */
#define PHYLO_OPT_DEFER_SCALAR 4
/* Environment variables exist for measurements and cross-checks only, never as configuration
 * (tools/tune_treew.sh, tools/mma_ab.sh, tools/grad_mma_ab.sh, tools/uppass_ab.sh):
 * PHYLO_TREEW_TUNE="slev,R,il" overrides the geometry the warp-autonomous tree kernel picks, and
 * PHYLO_TT_TABLE=0 sends 20-state tip+tip updates through the DMMA kernel instead of the table
 * copy. PHYLO_TREEW_TUNE never changes a bit of the result; the two tip+tip paths give identical
 * CLVs for observed tips and agree to rounding (1e-16 relative) where a tip is ambiguous.
 * PHYLO_TREEM_PAIRED / PHYLO_TREEM_STSWAP / PHYLO_TREEM_TIMING / PHYLO_FITCH_TIMING: A/B switches and
 * per-CTA timelines of the tree-fused DMMA and Fitch tile kernels (bit-identical results).
 * PHYLO_UPPASS_BATCH=0: phylo_lk_uppass launches one kernel per update instead of one per tree
 * level (bit-identical). PHYLO_GRAD_MMA=0: phylo_lk_param_gradient runs the scalar per-branch kernel
 * for 4 states instead of the tensor-core kernel (agreement 1e-14 relative); PHYLO_GRAD_CHUNK_BYTES
 * caps the scratch of that kernel (tests drive the several-launches path with it; bit-identical). */
int phylo_engine_set_option(phylo_engine *e, int option, int64_t value);
int phylo_engine_get_option(phylo_engine *e, int option, int64_t *value);
/* CUDA-event profiler: while enabled, every kernel launch is bracketed by an event pair on
 * the engine's stream and its duration accumulated per kernel class (0 <= class <
 * phylo_kernel_class_count()). bench.py's roofline figures come from here. */
int phylo_engine_profile(phylo_engine *e, int enable);
int phylo_engine_profile_reset(phylo_engine *e);
int phylo_engine_profile_get(phylo_engine *e, int kernel_class, double *ms_total, uint64_t *launches);
int phylo_kernel_class_count(void);
const char *phylo_kernel_class_name(int kernel_class);
/* page-locked host memory for Bigarray-backed staging buffers (full-rate H2D/D2H) */
int phylo_host_alloc(void **out, uint64_t bytes);
int phylo_host_free(void *p);

/* ---------------------------------------------- MlModel native half (lib/mlmodel.c) ---- */
/* Replaces diagonalize_sym (lib/mlmodel.c:163-197; OCaml external lib/mlModel.ml:78-79).
 * In: Q (n*n, symmetric). Out: Q overwritten with U whose ROWS are eigenvectors, D = n*n
 * matrix with eigenvalues on the diagonal (mlmodel.c:81-88). Host-side, once per model. */
int phylo_diagonalize_sym(double *Q_inout_U, double *D, int n);
/* Replaces diagonalize_gtr (lib/mlmodel.c:208-262; lib/mlModel.ml:73-74). In: Q (n*n).
 * Out: Q overwritten with U, D diagonal matrix, Ui = U^-1, such that Q = U D Ui row-major.
 * Fails with PHYLO_ERR_NUMERIC on complex eigenvalues (mlmodel.c:248-250). */
int phylo_diagonalize_gtr(double *Q_inout_U, double *D, double *Ui, int n);
/* Discrete-Gamma rate classes (lib/mlModel.ml:93-99, :676-694; the reference computes them with
 * Pareto/GSL, which is neither vendored nor pinned). mode 0 = what the reference's code does:
 * rates[i] = quantile at p = i/k of Gamma(shape = alpha, scale = alpha), so rates[0] = 0;
 * mode 1 = what lib/mlModel.mli:12 documents: Yang's (1994) class means of Gamma(alpha, rate alpha),
 * average 1. probs (may be NULL) = 1/k each. Host-side, self-contained incomplete-gamma code. */
int phylo_gamma_rates(double alpha, int k, int mode, double *rates, double *probs);
/* Replace compose_sym / compose_gtr (lib/mlmodel.c:280-302, :325-342; lib/mlModel.ml:83-90):
 * P = exp(Q t) from the eigensystem, computed by the pt_build kernel. Same special cases:
 * t == -1.0 -> Q; t < 1e-10 -> I; compose_sym rounds t to float first (mlmodel.c:280).
 * D is the full n*n matrix. P_out: n*n host buffer. */
int phylo_compose_sym(phylo_engine *e, const double *U, const double *D, double t, int n,
                      double *P_out);
int phylo_compose_gtr(phylo_engine *e, const double *U, const double *D, const double *Ui,
                      double t, int n, double *P_out);

/* ------------------------------------ Likelihood node data (lib/likelihood_c.ml) ---- */
/* The model record MlModel.t (lib/mlModel.ml:53-63): S states, K rate classes; u, d (full
 * S*S), ui (NULL <=> `ui = None`, symmetric); priors[S]; rates[K], probs[K];
 * pinvar < 0 <=> `pinvar = None`. Rate-scaled lengths t*rates[k] are formed on device. */
int phylo_lk_set_model(phylo_engine *e, int S, int K, const double *U, const double *D,
                       const double *Ui, const double *priors, const double *rates,
                       const double *probs, double pinvar);
/* Tip data: T taxa x N site patterns of state masks (bit i <=> state i possible,
 * lib/alphabet.ml:193-196), mask_bytes in {1,2,4,8} like the bitvector widths
 * (lib/bitvector/bv.h:29-55), tip-major. weights: N pattern weights or NULL (all 1).
 * capacity >= T: number of node slots. A mask with none of the low S bits set is rejected.
 * mask_bytes == 0 (4-state models only, no symbol table): PACKED input, two 4-bit masks per byte
 * (pattern 2j in the low nibble of byte j), rows of ceil(N/2) bytes -- half the PCIe bytes of the
 * one-byte form and no conversion pass on the device (the nibbles go straight into the tiles the
 * tree-fused kernel reads). phylo_pack_nibbles is the host-side packer. */
int phylo_lk_set_tips(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                      const double *weights, int capacity);
/* masks (T x N, one byte per cell, tip-major) -> packed (T rows of ceil(N/2) bytes). Host side. */
int phylo_pack_nibbles(const uint8_t *masks, int T, int64_t N, uint8_t *packed);
/* Same, for a column slab of a wider host matrix: consecutive taxon rows are host_pitch_bytes
 * apart (0 = N * mask_bytes, or ceil(N/2) for packed input; a packed slab must start at an even pattern). This is how phylo_group hands each device its shard without
 * repacking the alignment on the host. */
int phylo_lk_set_tips_pitched(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                              uint64_t host_pitch_bytes, const double *weights, int capacity);
/* Node-slot lifetime. NodeData.S is functional: every median_2 returns a NEW node value and the old
 * ones die with the OCaml GC. The reference ties native node data to the GC with a custom block
 * whose finalizer frees it (lib/bitvector/bv.c:183-189 bv_CAML_free, :229-244 the ops table); here
 * the custom block (stubs/phylo_stubs.c, `node`) holds a slot taken with phylo_lk_node_alloc and its
 * finalizer calls phylo_lk_node_release. A released slot keeps its device buffer and is handed out
 * again first, so a tree search allocates no device memory in steady state; when every slot is
 * live the slot table grows (total slots double; phylo_lk_node_stats reports it). `generation`
 * identifies the loaded alignment: releasing a slot of an alignment that has since been replaced
 * (another shape / alphabet: every slot was dropped) is a no-op, so late finalizers are harmless.
 * Callers that name slots themselves in a schedule (bench, tests) need none of this. */
int phylo_lk_node_alloc(phylo_engine *e, int *slot_out, uint64_t *generation_out);
int phylo_lk_node_release(phylo_engine *e, int slot, uint64_t generation);
/* any pointer may be NULL. capacity: interior slots in the table (total slots - T); in_use: slots
 * currently handed out; with_buffers: interior slots that own a device CLV buffer (live or pooled) */
int phylo_lk_node_stats(phylo_engine *e, int *capacity, int *in_use, int *with_buffers);
/* the loaded likelihood alignment: taxa, site patterns, and the current size of the slot table (tips
 * included; what an array indexed by slot, like phylo_lk_uppass's up_slot, must cover). NULL = not wanted */
int phylo_lk_shape(phylo_engine *e, int *n_taxa, int64_t *n_patterns, int *n_slots);
/* Likelihood.median_2 (lib/nodeData.ml:21, lib/likelihood_c.ml:15): CLV of `parent` from its
 * two children with per-site rescaling. */
int phylo_lk_median_2(phylo_engine *e, int parent, int left, double t_left, int right,
                      double t_right);
/* Likelihood.median_3 (lib/nodeData.ml:22; TODO in lib/likelihood_c.ml:16): the CLV of a node given all
 * three neighbours, (P_a L_a) o (P_b L_b) o (P_c L_c), rescaled like median_2. Uses the edge sum
 * table's buffer as scratch (a prepared edge must be re-prepared). */
int phylo_lk_median_3(phylo_engine *e, int parent, int a, double t_a, int b, double t_b, int c, double t_c);
/* Whole-tree entry point: run the schedule, join across the root edge (a,b) of length
 * root_t, return lnL. Every interior CLV stays resident in its slot for later
 * phylo_lk_edge_lnl / phylo_lk_get_clv / incremental re-scoring. */
int phylo_lk_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                        double root_t, double *lnl_out);
/* 3-directional CLVs -- what Node.Make3D keeps per node (lib/node.ml:363-477: one value per excluded
 * neighbour, `dir`) and readjust_3 consumes (lib/node.ml:239-256). After phylo_lk_score_tree over the same
 * schedule (CLVs retained), a pre-order pass fills, for every node v below the root edge with
 * up_slot[v] >= 0, the slot up_slot[v] with up[v]: the CLV of the rest of the tree at the far end of the
 * branch above v,  up[v] = (P(t_sibling) down[sibling]) o (P(t_parent) up[parent]),  up[a] = down[b] and
 * up[b] = down[a] across the root edge (their up_slot entries are ignored). up_slot: `capacity` entries
 * (-1 = not wanted; a wanted node needs its parent's up value, so its ancestors' entries must be set
 * too); the slots must be interior, distinct, and not used by the schedule. One pruning update per
 * filled slot (2 T - 4 for a whole tree). Afterwards ANY edge is a root edge:
 * phylo_lk_edge_lnl / _edge_prepare / _optimize_branch (v, up_slot[v]) work on the branch above v, and for
 * a reversible model every edge gives the same lnL (pulley principle) -- an SPR / TBR candidate costs one
 * edge join instead of a re-prune (lib/tree.ml:299-494). */
int phylo_lk_uppass(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b, double root_t,
                    const int32_t *up_slot);
/* Model-parameter derivatives -- what gen_subst_opt_func / gen_rates_opt_func / gen_prior_opt_func
 * (lib/mlModel.ml:822-829, all `failwith "todo"`) need from the native side. After phylo_lk_score_tree and
 * phylo_lk_uppass with EVERY up slot set (so that each branch has its directional pair), returns
 * grad_out[p] = d lnL / d theta_p for n_params (<= 64) parameters at once. The caller describes a parameter
 * by its effect on the model record (any pointer may be NULL = no effect):
 *   dQ     [n_params][S][S]  d Q / d theta_p of the rate matrix the eigensystem was taken from
 *                            (normalisation included: MlModel.m_meanrate, lib/mlModel.ml:204-213),
 *   drates [n_params][K]     d rates[k] / d theta_p (the Gamma shape alpha moves only these),
 *   dpi    [n_params][S]     d priors / d theta_p at the root (not together with an invariant-sites class).
 * lnL is multilinear in the branches' P(t): d P = dexp_{Q t r}[t (r dQ + dr Q)] is formed per branch on the
 * host from the eigensystem and applied across the branch's directional pair (all parameters together;
 * 2 T - 3 pairs of 2 C bytes per pattern -- against 2 n_params full evaluations for central differences).
 * 4 states: ONE launch over all (branch, 1024-pattern block) items on the fp64 tensor cores, seven
 * parameters per pass (param_grad4_mma_kernel); other alphabets: one scalar kernel pass per branch.
 * Reversible models (the pulley principle behind phylo_lk_uppass). *lnl_out (may be NULL) = the lnL of the
 * preceding phylo_lk_score_tree. */
int phylo_lk_param_gradient(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b, double root_t,
                            const int32_t *up_slot, int n_params, const double *dQ, const double *drates,
                            const double *dpi, double *lnl_out, double *grad_out);
/* Host-only (no GPU needed), for tests and tooling: the compiled form of a schedule as the
 * tree-fused likelihood kernels and the Fitch register walk execute it. One row of 6 int32 per
 * step (n_ops medians in depth-first order + the root-edge join, whose out_slot is -1):
 * left kind, left slot, right kind, right slot, push_first, out_slot. Kinds: 0 = tip, 1 = the value
 * the previous step produced (in registers), 2 = popped from the on-chip stack, 3 = a CLV / set
 * already resident in its node slot from an earlier call. push_first = 1: the live value is
 * pushed before this step. *depth_out = stack levels needed (the subtree with the larger need is
 * evaluated first). PHYLO_ERR_UNSUPPORTED: not a plain tree -- the per-node path runs such
 * schedules. */
int phylo_plan_compile(const phylo_op *ops, int n_ops, int T, int capacity, int root_a, int root_b,
                       int32_t *steps_out, int *depth_out);
/* Host-only (no GPU needed), for tests and tooling: the schedule phylo_fitch_score_tree evaluates for a
 * length-only call (PHYLO_OPT_RETAIN_CLV = 0) -- the same unrooted tree, re-rooted on the edge that minimises
 * the height of its two halves (Fitch length does not depend on the root; the chain of dependent medians
 * does). ops_out has room for n_ops ops; a schedule that is not one tree comes back unchanged. */
int phylo_fitch_reroot(const phylo_op *ops, int n_ops, int capacity, int root_a, int root_b, phylo_op *ops_out,
                       int *root_a_out, int *root_b_out);
/* phylo_lk_set_tips + phylo_lk_score_tree in one call for an alignment that is still in host
 * memory: the upload is cut into pattern slabs on a second stream and each slab is scored
 * (tree-fused kernel) while the next one is still crossing PCIe. Same result, bit for bit. */
int phylo_lk_score_alignment(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                             const double *weights, int capacity, const phylo_op *ops, int n_ops,
                             int root_a, int root_b, double root_t, double *lnl_out);
/* Likelihood.root_cost / distance_1 (lib/nodeData.ml:29,32): lnL of joining the directed
 * CLVs a and b across an edge, for n_t candidate lengths (branch-length loop). */
int phylo_lk_edge_lnl(phylo_engine *e, int a, int b, const double *t, int n_t, double *lnl_out);
/* The same join for n_edges different edges (a_slots[i], b_slots[i]) of lengths t[i] in one call: one P(t)
 * build, one synchronisation. With phylo_lk_uppass's directional CLVs this is the scoring loop of a
 * neighbourhood search: one entry per candidate (lib/tree.ml:299-494). Values equal phylo_lk_edge_lnl's. */
int phylo_lk_edge_lnl_batch(phylo_engine *e, int n_edges, const int32_t *a_slots, const int32_t *b_slots,
                            const double *t, double *lnl_out);
/* Branch-length loop on the edge (a, b) -- Likelihood.adjust_3 / readjust (lib/nodeData.ml:25,
 * lib/node.ml:239-256; TODO in lib/likelihood_c.ml:19-24). phylo_lk_edge_prepare builds the
 * edge's sum table (one CLV-sized array, eigen-space products of the two CLVs) once;
 * phylo_lk_edge_eval then returns lnL(t) and its first and second derivatives for n_t lengths,
 * each pass streaming only that table (d1_out / d2_out may be NULL). Values agree with
 * phylo_lk_edge_lnl to rounding (<= 1e-12 relative). The prepared edge stays valid until the
 * model or the tips change; re-prepare after the CLVs of a or b change. */
int phylo_lk_edge_prepare(phylo_engine *e, int a, int b);
int phylo_lk_edge_eval(phylo_engine *e, const double *t, int n_t, double *lnl_out, double *d1_out,
                       double *d2_out);
/* Maximises lnL over the length of edge (a, b) in [t_min, t_max] by safeguarded Newton steps
 * from t0 (tol: relative step tolerance, <= 0 -> 1e-8; max_iter < 1 -> 50). */
int phylo_lk_optimize_branch(phylo_engine *e, int a, int b, double t0, double t_min, double t_max,
                             double tol, int max_iter, double *t_opt, double *lnl_opt, int *iters_out);
/* Read-back. clv_out: N*K*S doubles [pattern][k][i]; scale_out: N int32 or NULL. */
int phylo_lk_get_clv(phylo_engine *e, int node, double *clv_out, int32_t *scale_out);
/* per-pattern ln-likelihoods (unweighted) of the last score_tree / edge_lnl (last t) */
int phylo_lk_get_site_lnl(phylo_engine *e, double *out);
/* level-1 block partials of the last evaluation (ceil(N/1024) doubles) and the canonical
 * reduction of such partials -- what ranks all-gather for a bit-reproducible multi-GPU sum */
int phylo_lk_get_block_partials(phylo_engine *e, double *out, int64_t *n_out);
double phylo_reduce_partials(const double *partials, int64_t n);

/* --------------------------- the scalar exchange, on the device over peer-mapped memory ---- */
/* patterns are sharded over ranks (one engine per GPU) and the ONLY exchange is the final scalar.
 * Instead of host -> NCCL all-reduce -> host, every engine owns a small mailbox in its HBM that all other
 * engines map (NVLink / NVSwitch peer access): phylo_exchange_alloc creates it and exports a 64-byte
 * cudaIpcMemHandle for other PROCESSES (one process per GPU: ship the handles with any host-side
 * all-gather, open each with phylo_exchange_open); engines of one process pass the raw pointers.
 * phylo_exchange_set(world, rank, mailboxes[world]) installs the table (mailboxes[rank] = the own one).
 * phylo_lk_exchange_reduce: call on EVERY rank after phylo_lk_score_tree (best with PHYLO_OPT_DEFER_SCALAR):
 * one CTA writes the rank's level-1 block partials into every peer's mailbox, waits for all ranks' partials in
 * its own, and folds their concatenation in rank order with the canonical 1024-fold -- *lnl_out is the
 * whole alignment's lnL, bit-identical to a single-engine evaluation for any rank count (shards must start
 * at multiples of PHYLO_LNL_BLOCK patterns; at most 8190 blocks per rank). phylo_exchange_sum_u64: the exact
 * integer sum of one value per rank (Fitch / TCM lengths). A rank that does not arrive within ~2 s makes the
 * call fail with PHYLO_ERR_CUDA instead of hanging. bench.py uses these (NCCL all-reduce = --exchange nccl). */
int phylo_exchange_alloc(phylo_engine *e, void **mailbox_out, unsigned char *ipc_handle64);
int phylo_exchange_open(phylo_engine *e, const unsigned char *ipc_handle64, void **mailbox_out);
int phylo_exchange_set(phylo_engine *e, int world, int rank, void *const *mailboxes);
int phylo_lk_exchange_reduce(phylo_engine *e, double *lnl_out);
int phylo_exchange_sum_u64(phylo_engine *e, uint64_t value, uint64_t *sum_out);

/* ------------------------------------------------ site-pattern compression (next to the path) ---- */
/* The step before scoring: identical alignment columns are merged into one site pattern whose
 * weight is the sum of the columns' weights (the `weights` of NonAdditive_c.t,
 * lib/nonAdditive_c.ml:3, and of phylo_lk_set_tips). masks: T x N, tip-major, elements of
 * mask_bytes; weights_in: N doubles or NULL (1 each). Outputs are HOST buffers sized for the
 * worst case: patterns_out T x N elements (written compactly as T rows of *n_patterns),
 * weights_out N doubles, site_to_pattern N int32 (may be NULL). Patterns come out in order of
 * first occurrence, so the result is deterministic. Done on the device (hashing into an
 * open-addressing table, full-column verification, prefix sum); no CPU fallback. */
int phylo_compress_patterns(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                            const double *weights_in, void *patterns_out, double *weights_out,
                            int32_t *site_to_pattern, int64_t *n_patterns);
/* The same for a column slab of a wider host matrix: row t starts at masks + t * host_pitch_bytes
 * (0 = dense rows). */
int phylo_compress_patterns_pitched(phylo_engine *e, int T, int64_t N, const void *masks, int mask_bytes,
                                    uint64_t host_pitch_bytes, const double *weights_in, void *patterns_out,
                                    double *weights_out, int32_t *site_to_pattern, int64_t *n_patterns);

/* -------------------------- NonAdditive node data / Bitvector (lib/nonAdditive_c.ml) ---- */
/* T taxa x N characters, one character per element of elt_bytes (W/8, W in {8,16,32,64},
 * lib/bitvector/bv.h:29-55), n_states = number of usable low bits (vect.msize, bv.h:61).
 * weights: N non-negative integer-valued doubles (lib/nonAdditive_c.ml:3) or NULL (all 1,
 * the bv_fitch case). On device the characters live bit-sliced (one 32-bit word per state
 * plane per 32 characters). An all-zero element is rejected. */
/* elt_bytes == 0: `codes` already is that device layout -- per taxon ceil(N/32) words x NP planes of
 * uint32 (NP = phylo_fitch_plane_count(n_states); DNA: 16 bytes per 32 characters = 0.5 B/char), rows
 * ceil(N/32)*NP*4 bytes apart, as phylo_fitch_pack_planes writes it. The upload is then a plain copy
 * into the node buffers plus one checking pass (no transcoding), and phylo_fitch_get_states returns
 * sets in the narrowest element that holds n_states bits. */
int phylo_fitch_set_tips(phylo_engine *e, int T, int64_t N, int elt_bytes, int n_states,
                         const void *codes, const double *weights, int capacity);
/* host-side packer: reference layout (one character per element of elt_bytes) -> planes; and the plane
 * count the device uses for n_states (one of 1,2,3,4,5,6,8,12,16,24,32,64) */
int phylo_fitch_pack_planes(const void *codes, int elt_bytes, int n_states, int T, int64_t N, uint32_t *planes);
int phylo_fitch_plane_count(int n_states);
/* column slab of a wider host matrix, rows host_pitch_bytes apart (0 = N * elt_bytes; plane input: a
 * slab starts at a multiple of 32 characters and 0 = ceil(N/32) * NP * 4) */
int phylo_fitch_set_tips_pitched(phylo_engine *e, int T, int64_t N, int elt_bytes, int n_states,
                                 const void *codes, uint64_t host_pitch_bytes, const double *weights,
                                 int capacity);
/* node-slot lifetime for the state sets: same contract as phylo_lk_node_alloc / _release / _stats
 * (what bv_CAML_free, lib/bitvector/bv.c:183-189, does for a `vect`) */
int phylo_fitch_node_alloc(phylo_engine *e, int *slot_out, uint64_t *generation_out);
int phylo_fitch_node_release(phylo_engine *e, int slot, uint64_t generation);
int phylo_fitch_node_stats(phylo_engine *e, int *capacity, int *in_use, int *with_buffers);
/* NonAdditive.median_2 (lib/nonAdditive_c.ml:19-35) == bv_fitch (lib/bitvector/bv.c:148-160;
 * stub bv_CAML_fitch_median2 :463-480): parent set + cost of this node alone. */
int phylo_fitch_median_2(phylo_engine *e, int parent, int left, int right, uint64_t *cost_out);
/* NonAdditive.median_3 (lib/nodeData.ml:22; only sketched in the reference: the commented-out
 * bv_CAML_fitch_median3(vb0, vb1, vb2, vb3) of lib/bitvector/bv.h:94): final state set of a node from
 * its own preliminary set `prelim`, its parent's final set and its children's preliminary sets
 * (Fitch's second pass, the rule phylo_fitch_uppass applies to the whole tree). Written to `dst`
 * as an ordinary node value. */
int phylo_fitch_median_3(phylo_engine *e, int dst, int prelim, int parent_final, int left, int right);
/* bv_distance (lib/bitvector/bv.c:46-55; stub :455-461) */
int phylo_fitch_distance(phylo_engine *e, int a, int b, uint64_t *dist_out);
/* Whole-tree down-pass: sum of interior median costs + root-edge distance. */
int phylo_fitch_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a,
                           int root_b, uint64_t *length_out);
/* per-node costs of the last score_tree (one uint64 per slot: `capacity` of them, or what
 * phylo_fitch_node_stats reports (+ T) once phylo_fitch_node_alloc has grown the table; Node.cost is
 * node-local, lib/node.ml:191) */
int phylo_fitch_get_node_costs(phylo_engine *e, uint64_t *out);
/* Up-pass / final state sets (Node.final_states, lib/node.ml:260-268 -- TODO in the
 * reference; rule in DESIGN.md). Requires a preceding down-pass over the same schedule. */
int phylo_fitch_uppass(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b);
/* which: 0 = preliminary (down-pass) sets, 1 = final sets. out: N elements of elt_bytes. */
int phylo_fitch_get_states(phylo_engine *e, int node, int which, void *out);
/* upload one node's preliminary sets (Bitvector.of_array, lib/bitvector/bv.c:305-327) */
int phylo_fitch_set_states(phylo_engine *e, int node, const void *codes);

/* General-TCM parsimony median on state sets -- CostMatrix.find_median_general / _metric
 * (lib/costMatrix.ml:68-86, :107-124): for two sets a, b the cost is the minimum over i in a,
 * j in b and a candidate median state k of M[i][k] + M[j][k], the median is the set of all k
 * that reach it (candidates: every state; metric != 0: only the states of a | b). M is
 * n_states x n_states, row-major, 0 <= M <= 30000, n_states <= 6 and equal to the n_states of
 * phylo_fitch_set_tips; the medians live in the same node slots as the Fitch sets. With the
 * 0/1 matrix both variants are the Fitch rule (the reference tests exactly that,
 * test/costMatrixTest.ml:110-125). parent < 0 in phylo_tcm_median_2: cost only. */
int phylo_tcm_set_matrix(phylo_engine *e, int n_states, const int32_t *M, int metric);
/* MlModel.integerized_model (lib/mlModel.ml:639-660), the bridge from a likelihood model to such
 * a cost matrix: cost[i][j] = -trunc(10^sigma * ln(priors[i] * P[i][j])) for P = P(t) from
 * phylo_compose_* (priors == NULL for JC69 / K2P, which the reference leaves out there). Host
 * side. PHYLO_ERR_NUMERIC when an entry of P is not positive. Mind phylo_tcm_set_matrix's range
 * (costs <= 30000): sigma = 3 is the largest that fits for usual branch lengths (the reference's
 * default sigma = 4 gives costs up to ~10^5). */
int phylo_integerize_matrix(const double *P, const double *priors, int n, int sigma, int32_t *cost_out);
int phylo_tcm_median_2(phylo_engine *e, int parent, int left, int right, uint64_t *cost_out);
int phylo_tcm_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                         uint64_t *length_out);

/* Cost-vector (Sankoff) parsimony under a general transformation cost matrix -- the weighted-state form of the
 * NonAdditive path. The reference names the passes but never wrote them (commented-out
 * bv_CAML_sankoff_median2_downpass / _uppass, lib/bitvector/bv.h:97-98); its set-based medians are
 * phylo_tcm_* above. Per character and node a vector c[s] = cheapest cost of the subtree given state s;
 *   median: c_p[s] = min_i (M[s][i] + c_l[i]) + min_j (M[s][j] + c_r[j])   (M[s][i]: parent state s -> child state i),
 *   tips: 0 for the states of the mask, "infinity" elsewhere; length = sum_chars w min_{i,j} c_a[i] + M[i][j] + c_b[j]
 * across the root edge. Up to 32 states (amino acids fit), any non-negative integer costs <= 10^6, no metricity
 * assumed. codes: T x N state masks, one per element of elt_bytes in {1,2,4}; weights: non-negative integers as
 * doubles or NULL. phylo_sankoff_score_tree runs a plain tree in ONE launch (running vector in registers, parked
 * vectors on a per-thread stack: only the tip masks are read) and leaves every node's vectors in its slot when
 * PHYLO_OPT_RETAIN_CLV is set; other schedules go node by node. With the 0/1 matrix the length is the Fitch
 * length. Loading characters (or a matrix) of another alphabet size drops the matrix (the characters).
 * phylo_sankoff_median_2 also returns sum_chars w min_s c_p[s], the cost of the subtree below `parent`. */
int phylo_sankoff_set_matrix(phylo_engine *e, int n_states, const int32_t *M);
int phylo_sankoff_set_tips(phylo_engine *e, int T, int64_t N, int elt_bytes, int n_states, const void *codes,
                           const double *weights, int capacity);
int phylo_sankoff_median_2(phylo_engine *e, int parent, int left, int right, uint64_t *subtree_cost_out);
int phylo_sankoff_score_tree(phylo_engine *e, const phylo_op *ops, int n_ops, int root_a, int root_b,
                             uint64_t *length_out);
/* out: N x n_states int32, values >= 2^28 mean "impossible" */
int phylo_sankoff_get_costs(phylo_engine *e, int node, int32_t *out);

/* Bitvector set algebra over node slots (lib/bitvector/bv.c:59-144; stubs :405-453) */
int phylo_bv_union(phylo_engine *e, int dst, int a, int b);
int phylo_bv_inter(phylo_engine *e, int dst, int a, int b);
int phylo_bv_popcount(phylo_engine *e, int a, uint64_t *out);
int phylo_bv_saturation(phylo_engine *e, int a, uint64_t state_mask, uint64_t *out);
int phylo_bv_poly_saturation(phylo_engine *e, int a, int n, uint64_t *out);
int phylo_bv_compare(phylo_engine *e, int a, int b, int *out);
/* bv_eltcount (lib/bitvector/bv.c:59-69): states in the set of character i */
int phylo_bv_eltcount(phylo_engine *e, int a, int64_t i, int *out);

/* --------------------------------------- several GPUs behind one handle (one process) ---- */
/* The reference is a single OCaml process (no threads: no caml_enter_blocking_section anywhere under lib/),
 * so a drop-in that wants all GPUs of a box cannot rely on one-process-per-GPU launchers. A
 * phylo_group owns one engine per listed device (a device may be listed more than once) and one
 * host worker thread per engine; every call below fans out to the workers and returns when all
 * are done. Site patterns / characters are cut into contiguous shards whose boundaries are
 * multiples of PHYLO_LNL_BLOCK; the model, the schedule and P(t) are replicated.
 * The path's only exchange is the final scalar: the level-1 block partials of every shard are
 * concatenated in shard order and folded by phylo_reduce_partials, so lnL is BIT-IDENTICAL to the
 * single-engine value for any device count; Fitch / TCM lengths are exact integer sums. No
 * device-to-device traffic is needed (the scalars come back through each engine's mapped host
 * block), hence no NCCL dependency in the library; bench.py's one-process-per-GPU arm does the same
 * sum with an NCCL allreduce. Errors: the first failing shard's message, phylo_group_last_error. */
typedef struct phylo_group phylo_group;
int phylo_group_create(const int *devices, int n_devices, phylo_group **out);
void phylo_group_destroy(phylo_group *g);
const char *phylo_group_last_error(const phylo_group *g);
int phylo_group_size(const phylo_group *g);
/* engine of shard i (options, profiler, per-shard read-back); NULL when out of range */
phylo_engine *phylo_group_engine(phylo_group *g, int i);
/* patterns [*lo, *hi) of the loaded likelihood alignment (which = 0) or Fitch characters
 * (which = 1) that shard i holds; lo == hi for a shard left empty by a short alignment */
int phylo_group_shard(const phylo_group *g, int which, int i, int64_t *lo, int64_t *hi);
int phylo_group_set_option(phylo_group *g, int option, int64_t value);
int phylo_group_set_symbol_table(phylo_group *g, const uint64_t *table256);
int phylo_group_lk_set_model(phylo_group *g, int S, int K, const double *U, const double *D,
                             const double *Ui, const double *priors, const double *rates,
                             const double *probs, double pinvar);
int phylo_group_lk_set_tips(phylo_group *g, int T, int64_t N, const void *masks, int mask_bytes,
                            const double *weights, int capacity);
int phylo_group_lk_score_tree(phylo_group *g, const phylo_op *ops, int n_ops, int root_a, int root_b,
                              double root_t, double *lnl_out);
/* sums over shards in shard order (deterministic; not the blocked reduction) */
int phylo_group_lk_edge_lnl(phylo_group *g, int a, int b, const double *t, int n_t, double *lnl_out);
/* safeguarded Newton of phylo_lk_optimize_branch on the summed lnL, d1, d2 of all shards */
int phylo_group_lk_optimize_branch(phylo_group *g, int a, int b, double t0, double t_min, double t_max,
                                   double tol, int max_iter, double *t_opt, double *lnl_opt,
                                   int *iters_out);
int phylo_group_lk_get_site_lnl(phylo_group *g, double *out);
int phylo_group_lk_get_clv(phylo_group *g, int node, double *clv_out, int32_t *scale_out);
int phylo_group_fitch_set_tips(phylo_group *g, int T, int64_t N, int elt_bytes, int n_states,
                               const void *codes, const double *weights, int capacity);
int phylo_group_fitch_score_tree(phylo_group *g, const phylo_op *ops, int n_ops, int root_a,
                                 int root_b, uint64_t *length_out);
int phylo_group_fitch_uppass(phylo_group *g, const phylo_op *ops, int n_ops, int root_a, int root_b);
int phylo_group_fitch_get_states(phylo_group *g, int node, int which, void *out);
/* phylo_compress_patterns over all devices of the group: every device compresses a contiguous slab of sites,
 * device 0 merges the slabs' pattern tables (a second pass with their weights as input). Patterns, weights and the
 * site map are identical to those of one device for the whole alignment. */
int phylo_group_compress_patterns(phylo_group *g, int T, int64_t N, const void *masks, int mask_bytes,
                                  const double *weights_in, void *patterns_out, double *weights_out,
                                  int32_t *site_to_pattern, int64_t *n_patterns);

#ifdef __cplusplus
}
#endif
#endif /* PHYLO_ENGINE_H */
