/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the phylocaml tree-scoring hot path.
 *
 * This is a plain-C restatement of the algorithm the CUDA engine must reproduce. It is
 * used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs, as the checker. The product library (phylocaml_b200/lib) never links,
 * loads or calls anything in oracle/.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - oracle_compose            pinned against the reference's own compose_gtr/compose_sym
 *                               (lib/mlmodel.c:280-342) compiled in oracle/_ref.
 *   - oracle_fitch_median2 etc. pinned against the reference's bv_fitch / bv_distance
 *                               (lib/bitvector/bv.c:46-55,148-160) in oracle/_ref and
 *                               against the Fitch truth table of test/costMatrixTest.ml:83-108.
 *   - pruning / scaling / lnL / up-pass: the reference never wrote these
 *                               (lib/likelihood_c.ml:1-33 is all TODO) => PARITY UNPINNED by the
 *                               reference; pinned by our own independent checks instead
 *                               (brute-force state enumeration, closed forms, pulley principle).
 */
#ifndef PHYLO_ORACLE_H
#define PHYLO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One post-order step: node `parent` is the median of `left` and `right`, which hang off it
 * by branches of length t_left / t_right (ignored by Fitch). Mirrors the call
 * Node.median_2 makes per interior node of Tree.post_order_edges (lib/tree.ml:171-187,
 * lib/node.ml:183-198). Same memory layout as phylo_op in include/phylo_engine.h. */
typedef struct {
  int32_t parent, left, right, pad_;
  double t_left, t_right;
} oracle_op;

#define ORACLE_SCALE_EXP 256 /* rescale by 2^256 when the site maximum drops below 2^-256 */
#define ORACLE_LNL_BLOCK 1024

/* lib/mlmodel.c:280-342. D is the full n*n matrix with eigenvalues on the diagonal, as the
 * reference stores it (mlmodel.c:81-88). Ui == NULL selects the symmetric path, which
 * (like the reference, mlmodel.c:280) rounds t to float first. Row-major in and out. */
void oracle_compose(double *P, const double *U, const double *D, const double *Ui, double t, int n);

/* Canonical fixed-shape reduction of n doubles (see phylo_oracle.c). */
double oracle_reduce(const double *v, long n);
/* Level-1 only: writes ceil(n/1024) block partials. */
long oracle_reduce_blocks(const double *v, long n, double *partials);

/* Felsenstein pruning of a whole tree + root-edge log-likelihood.
 *  S states, K rate categories. U/D(full SxS)/Ui as held by MlModel.t (lib/mlModel.ml:53-63),
 *  Ui == NULL => symmetric model. pinvar < 0 => no invariant-sites class.
 *  tips: T*N masks of mask_bytes (1,2,4,8) each, tip-major; bit i set <=> state i possible.
 *  weights: N doubles or NULL (all 1).
 *  Nodes are numbered 0..n_nodes-1 with tips 0..T-1.
 *  clv_out (optional): n_nodes*N*K*S doubles, [node][pattern][k][i]; tips are filled too.
 *  scale_out (optional): n_nodes*N int32. site_lnl_out (optional): N doubles (unweighted).
 *  nthreads: patterns are split in contiguous 1024-aligned slabs over this many pthreads.
 *  Returns lnL. */
double oracle_lk_score_tree(int S, int K, const double *U, const double *D, const double *Ui,
                            const double *pi, const double *rates, const double *probs,
                            double pinvar, int T, long N, const void *tips, int mask_bytes,
                            const double *weights, const oracle_op *ops, int n_ops, int n_nodes,
                            int root_a, int root_b, double root_t, double *clv_out,
                            int32_t *scale_out, double *site_lnl_out, int nthreads);

/* One CLV update (the body of Likelihood.median_2): parent from two child CLVs. */
void oracle_lk_median2(int S, int K, long N, const double *Pl, const double *Pr,
                       const double *clv_l, const int32_t *sc_l, const double *clv_r,
                       const int32_t *sc_r, double *clv_p, int32_t *sc_p);

/* lib/bitvector/bv.c:148-160 restated for one-char-per-element vectors of elt_bytes. */
uint64_t oracle_fitch_median2(void *c, const void *a, const void *b, long chars, int elt_bytes);
/* lib/bitvector/bv.c:46-55 */
uint64_t oracle_fitch_distance(const void *a, const void *b, long chars, int elt_bytes);

/* Whole-tree Fitch down-pass: tree length = sum of interior
 * median costs + root-edge distance. weights (optional): N non-negative integers as doubles.
 * prelim_out (optional): n_nodes*N elements (tips copied in). node_cost_out (optional):
 * n_nodes uint64 (weighted). */
uint64_t oracle_fitch_score_tree(int T, long N, int elt_bytes, const void *tips,
                                 const double *weights, const oracle_op *ops, int n_ops,
                                 int n_nodes, int root_a, int root_b, void *prelim_out,
                                 uint64_t *node_cost_out, int nthreads);

/* Fitch up-pass (final state sets; rule of SURVEY.md section 8 a11). prelim: n_nodes*N elements
 * from the down-pass. final_out: n_nodes*N elements. Leaves keep their observed sets. */
void oracle_fitch_uppass(int T, long N, int elt_bytes, const void *prelim, const oracle_op *ops,
                         int n_ops, int n_nodes, int root_a, int root_b, void *final_out);

#ifdef __cplusplus
}
#endif
#endif
