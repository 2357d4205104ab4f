/* TEST INFRASTRUCTURE ONLY -- see phylo_oracle.h. Build with -ffp-contract=off (Makefile)
 * so every multiply-add below is two roundings, independent of the host compiler.
 *
 * Each function cites the reference file:line it restates, or "spec" = SURVEY.md Appendix C
 * where the reference has no implementation at all. */
#include "phylo_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ P(t) */

/* lib/mlmodel.c:280-302 (compose_sym) and :325-342 (compose_gtr), read row-major:
 *   gtr: P_ij = sum_k U[i][k] * (exp(D[k][k] t) * Ui[k][j])      (Ui != NULL)
 *   sym: P_ij = sum_k U[k][i] * (exp(D[k][k] t) * U[k][j])       (Ui == NULL)
 * t == -1.0 returns Q (:286-288, :332-334); t < EPSILON=1e-10 returns I (:299-301, :339-341).
 * compose_sym takes `const float t` (:280): t is rounded to float before use. */
void oracle_compose(double *P, const double *U, const double *D, const double *Ui, double t, int n)
{
  int i, j, k;
  double tt = t;
  if (Ui == NULL) tt = (double)(float)t;
  if (tt == -1.0 || tt >= 1e-10) {
    for (i = 0; i < n; ++i)
      for (j = 0; j < n; ++j) {
        double acc = 0.0;
        for (k = 0; k < n; ++k) {
          double lam = D[k * n + k];
          double e = (tt == -1.0) ? lam : exp(lam * tt); /* apply_exp, mlmodel.c:128-134 */
          if (Ui != NULL)
            acc += U[i * n + k] * (e * Ui[k * n + j]);
          else
            acc += U[k * n + i] * (e * U[k * n + j]);
        }
        P[i * n + j] = acc;
      }
  } else { /* create_identity, mlmodel.c:146-153 */
    for (i = 0; i < n; ++i)
      for (j = 0; j < n; ++j) P[i * n + j] = (i == j) ? 1.0 : 0.0;
  }
}

/* ------------------------------------------------- canonical reduction (spec C.2.6) */

/* One block of up to 1024 values, zero padded: 32 groups of 32 consecutive values are each
 * folded with offsets 16,8,4,2,1 (x[j] += x[j+off]), then the 32 group sums are folded the
 * same way. This is the shape a CUDA block gets from __shfl_down_sync within warps followed
 * by one warp over the per-warp sums, so the kernels reproduce it bit for bit. */
static double reduce_block(const double *v, long n)
{
  double w[32];
  int g, j, off;
  for (g = 0; g < 32; ++g) {
    double y[32];
    for (j = 0; j < 32; ++j) {
      long idx = (long)g * 32 + j;
      y[j] = idx < n ? v[idx] : 0.0;
    }
    for (off = 16; off >= 1; off >>= 1)
      for (j = 0; j < off; ++j) y[j] += y[j + off];
    w[g] = y[0];
  }
  for (off = 16; off >= 1; off >>= 1)
    for (j = 0; j < off; ++j) w[j] += w[j + off];
  return w[0];
}

long oracle_reduce_blocks(const double *v, long n, double *partials)
{
  long nb = (n + ORACLE_LNL_BLOCK - 1) / ORACLE_LNL_BLOCK, b;
  for (b = 0; b < nb; ++b) {
    long lo = b * ORACLE_LNL_BLOCK;
    long len = n - lo < ORACLE_LNL_BLOCK ? n - lo : ORACLE_LNL_BLOCK;
    partials[b] = reduce_block(v + lo, len);
  }
  return nb;
}

/* Repeated level by level until one value is left. */
double oracle_reduce(const double *v, long n)
{
  double *cur, *nxt, r;
  long m;
  if (n <= 0) return 0.0;
  if (n <= ORACLE_LNL_BLOCK) return reduce_block(v, n);
  m = (n + ORACLE_LNL_BLOCK - 1) / ORACLE_LNL_BLOCK;
  cur = (double *)malloc(sizeof(double) * (size_t)m);
  oracle_reduce_blocks(v, n, cur);
  while (m > ORACLE_LNL_BLOCK) {
    long m2 = (m + ORACLE_LNL_BLOCK - 1) / ORACLE_LNL_BLOCK;
    nxt = (double *)malloc(sizeof(double) * (size_t)m2);
    oracle_reduce_blocks(cur, m, nxt);
    free(cur);
    cur = nxt;
    m = m2;
  }
  r = reduce_block(cur, m);
  free(cur);
  return r;
}

/* ------------------------------------------------------------ likelihood (spec C.2) */

static uint64_t load_mask(const void *p, long idx, int bytes)
{
  switch (bytes) {
  case 1: return ((const uint8_t *)p)[idx];
  case 2: return ((const uint16_t *)p)[idx];
  case 4: return ((const uint32_t *)p)[idx];
  default: return ((const uint64_t *)p)[idx];
  }
}

static void store_mask(void *p, long idx, int bytes, uint64_t v)
{
  switch (bytes) {
  case 1: ((uint8_t *)p)[idx] = (uint8_t)v; break;
  case 2: ((uint16_t *)p)[idx] = (uint16_t)v; break;
  case 4: ((uint32_t *)p)[idx] = (uint32_t)v; break;
  default: ((uint64_t *)p)[idx] = v; break;
  }
}

/* spec C.2.3 + C.2.4: X = P_l L_l, Y = P_r L_r, L_p = X*Y (ascending j, no FMA), then if the
 * site maximum is below 2^-256 multiply by 2^256 and bump the counter. */
void oracle_lk_median2(int S, int K, long N, const double *Pl, const double *Pr,
                       const double *clv_l, const int32_t *sc_l, const double *clv_r,
                       const int32_t *sc_r, double *clv_p, int32_t *sc_p)
{
  const double thresh = ldexp(1.0, -ORACLE_SCALE_EXP), factor = ldexp(1.0, ORACLE_SCALE_EXP);
  long s;
  int k, i, j;
  for (s = 0; s < N; ++s) {
    const double *ll = clv_l + (size_t)s * K * S, *lr = clv_r + (size_t)s * K * S;
    double *lp = clv_p + (size_t)s * K * S;
    double m = 0.0;
    int first = 1;
    for (k = 0; k < K; ++k) {
      const double *pl = Pl + (size_t)k * S * S, *pr = Pr + (size_t)k * S * S;
      for (i = 0; i < S; ++i) {
        double x = 0.0, y = 0.0, v;
        for (j = 0; j < S; ++j) x += pl[i * S + j] * ll[k * S + j];
        for (j = 0; j < S; ++j) y += pr[i * S + j] * lr[k * S + j];
        v = x * y;
        lp[k * S + i] = v;
        if (first || v > m) { m = v; first = 0; }
      }
    }
    sc_p[s] = sc_l[s] + sc_r[s];
    if (m < thresh) {
      for (k = 0; k < K * S; ++k) lp[k] *= factor;
      sc_p[s] += 1;
    }
  }
}

typedef struct {
  int S, K, T, mask_bytes, n_ops, n_nodes, root_a, root_b;
  long N, lo, hi;
  const double *pi, *probs, *weights, *Pmats, *Proot;
  double pinvar;
  const void *tips;
  const oracle_op *ops;
  double *clv, *site_lnl, *wsite;
  int32_t *scale;
} lk_job;

/* spec C.2.5 with an invariant-sites class: ln((1-v) l 2^(-256 c) + v pv), evaluated in the scaled
 * domain as ln((1-v) l + v pv 2^(256 c)) - c 256 ln 2 so that a summed scale counter c >= 4-5
 * (site likelihood below 2^-1074: deep trees) does not underflow a variable site to -inf. When
 * v pv 2^(256 c) would overflow it exceeds (1-v) l <= 1 by more than 2^64: the sum is v pv to
 * rounding. Same statement as lnl_pinvar in phylocaml_b200/csrc/common.cuh. */
static double lnl_pinvar(double l, int c, double pinvar, double pv)
{
  const double ln_scale = ORACLE_SCALE_EXP * 0.6931471805599453094;
  const double inv_term = pinvar * pv, var = (1.0 - pinvar) * l;
  if (inv_term == 0.0) return log(var) - (double)c * ln_scale;
  if (c >= 4) {
    int ex;
    (void)frexp(inv_term, &ex);
    if (ex + ORACLE_SCALE_EXP * c > 64) return log(inv_term);
  }
  return log(var + ldexp(inv_term, ORACLE_SCALE_EXP * c)) - (double)c * ln_scale;
}

static void *lk_worker(void *arg)
{
  lk_job *J = (lk_job *)arg;
  const int S = J->S, K = J->K;
  const size_t C = (size_t)K * S;
  const long N = J->N, lo = J->lo, n = J->hi - J->lo;
  const double ln_scale = (double)ORACLE_SCALE_EXP * 0.6931471805599453094;
  long s;
  int t, k, i, j, o;
  if (n <= 0) return NULL;
  /* spec C.2.1: tip partials straight from the bit masks (lib/alphabet.ml:193-196: state i
   * <-> bit 1<<i); scale counter 0 */
  for (t = 0; t < J->T; ++t)
    for (s = lo; s < J->hi; ++s) {
      uint64_t m = load_mask(J->tips, (long)t * N + s, J->mask_bytes);
      double *L = J->clv + ((size_t)t * N + s) * C;
      for (k = 0; k < K; ++k)
        for (i = 0; i < S; ++i) L[k * S + i] = ((m >> i) & 1) ? 1.0 : 0.0;
      J->scale[(size_t)t * N + s] = 0;
    }
  /* spec C.2.3-4, in schedule order (lib/tree.ml:171-187 post-order) */
  for (o = 0; o < J->n_ops; ++o) {
    const oracle_op *op = &J->ops[o];
    const double *Pl = J->Pmats + (size_t)(2 * o) * K * S * S;
    const double *Pr = J->Pmats + (size_t)(2 * o + 1) * K * S * S;
    oracle_lk_median2(S, K, n, Pl, Pr, J->clv + ((size_t)op->left * N + lo) * C,
                      J->scale + (size_t)op->left * N + lo,
                      J->clv + ((size_t)op->right * N + lo) * C,
                      J->scale + (size_t)op->right * N + lo,
                      J->clv + ((size_t)op->parent * N + lo) * C,
                      J->scale + (size_t)op->parent * N + lo);
  }
  /* spec C.2.5: join the two directed CLVs across the root edge */
  for (s = lo; s < J->hi; ++s) {
    const double *La = J->clv + ((size_t)J->root_a * N + s) * C;
    const double *Lb = J->clv + ((size_t)J->root_b * N + s) * C;
    int32_t c = J->scale[(size_t)J->root_a * N + s] + J->scale[(size_t)J->root_b * N + s];
    double l = 0.0, lnl;
    for (k = 0; k < K; ++k) {
      const double *P = J->Proot + (size_t)k * S * S;
      double lk = 0.0;
      for (i = 0; i < S; ++i) {
        double y = 0.0;
        for (j = 0; j < S; ++j) y += P[i * S + j] * Lb[k * S + j];
        lk += (J->pi[i] * La[k * S + i]) * y;
      }
      l += J->probs[k] * lk;
    }
    if (J->pinvar >= 0.0) {
      /* invariant-sites class (lib/mlModel.ml:56,684-694): a pattern can be invariant in
       * state i iff every tip mask has bit i */
      uint64_t inv = ~(uint64_t)0;
      double pinv = 0.0;
      for (t = 0; t < J->T; ++t) inv &= load_mask(J->tips, (long)t * N + s, J->mask_bytes);
      for (i = 0; i < S; ++i)
        if ((inv >> i) & 1) pinv += J->pi[i];
      lnl = lnl_pinvar(l, c, J->pinvar, pinv);
    } else {
      lnl = log(l) - (double)c * ln_scale;
    }
    J->site_lnl[s] = lnl;
    J->wsite[s] = (J->weights ? J->weights[s] : 1.0) * lnl;
  }
  return NULL;
}

static void split_slabs(long N, int nthreads, long align, long *lo, long *hi)
{
  long blocks = (N + align - 1) / align, per = blocks / nthreads, rem = blocks % nthreads;
  long b = 0;
  int t;
  for (t = 0; t < nthreads; ++t) {
    long nb = per + (t < rem ? 1 : 0);
    lo[t] = b * align < N ? b * align : N;
    b += nb;
    hi[t] = b * align < N ? b * align : N;
  }
}

double oracle_lk_score_tree(int S, int K, const double *U, const double *D, const double *Ui,
                            const double *pi, const double *rates, const double *probs,
                            double pinvar, int T, long N, const void *tips, int mask_bytes,
                            const double *weights, const oracle_op *ops, int n_ops, int n_nodes,
                            int root_a, int root_b, double root_t, double *clv_out,
                            int32_t *scale_out, double *site_lnl_out, int nthreads)
{
  const size_t C = (size_t)K * S, SS = (size_t)S * S;
  double *Pmats = (double *)malloc(sizeof(double) * (size_t)(2 * n_ops + 1) * K * SS);
  double *clv = clv_out ? clv_out : (double *)malloc(sizeof(double) * (size_t)n_nodes * N * C);
  int32_t *scale = scale_out ? scale_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes * N);
  double *site = site_lnl_out ? site_lnl_out : (double *)malloc(sizeof(double) * (size_t)N);
  double *wsite = (double *)malloc(sizeof(double) * (size_t)N);
  pthread_t *th;
  lk_job *jobs;
  long *lo, *hi;
  double lnl;
  int o, k, t;
  if (nthreads < 1) nthreads = 1;
  /* spec C.2.2: tau = t * rates[k]; P_k = reference compose at tau */
  for (o = 0; o < n_ops; ++o)
    for (k = 0; k < K; ++k) {
      oracle_compose(Pmats + ((size_t)(2 * o) * K + k) * SS, U, D, Ui, ops[o].t_left * rates[k], S);
      oracle_compose(Pmats + ((size_t)(2 * o + 1) * K + k) * SS, U, D, Ui, ops[o].t_right * rates[k], S);
    }
  for (k = 0; k < K; ++k)
    oracle_compose(Pmats + ((size_t)(2 * n_ops) * K + k) * SS, U, D, Ui, root_t * rates[k], S);

  th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  jobs = (lk_job *)malloc(sizeof(lk_job) * nthreads);
  lo = (long *)malloc(sizeof(long) * nthreads);
  hi = (long *)malloc(sizeof(long) * nthreads);
  split_slabs(N, nthreads, ORACLE_LNL_BLOCK, lo, hi);
  for (t = 0; t < nthreads; ++t) {
    lk_job *J = &jobs[t];
    J->S = S; J->K = K; J->T = T; J->mask_bytes = mask_bytes; J->n_ops = n_ops;
    J->n_nodes = n_nodes; J->root_a = root_a; J->root_b = root_b;
    J->N = N; J->lo = lo[t]; J->hi = hi[t];
    J->pi = pi; J->probs = probs; J->weights = weights; J->Pmats = Pmats;
    J->Proot = Pmats + (size_t)(2 * n_ops) * K * SS;
    J->pinvar = pinvar; J->tips = tips; J->ops = ops;
    J->clv = clv; J->site_lnl = site; J->wsite = wsite; J->scale = scale;
    if (nthreads == 1) lk_worker(J);
    else pthread_create(&th[t], NULL, lk_worker, J);
  }
  if (nthreads > 1)
    for (t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  /* spec C.2.6 */
  lnl = oracle_reduce(wsite, N);
  free(th); free(jobs); free(lo); free(hi); free(wsite); free(Pmats);
  if (!clv_out) free(clv);
  if (!scale_out) free(scale);
  if (!site_lnl_out) free(site);
  return lnl;
}

/* ------------------------------------------------------------------ Fitch (spec C.4) */

/* lib/bitvector/bv.c:148-160 (bv_fitch); identical rule in lib/nonAdditive_c.ml:19-35 */
uint64_t oracle_fitch_median2(void *c, const void *a, const void *b, long chars, int elt_bytes)
{
  uint64_t res = 0;
  long i;
  for (i = 0; i < chars; ++i) {
    uint64_t x = load_mask(a, i, elt_bytes), y = load_mask(b, i, elt_bytes), m = x & y;
    if (m == 0) { m = x | y; ++res; }
    store_mask(c, i, elt_bytes, m);
  }
  return res;
}

/* lib/bitvector/bv.c:46-55 (bv_distance) */
uint64_t oracle_fitch_distance(const void *a, const void *b, long chars, int elt_bytes)
{
  uint64_t res = 0;
  long i;
  for (i = 0; i < chars; ++i)
    if ((load_mask(a, i, elt_bytes) & load_mask(b, i, elt_bytes)) == 0) ++res;
  return res;
}

typedef struct {
  int T, elt_bytes, n_ops, root_a, root_b;
  long N, lo, hi;
  const void *tips;
  const double *weights;
  const oracle_op *ops;
  uint8_t *prelim;
  uint64_t *node_cost; /* n_ops + 1 per thread */
} fitch_job;

static void *fitch_worker(void *arg)
{
  fitch_job *J = (fitch_job *)arg;
  const int eb = J->elt_bytes;
  const long N = J->N, lo = J->lo, n = J->hi - J->lo;
  int o, t;
  long i;
  for (o = 0; o <= J->n_ops; ++o) J->node_cost[o] = 0;
  if (n <= 0) return NULL;
  for (t = 0; t < J->T; ++t) /* leaves: median_1 = identity, cost 0 (nonAdditive_c.ml:18) */
    memcpy(J->prelim + ((size_t)t * N + lo) * eb, (const uint8_t *)J->tips + ((size_t)t * N + lo) * eb,
           (size_t)n * eb);
  for (o = 0; o <= J->n_ops; ++o) {
    const uint8_t *a, *b;
    uint8_t *c = NULL;
    if (o < J->n_ops) {
      a = J->prelim + ((size_t)J->ops[o].left * N + lo) * eb;
      b = J->prelim + ((size_t)J->ops[o].right * N + lo) * eb;
      c = J->prelim + ((size_t)J->ops[o].parent * N + lo) * eb;
    } else { /* root-edge join = bv_distance */
      a = J->prelim + ((size_t)J->root_a * N + lo) * eb;
      b = J->prelim + ((size_t)J->root_b * N + lo) * eb;
    }
    if (J->weights == NULL) {
      J->node_cost[o] = c ? oracle_fitch_median2(c, a, b, n, eb) : oracle_fitch_distance(a, b, n, eb);
    } else { /* nonAdditive_c.ml:31: cost += weight * change */
      uint64_t res = 0;
      for (i = 0; i < n; ++i) {
        uint64_t x = load_mask(a, i, eb), y = load_mask(b, i, eb), m = x & y;
        if (m == 0) { m = x | y; res += (uint64_t)J->weights[lo + i]; }
        if (c) store_mask(c, i, eb, m);
      }
      J->node_cost[o] = res;
    }
  }
  return NULL;
}

uint64_t oracle_fitch_score_tree(int T, long N, int elt_bytes, const void *tips,
                                 const double *weights, const oracle_op *ops, int n_ops,
                                 int n_nodes, int root_a, int root_b, void *prelim_out,
                                 uint64_t *node_cost_out, int nthreads)
{
  uint8_t *prelim = prelim_out ? (uint8_t *)prelim_out
                               : (uint8_t *)malloc((size_t)n_nodes * N * elt_bytes);
  pthread_t *th;
  fitch_job *jobs;
  uint64_t *costs, total = 0;
  long *lo, *hi;
  int t, o;
  if (nthreads < 1) nthreads = 1;
  th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  jobs = (fitch_job *)malloc(sizeof(fitch_job) * nthreads);
  costs = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nthreads * (n_ops + 1));
  lo = (long *)malloc(sizeof(long) * nthreads);
  hi = (long *)malloc(sizeof(long) * nthreads);
  split_slabs(N, nthreads, 128, lo, hi);
  for (t = 0; t < nthreads; ++t) {
    fitch_job *J = &jobs[t];
    J->T = T; J->elt_bytes = elt_bytes; J->n_ops = n_ops; J->root_a = root_a; J->root_b = root_b;
    J->N = N; J->lo = lo[t]; J->hi = hi[t]; J->tips = tips; J->weights = weights; J->ops = ops;
    J->prelim = prelim; J->node_cost = costs + (size_t)t * (n_ops + 1);
    if (nthreads == 1) fitch_worker(J);
    else pthread_create(&th[t], NULL, fitch_worker, J);
  }
  if (nthreads > 1)
    for (t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  if (node_cost_out) memset(node_cost_out, 0, sizeof(uint64_t) * (size_t)n_nodes);
  for (t = 0; t < nthreads; ++t)
    for (o = 0; o <= n_ops; ++o) {
      uint64_t c = costs[(size_t)t * (n_ops + 1) + o];
      total += c;
      if (node_cost_out && o < n_ops) node_cost_out[ops[o].parent] += c;
    }
  free(th); free(jobs); free(costs); free(lo); free(hi);
  if (!prelim_out) free(prelim);
  return total;
}

/* Up-pass, spec section 8 a11 / C.4 (no reference implementation: lib/node.ml:260-268 TODO).
 * Root set R = A&B if non-empty else A|B. Interior node with parent final A, prelim P and
 * children prelims L,R:  (P&A)==A -> A;  else (L&R)==0 -> P|A;  else P|(A&(L|R)).
 * Leaves keep their observed set. The schedule is walked in reverse (parents first). */
void oracle_fitch_uppass(int T, long N, int elt_bytes, const void *prelim, const oracle_op *ops,
                         int n_ops, int n_nodes, int root_a, int root_b, void *final_out)
{
  const int eb = elt_bytes;
  int *parent_of = (int *)malloc(sizeof(int) * n_nodes);
  int o, v;
  long i;
  for (v = 0; v < n_nodes; ++v) parent_of[v] = -1;
  for (o = 0; o < n_ops; ++o) { parent_of[ops[o].left] = ops[o].parent; parent_of[ops[o].right] = ops[o].parent; }
  memcpy(final_out, prelim, (size_t)n_nodes * N * eb); /* leaves (and unreached nodes) = prelim */
  for (o = n_ops - 1; o >= 0; --o) {
    const int p = ops[o].parent, l = ops[o].left, r = ops[o].right;
    for (i = 0; i < N; ++i) {
      uint64_t P = load_mask(prelim, (long)p * N + i, eb), L = load_mask(prelim, (long)l * N + i, eb),
               R = load_mask(prelim, (long)r * N + i, eb), A, F;
      if (p == root_a || p == root_b) {
        uint64_t a = load_mask(prelim, (long)root_a * N + i, eb), b = load_mask(prelim, (long)root_b * N + i, eb);
        A = (a & b) ? (a & b) : (a | b);
      } else {
        A = load_mask(final_out, (long)parent_of[p] * N + i, eb);
      }
      if ((P & A) == A) F = A;
      else if ((L & R) == 0) F = P | A;
      else F = P | (A & (L | R));
      store_mask(final_out, (long)p * N + i, eb, F);
    }
  }
  (void)T;
  free(parent_of);
}
