/* phylo_stubs.c -- OCaml `value` stubs over the C ABI of include/phylo_engine.h.
 *
 * These are the functions phylocaml's `external` declarations bind. They follow the
 * reference's own FFI style (lib/mlmodel.c:200-205, :304-365; lib/bitvector/bv.c:251-480):
 * CAMLparam/CAMLreturn, Bigarray c_layout float64 payloads read with Data_bigarray_val,
 * opaque native state in custom blocks with a finalizer, errors as `Failure msg`
 * (failwith). Differences, all deliberate: the runtime lock is released while the GPU
 * works; compose results are BIGARRAY_MANAGED (the reference leaks them,
 * lib/mlmodel.c:320,362); device node data is named by small integer slots owned by the
 * engine, and every interior node value the plugins return is a `node` custom block whose
 * finalizer gives the slot back (phylo_lk_node_release / phylo_fitch_node_release) -- the
 * pattern of the reference's bitvector blocks (lib/bitvector/bv.c:183-189 bv_CAML_free,
 * :229-244 ops table). An engine stays alive until its own block AND every node block that
 * names it have been finalized (reference count), whatever order the GC picks.
 *
 * Link-compatible names kept from the reference (lib/mlmodel.h:39-43):
 *   likelihood_CAML_diagonalize_sym / _gtr, likelihood_CAML_compose_sym / _gtr.
 * New names follow the same <module>_CAML_<fn> scheme: likelihood_CAML_*, nonadd_CAML_*.
 * Build: part of libphyloc (libphyloc.clib:1-7), linked with -lphyloc_b200.
 */
#include <stdlib.h>
#include <string.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "phylo_engine.h"

/* ------------------------------------------------------------ engine custom block ---- */
/* The engine handle is shared by its own custom block and by every node block: a reference count
 * decides who destroys it (the GC finalizes unreachable blocks in no particular order). */
struct engine_ref {
  phylo_engine *e;
  long refs;
};
#define Eref_val(v) (*((struct engine_ref **)Data_custom_val(v)))
#define Engine_val(v) (Eref_val(v)->e)

static void eref_drop(struct engine_ref *r)
{
  if (r && --r->refs == 0) {
    phylo_engine_destroy(r->e);
    free(r);
  }
}

static void engine_finalize(value v)
{
  eref_drop(Eref_val(v));
  Eref_val(v) = NULL;
}

static struct custom_operations engine_ops = {
    "AMNH/phylo_b200/engine/0.2", engine_finalize, custom_compare_default, custom_hash_default,
    custom_serialize_default, custom_deserialize_default, custom_compare_ext_default};

static void check(phylo_engine *e, int rc)
{
  if (rc != PHYLO_OK) caml_failwith(phylo_last_error(e)); /* -> OCaml Failure */
}

/* external engine_create : int -> engine = "phylo_CAML_engine_create" */
CAMLprim value phylo_CAML_engine_create(value vdev)
{
  CAMLparam1(vdev);
  CAMLlocal1(res);
  phylo_engine *e = NULL;
  struct engine_ref *r;
  int rc = phylo_engine_create(Int_val(vdev), &e);
  if (rc != PHYLO_OK) caml_failwith(phylo_last_error(NULL));
  r = (struct engine_ref *)malloc(sizeof(*r));
  r->e = e;
  r->refs = 1;
  res = caml_alloc_custom(&engine_ops, sizeof(struct engine_ref *), 0, 1);
  Eref_val(res) = r;
  CAMLreturn(res);
}

/* external engine_set_option : engine -> int -> int -> unit = "phylo_CAML_engine_set_option"
 * (PHYLO_OPT_*: 2 = retain interior CLVs / state sets (1, default) or score only (0: lnL / Fitch length alone,
 * the fastest form of a candidate evaluation), 1 = tree-fused kernels, 3 = whole-tree Fitch kernel) */
CAMLprim value phylo_CAML_engine_set_option(value ve, value vopt, value vval)
{
  CAMLparam3(ve, vopt, vval);
  phylo_engine *e = Engine_val(ve);
  check(e, phylo_engine_set_option(e, Int_val(vopt), (int64_t)Long_val(vval)));
  CAMLreturn(Val_unit);
}

/* one process-wide default engine for the reference-named entry points that carry no
 * engine argument (compose_*): created on first use on device 0 */
static phylo_engine *default_engine(void)
{
  static phylo_engine *e = NULL;
  if (!e && phylo_engine_create(0, &e) != PHYLO_OK) caml_failwith(phylo_last_error(NULL));
  return e;
}

/* ------------------------------------------------------------- node custom block ---- */
/* A node value = one device-resident CLV (kind 0) or state-set vector (kind 1), named by a slot of
 * its engine. Interior nodes own their slot: the finalizer returns it, the engine hands it (and
 * the device buffer still attached to it) to the next median. Tips (slot < n_taxa) are not owned.
 * Like the reference's bitvector blocks, comparison is by content for state sets (bv_CAML_compare_values,
 * lib/bitvector/bv.c:191-194). Serialisation is not supported (the data live in HBM; copy them out
 * with get_clv / get_states first). */
struct node_blk {
  struct engine_ref *er;
  int32_t slot;
  int16_t kind;   /* 0 = likelihood CLV, 1 = Fitch / bitvector state sets */
  int16_t owned;  /* 1: release the slot on finalization */
  uint64_t gen;   /* generation of the loaded alignment the slot belongs to */
};
#define Node_val(v) ((struct node_blk *)Data_custom_val(v))
#define Node_engine(v) (Node_val(v)->er->e)
#define Node_slot(v) (Node_val(v)->slot)

static void node_finalize(value v)
{
  struct node_blk *n = Node_val(v);
  if (!n->er) return;
  if (n->owned) {
    if (n->kind == 0) phylo_lk_node_release(n->er->e, n->slot, n->gen);
    else phylo_fitch_node_release(n->er->e, n->slot, n->gen);
  }
  eref_drop(n->er);
  n->er = NULL;
}

static int node_compare(value a, value b)
{
  struct node_blk *x = Node_val(a), *y = Node_val(b);
  int r = 0;
  if (x->kind == 1 && y->kind == 1 && x->er == y->er && phylo_bv_compare(x->er->e, x->slot, y->slot, &r) == PHYLO_OK)
    return r;
  return (x->slot > y->slot) - (x->slot < y->slot);
}

static struct custom_operations node_ops = {
    "AMNH/phylo_b200/node/0.2", node_finalize, node_compare, custom_hash_default,
    custom_serialize_default, custom_deserialize_default, custom_compare_ext_default};

/* GC pressure: one node block stands for N*K*S*8 bytes of HBM that the OCaml heap does not see.
 * caml_alloc_custom(mem = 1, max = node_gc_max) makes a major collection due after about that
 * many node allocations -- the knob the reference exposes as bv_CAML_custom_max
 * (lib/bitvector/bv.c:250-256, default 10000 there). */
static int node_gc_max = 256;

/* external custom_max : int -> unit = "phylo_CAML_custom_max" */
CAMLprim value phylo_CAML_custom_max(value n)
{
  CAMLparam1(n);
  if (Int_val(n) > 0) node_gc_max = Int_val(n);
  CAMLreturn(Val_unit);
}

extern value caml_gc_full_major(value unit); /* runtime primitive behind Gc.full_major */

static value node_wrap(value ve, int kind, int slot, int owned, uint64_t gen)
{
  CAMLparam1(ve);
  CAMLlocal1(res);
  struct node_blk *n;
  res = caml_alloc_custom(&node_ops, sizeof(struct node_blk), owned ? 1 : 0, node_gc_max);
  n = Node_val(res);
  n->er = Eref_val(ve);
  n->er->refs++;
  n->slot = slot;
  n->kind = (int16_t)kind;
  n->owned = (int16_t)owned;
  n->gen = gen;
  CAMLreturn(res);
}

/* A fresh interior slot. When every slot of the table is handed out, dead node values may simply
 * not have been collected yet: run a full major GC first (their finalizers release slots) and only
 * let the engine grow its table if the slots are genuinely all live. */
static int slot_new(value ve, int kind, uint64_t *gen)
{
  CAMLparam1(ve);
  phylo_engine *e = Engine_val(ve);
  int cap = 0, used = 0, slot = -1, rc;
  if (kind == 0) phylo_lk_node_stats(e, &cap, &used, NULL);
  else phylo_fitch_node_stats(e, &cap, &used, NULL);
  if (cap > 0 && used >= cap) caml_gc_full_major(Val_unit);
  rc = kind == 0 ? phylo_lk_node_alloc(e, &slot, gen) : phylo_fitch_node_alloc(e, &slot, gen);
  check(e, rc);
  CAMLreturnT(int, slot);
}

static void slot_drop(phylo_engine *e, int kind, int slot, uint64_t gen)
{
  if (kind == 0) phylo_lk_node_release(e, slot, gen);
  else phylo_fitch_node_release(e, slot, gen);
}

/* external node_slot : node -> int = "phylo_CAML_node_slot"   (diagnostics, to_string) */
CAMLprim value phylo_CAML_node_slot(value vn)
{
  CAMLparam1(vn);
  CAMLreturn(Val_int(Node_slot(vn)));
}

/* external node_stats : engine -> bool -> int * int * int = "phylo_CAML_node_stats"
 * (fitch?) -> (interior slots in the table, slots handed out, interior slots owning a device buffer) */
CAMLprim value phylo_CAML_node_stats(value ve, value vfitch)
{
  CAMLparam2(ve, vfitch);
  CAMLlocal1(res);
  int cap = 0, used = 0, buf = 0;
  if (Int_val(vfitch)) phylo_fitch_node_stats(Engine_val(ve), &cap, &used, &buf);
  else phylo_lk_node_stats(Engine_val(ve), &cap, &used, &buf);
  res = caml_alloc_tuple(3);
  Store_field(res, 0, Val_int(cap));
  Store_field(res, 1, Val_int(used));
  Store_field(res, 2, Val_int(buf));
  CAMLreturn(res);
}

/* ------------------------------------------ MlModel externs (lib/mlModel.ml:73-90) ---- */
/* replaces lib/mlmodel.c:200-205 */
CAMLprim value likelihood_CAML_diagonalize_sym(value Q, value D)
{
  CAMLparam2(Q, D);
  if (phylo_diagonalize_sym((double *)Data_bigarray_val(Q), (double *)Data_bigarray_val(D),
                            (int)Bigarray_val(Q)->dim[0]) != PHYLO_OK)
    caml_failwith("dsyev_ diagonalization failed to converge. Singular matrix?");
  CAMLreturn(Val_unit);
}

/* replaces lib/mlmodel.c:265-271 */
CAMLprim value likelihood_CAML_diagonalize_gtr(value Q, value D, value Qi)
{
  CAMLparam3(Q, D, Qi);
  if (phylo_diagonalize_gtr((double *)Data_bigarray_val(Q), (double *)Data_bigarray_val(D),
                            (double *)Data_bigarray_val(Qi), (int)Bigarray_val(Q)->dim[0]) != PHYLO_OK)
    caml_failwith("Imaginary eigenvalues");
  CAMLreturn(Val_unit);
}

static value compose_common(value U, value D, value Ui, double t)
{
  CAMLparam3(U, D, Ui);
  CAMLlocal1(res);
  intptr_t dims[2];
  phylo_engine *e = default_engine();
  int n = (int)Bigarray_val(U)->dim[0], rc;
  dims[0] = n;
  dims[1] = n;
  /* runtime-allocated and owned by the GC (the reference leaks P: no BIGARRAY_MANAGED) */
  res = caml_ba_alloc(CAML_BA_FLOAT64 | CAML_BA_C_LAYOUT, 2, NULL, dims);
  if (Ui == Val_unit)
    rc = phylo_compose_sym(e, (double *)Data_bigarray_val(U), (double *)Data_bigarray_val(D), t, n,
                           (double *)Data_bigarray_val(res));
  else
    rc = phylo_compose_gtr(e, (double *)Data_bigarray_val(U), (double *)Data_bigarray_val(D),
                           (double *)Data_bigarray_val(Ui), t, n, (double *)Data_bigarray_val(res));
  check(e, rc);
  CAMLreturn(res);
}

/* replaces lib/mlmodel.c:304-323 */
CAMLprim value likelihood_CAML_compose_sym(value U, value D, value t)
{
  return compose_common(U, D, Val_unit, Double_val(t));
}

/* replaces lib/mlmodel.c:344-365 */
CAMLprim value likelihood_CAML_compose_gtr(value U, value D, value Ui, value t)
{
  return compose_common(U, D, Ui, Double_val(t));
}

/* --------------------------------------------------- Likelihood_c (lib/nodeData.ml) ---- */
/* external set_model : engine -> u:matrix -> d:matrix -> ui:matrix option ->
 *                      (priors:vector * rates:vector * probs:vector * pinvar:float option) -> unit
 * carries MlModel.t (lib/mlModel.ml:53-63) */
CAMLprim value likelihood_CAML_set_model(value ve, value U, value D, value Uio, value rest)
{
  CAMLparam5(ve, U, D, Uio, rest);
  phylo_engine *e = Engine_val(ve);
  value pri = Field(rest, 0), rates = Field(rest, 1), probs = Field(rest, 2), pinv = Field(rest, 3);
  const double *ui = (Uio == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(Uio, 0));
  double pv = (pinv == Val_int(0)) ? -1.0 : Double_val(Field(pinv, 0));
  check(e, phylo_lk_set_model(e, (int)Bigarray_val(U)->dim[0], (int)Bigarray_val(rates)->dim[0],
                              (double *)Data_bigarray_val(U), (double *)Data_bigarray_val(D), ui,
                              (double *)Data_bigarray_val(pri), (double *)Data_bigarray_val(rates),
                              (double *)Data_bigarray_val(probs), pv));
  CAMLreturn(Val_unit);
}

/* external set_tips : engine -> (int, int8_unsigned_elt, c_layout) Array2.t (taxa x patterns)
 *                     -> weights:vector option -> capacity:int -> unit */
CAMLprim value likelihood_CAML_set_tips(value ve, value masks, value wo, value vcap)
{
  CAMLparam4(ve, masks, wo, vcap);
  phylo_engine *e = Engine_val(ve);
  struct caml_ba_array *b = Bigarray_val(masks);
  int kind = (int)(b->flags & 0xff), bytes = 1, rc;
  const double *w = (wo == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(wo, 0));
  if (kind == CAML_BA_INT32) bytes = 4;
  else if (kind == CAML_BA_INT64) bytes = 8;
  else if (kind != CAML_BA_UINT8) caml_failwith("likelihood set_tips: masks must be uint8, int32 or int64");
  caml_release_runtime_system(); /* H2D of the alignment can take milliseconds */
  rc = phylo_lk_set_tips(e, (int)b->dim[0], (int64_t)b->dim[1], b->data, bytes, w, Int_val(vcap));
  caml_acquire_runtime_system();
  check(e, rc);
  CAMLreturn(Val_unit);
}

/* external tip : engine -> int -> node = "likelihood_CAML_tip"   (taxon i's node value; not owned) */
CAMLprim value likelihood_CAML_tip(value ve, value vi)
{
  CAMLparam2(ve, vi);
  CAMLreturn(node_wrap(ve, 0, Int_val(vi), 0, 0));
}

/* external median_2 : engine -> node -> float -> node -> float -> node
 * body of Likelihood_c.median_2 (lib/nodeData.ml:21): a NEW node value, the CLV of the parent of
 * (left over a branch of t_left, right over t_right) */
CAMLprim value likelihood_CAML_median2(value ve, value l, value tl, value r, value tr)
{
  CAMLparam5(ve, l, tl, r, tr);
  phylo_engine *e = Engine_val(ve);
  uint64_t gen = 0;
  int slot = slot_new(ve, 0, &gen), rc;
  rc = phylo_lk_median_2(e, slot, Node_slot(l), Double_val(tl), Node_slot(r), Double_val(tr));
  if (rc != PHYLO_OK) slot_drop(e, 0, slot, gen);
  check(e, rc);
  CAMLreturn(node_wrap(ve, 0, slot, 1, gen));
}

/* external median_3 : engine -> (node * float) -> (node * float) -> (node * float) -> node
 * Likelihood_c.median_3 (lib/nodeData.ml:22): the CLV conditioned on all three neighbours */
CAMLprim value likelihood_CAML_median3(value ve, value a, value b, value c)
{
  CAMLparam4(ve, a, b, c);
  phylo_engine *e = Engine_val(ve);
  uint64_t gen = 0;
  int slot = slot_new(ve, 0, &gen), rc;
  rc = phylo_lk_median_3(e, slot, Node_slot(Field(a, 0)), Double_val(Field(a, 1)), Node_slot(Field(b, 0)),
                         Double_val(Field(b, 1)), Node_slot(Field(c, 0)), Double_val(Field(c, 1)));
  if (rc != PHYLO_OK) slot_drop(e, 0, slot, gen);
  check(e, rc);
  CAMLreturn(node_wrap(ve, 0, slot, 1, gen));
}

/* external score_tree : engine -> (int32, int32_elt, c_layout) Array2.t (n_ops x 3: parent,left,right;
 *                       ids < n_taxa are tips, the others name the tree's interior nodes in any numbering)
 *                       -> (float, float64_elt, c_layout) Array2.t (n_ops x 2: t_left,t_right)
 *                       -> (root_a:int * root_b:int * root_t:float) -> float * node array
 * One launch sequence for a whole tree (Tree.post_order_edges flattened, lib/tree.ml:171-187). Every
 * interior id gets a fresh slot; the result is lnL and the node values in schedule order. */
CAMLprim value likelihood_CAML_score_tree(value ve, value ids, value lens, value root)
{
  CAMLparam4(ve, ids, lens, root);
  CAMLlocal3(res, arr, nd);
  phylo_engine *e = Engine_val(ve);
  int n = (int)Bigarray_val(ids)->dim[0], i, rc = PHYLO_OK, max_id = 0;
  const int32_t *id = (const int32_t *)Data_bigarray_val(ids);
  const double *tl = (const double *)Data_bigarray_val(lens);
  phylo_op *ops = (phylo_op *)calloc(n > 0 ? n : 1, sizeof(phylo_op));
  uint64_t *gens = (uint64_t *)calloc(n > 0 ? n : 1, sizeof(uint64_t));
  int *map;
  double lnl = 0.0;
  int ra = Int_val(Field(root, 0)), rb = Int_val(Field(root, 1));
  double rt = Double_val(Field(root, 2));
  for (i = 0; i < 3 * n; ++i) if (id[i] > max_id) max_id = id[i];
  if (ra > max_id) max_id = ra;
  if (rb > max_id) max_id = rb;
  map = (int *)malloc(sizeof(int) * (size_t)(max_id + 1));
  for (i = 0; i <= max_id; ++i) map[i] = i; /* tips and already-resident slots map to themselves */
  for (i = 0; i < n; ++i) { /* allocation may run the GC: ids / lens are re-read afterwards */
    int slot = slot_new(ve, 0, &gens[i]);
    id = (const int32_t *)Data_bigarray_val(ids);
    map[id[3 * i]] = slot;
  }
  tl = (const double *)Data_bigarray_val(lens);
  for (i = 0; i < n; ++i) {
    ops[i].parent = map[id[3 * i]]; ops[i].left = map[id[3 * i + 1]]; ops[i].right = map[id[3 * i + 2]];
    ops[i].pad_ = 0; ops[i].t_left = tl[2 * i]; ops[i].t_right = tl[2 * i + 1];
  }
  ra = map[ra];
  rb = map[rb];
  caml_release_runtime_system();
  rc = phylo_lk_score_tree(e, ops, n, ra, rb, rt, &lnl);
  caml_acquire_runtime_system();
  if (rc != PHYLO_OK)
    for (i = 0; i < n; ++i) slot_drop(e, 0, ops[i].parent, gens[i]);
  if (rc == PHYLO_OK) {
    arr = caml_alloc_tuple(n);
    for (i = 0; i < n; ++i) {
      nd = node_wrap(ve, 0, ops[i].parent, 1, gens[i]);
      Store_field(arr, i, nd);
    }
  }
  free(ops);
  free(gens);
  free(map);
  check(e, rc);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(lnl));
  Store_field(res, 1, arr);
  CAMLreturn(res);
}

/* schedule of an already scored tree with the REAL slots: ids as in score_tree, interior id of op i = the
 * slot of down.(i) (score_tree's node array). Returns malloc'ed ops; *ra / *rb are mapped in place. */
static phylo_op *ops_of_scored_tree(value ids, value lens, value down, int *ra, int *rb, int *n_out)
{
  int n = (int)Bigarray_val(ids)->dim[0], i, max_id = 0;
  const int32_t *id = (const int32_t *)Data_bigarray_val(ids);
  const double *tl = (const double *)Data_bigarray_val(lens);
  phylo_op *ops = (phylo_op *)calloc(n > 0 ? n : 1, sizeof(phylo_op));
  int *map;
  for (i = 0; i < 3 * n; ++i) if (id[i] > max_id) max_id = id[i];
  if (*ra > max_id) max_id = *ra;
  if (*rb > max_id) max_id = *rb;
  map = (int *)malloc(sizeof(int) * (size_t)(max_id + 1));
  for (i = 0; i <= max_id; ++i) map[i] = i;
  for (i = 0; i < n; ++i) map[id[3 * i]] = Node_slot(Field(down, i));
  for (i = 0; i < n; ++i) {
    ops[i].parent = map[id[3 * i]]; ops[i].left = map[id[3 * i + 1]]; ops[i].right = map[id[3 * i + 2]];
    ops[i].pad_ = 0; ops[i].t_left = tl[2 * i]; ops[i].t_right = tl[2 * i + 1];
  }
  *ra = map[*ra];
  *rb = map[*rb];
  free(map);
  *n_out = n;
  return ops;
}

/* external uppass : engine -> ids -> lens -> (root_a * root_b * root_t) -> node array (score_tree's) -> node array
 * The second half of Node.Make3D (lib/node.ml:363-477): for op i the result holds at 2 i (2 i + 1) the value
 * of the REST OF THE TREE above its left (right) child -- the third direction of that child. Any branch
 * (child, its up value) can then be handed to edge_lnl / optimize_branch / edge_eval (readjust_3,
 * lib/node.ml:239-256). Fresh node values; the inputs are not touched. */
CAMLprim value likelihood_CAML_uppass(value ve, value ids, value lens, value root, value down)
{
  CAMLparam5(ve, ids, lens, root, down);
  CAMLlocal2(arr, nd);
  phylo_engine *e = Engine_val(ve);
  int n = (int)Bigarray_val(ids)->dim[0], i, rc, n_slots = 0, n_ops = 0;
  int ra = Int_val(Field(root, 0)), rb = Int_val(Field(root, 1));
  double rt = Double_val(Field(root, 2));
  int *ups = (int *)calloc(2 * (n > 0 ? n : 1), sizeof(int));
  uint64_t *gens = (uint64_t *)calloc(2 * (n > 0 ? n : 1), sizeof(uint64_t));
  int32_t *up_slot;
  phylo_op *ops;
  for (i = 0; i < 2 * n; ++i) ups[i] = slot_new(ve, 0, &gens[i]); /* may run the GC: Bigarrays are read afterwards */
  ops = ops_of_scored_tree(ids, lens, down, &ra, &rb, &n_ops);
  phylo_lk_shape(e, NULL, NULL, &n_slots);
  up_slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_slots > 0 ? n_slots : 1));
  for (i = 0; i < n_slots; ++i) up_slot[i] = -1;
  for (i = 0; i < n; ++i) { up_slot[ops[i].left] = ups[2 * i]; up_slot[ops[i].right] = ups[2 * i + 1]; }
  caml_release_runtime_system();
  rc = phylo_lk_uppass(e, ops, n, ra, rb, rt, up_slot);
  caml_acquire_runtime_system();
  if (rc != PHYLO_OK)
    for (i = 0; i < 2 * n; ++i) slot_drop(e, 0, ups[i], gens[i]);
  if (rc == PHYLO_OK) {
    arr = caml_alloc_tuple(2 * n);
    for (i = 0; i < 2 * n; ++i) {
      nd = node_wrap(ve, 0, ups[i], 1, gens[i]);
      Store_field(arr, i, nd);
    }
  }
  free(ops); free(ups); free(gens); free(up_slot);
  check(e, rc);
  CAMLreturn(arr);
}

/* external param_gradient : engine -> ids -> lens -> (root_a * root_b * root_t) -> (node array * node array)
 *                           -> (matrix option * matrix option * matrix option) -> vector -> unit
 * (down, up) = the node arrays of score_tree and uppass; the options are dQ (n_params x S*S), drates
 * (n_params x K), dpriors (n_params x S); out: n_params. What gen_subst_opt_func / gen_rates_opt_func /
 * gen_prior_opt_func (lib/mlModel.ml:822-829, `failwith "todo"`) need for a gradient-based optimiser. */
CAMLprim value likelihood_CAML_param_gradient(value ve, value ids, value lens, value root, value nodes, value dirs, value out)
{
  CAMLparam5(ve, ids, lens, root, nodes);
  CAMLxparam2(dirs, out);
  phylo_engine *e = Engine_val(ve);
  value down = Field(nodes, 0), up = Field(nodes, 1);
  int n = (int)Bigarray_val(ids)->dim[0], i, rc, n_slots = 0, n_ops = 0, n_params = (int)Bigarray_val(out)->dim[0];
  int ra = Int_val(Field(root, 0)), rb = Int_val(Field(root, 1));
  double rt = Double_val(Field(root, 2));
  const double *dq = (Field(dirs, 0) == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(Field(dirs, 0), 0));
  const double *dr = (Field(dirs, 1) == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(Field(dirs, 1), 0));
  const double *dp = (Field(dirs, 2) == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(Field(dirs, 2), 0));
  double *g = (double *)Data_bigarray_val(out);
  phylo_op *ops = ops_of_scored_tree(ids, lens, down, &ra, &rb, &n_ops);
  int32_t *up_slot;
  phylo_lk_shape(e, NULL, NULL, &n_slots);
  up_slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_slots > 0 ? n_slots : 1));
  for (i = 0; i < n_slots; ++i) up_slot[i] = -1;
  for (i = 0; i < n; ++i) { up_slot[ops[i].left] = Node_slot(Field(up, 2 * i)); up_slot[ops[i].right] = Node_slot(Field(up, 2 * i + 1)); }
  caml_release_runtime_system();
  rc = phylo_lk_param_gradient(e, ops, n, ra, rb, rt, up_slot, n_params, dq, dr, dp, NULL, g);
  caml_acquire_runtime_system();
  free(ops); free(up_slot);
  check(e, rc);
  CAMLreturn(Val_unit);
}
CAMLprim value likelihood_CAML_param_gradient_bc(value *argv, int argn)
{
  (void)argn;
  return likelihood_CAML_param_gradient(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6]);
}

/* external edge_lnl : engine -> node -> node -> vector (lengths) -> vector (out lnL) -> unit
 * Likelihood_c.root_cost / distance_1 (lib/nodeData.ml:29,32) for a batch of lengths */
CAMLprim value likelihood_CAML_edge_lnl(value ve, value va, value vb, value ts, value out)
{
  CAMLparam5(ve, va, vb, ts, out);
  phylo_engine *e = Engine_val(ve);
  int rc, n = (int)Bigarray_val(ts)->dim[0], a = Node_slot(va), b = Node_slot(vb);
  const double *t = (const double *)Data_bigarray_val(ts);
  double *o = (double *)Data_bigarray_val(out);
  caml_release_runtime_system();
  rc = phylo_lk_edge_lnl(e, a, b, t, n, o);
  caml_acquire_runtime_system();
  check(e, rc);
  CAMLreturn(Val_unit);
}

/* external optimize_branch : engine -> node -> node -> (float * float * float * float) -> float * float
 * (t0, t_min, t_max, tol) -> (t_opt, lnL): Likelihood_c.adjust_3 (lib/nodeData.ml:25; TODO in
 * lib/likelihood_c.ml:19) -- maximum-likelihood length of the edge between two directed CLVs */
CAMLprim value likelihood_CAML_optimize_branch(value ve, value va, value vb, value par)
{
  CAMLparam4(ve, va, vb, par);
  CAMLlocal1(res);
  phylo_engine *e = Engine_val(ve);
  int rc, iters = 0, a = Node_slot(va), b = Node_slot(vb);
  double t0 = Double_val(Field(par, 0)), tmin = Double_val(Field(par, 1)), tmax = Double_val(Field(par, 2));
  double tol = Double_val(Field(par, 3)), t = 0.0, lnl = 0.0;
  caml_release_runtime_system();
  rc = phylo_lk_optimize_branch(e, a, b, t0, tmin, tmax, tol, 50, &t, &lnl, &iters);
  caml_acquire_runtime_system();
  check(e, rc);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(t));
  Store_field(res, 1, caml_copy_double(lnl));
  CAMLreturn(res);
}

/* external edge_eval : engine -> node -> node -> vector (lengths) -> matrix (3 x n: lnL, d1, d2) -> unit
 * sum table of the edge, then lnL and its derivatives for a batch of lengths */
CAMLprim value likelihood_CAML_edge_eval(value ve, value va, value vb, value ts, value out)
{
  CAMLparam5(ve, va, vb, ts, out);
  phylo_engine *e = Engine_val(ve);
  int rc, n = (int)Bigarray_val(ts)->dim[0], a = Node_slot(va), b = Node_slot(vb);
  const double *t = (const double *)Data_bigarray_val(ts);
  double *o = (double *)Data_bigarray_val(out);
  caml_release_runtime_system();
  rc = phylo_lk_edge_prepare(e, a, b);
  if (rc == PHYLO_OK) rc = phylo_lk_edge_eval(e, t, n, o, o + n, o + 2 * n);
  caml_acquire_runtime_system();
  check(e, rc);
  CAMLreturn(Val_unit);
}

/* external get_clv : engine -> node -> (float, float64_elt, c_layout) Array3.t -> unit */
CAMLprim value likelihood_CAML_get_clv(value ve, value vnode, value out)
{
  CAMLparam3(ve, vnode, out);
  phylo_engine *e = Engine_val(ve);
  check(e, phylo_lk_get_clv(e, Node_slot(vnode), (double *)Data_bigarray_val(out), NULL));
  CAMLreturn(Val_unit);
}

/* ------------------------------------ NonAdditive_c / Bitvector (lib/nonAdditive_c.ml) ---- */
/* external set_tips : engine -> (int, int8_unsigned_elt, c_layout) Array2.t -> n_states:int
 *                     -> weights:vector option -> capacity:int -> unit */
CAMLprim value nonadd_CAML_set_tips(value ve, value codes, value vns, value wo, value vcap)
{
  CAMLparam5(ve, codes, vns, wo, vcap);
  phylo_engine *e = Engine_val(ve);
  struct caml_ba_array *b = Bigarray_val(codes);
  int kind = (int)(b->flags & 0xff), bytes = 1, rc;
  const double *w = (wo == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(wo, 0));
  if (kind == CAML_BA_INT32) bytes = 4;
  else if (kind == CAML_BA_INT64) bytes = 8;
  else if (kind != CAML_BA_UINT8) caml_failwith("nonadd set_tips: codes must be uint8, int32 or int64");
  caml_release_runtime_system();
  rc = phylo_fitch_set_tips(e, (int)b->dim[0], (int64_t)b->dim[1], bytes, Int_val(vns), b->data, w,
                            Int_val(vcap));
  caml_acquire_runtime_system();
  check(e, rc);
  CAMLreturn(Val_unit);
}

/* external tip : engine -> int -> node = "nonadd_CAML_tip" */
CAMLprim value nonadd_CAML_tip(value ve, value vi)
{
  CAMLparam2(ve, vi);
  CAMLreturn(node_wrap(ve, 1, Int_val(vi), 0, 0));
}

/* external median_2 : engine -> node -> node -> node * int   (* new node value, node-local cost *)
 * NonAdditive_c.median_2 (lib/nonAdditive_c.ml:19-35) == bv_CAML_fitch_median2 (bv.c:463-480),
 * which likewise returns (fresh vect, cost) */
CAMLprim value nonadd_CAML_median2(value ve, value vl, value vr)
{
  CAMLparam3(ve, vl, vr);
  CAMLlocal2(res, nd);
  phylo_engine *e = Engine_val(ve);
  uint64_t cost = 0, gen = 0;
  int slot = slot_new(ve, 1, &gen), rc;
  rc = phylo_fitch_median_2(e, slot, Node_slot(vl), Node_slot(vr), &cost);
  if (rc != PHYLO_OK) slot_drop(e, 1, slot, gen);
  check(e, rc);
  nd = node_wrap(ve, 1, slot, 1, gen);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, nd);
  Store_field(res, 1, Val_long((intptr_t)cost));
  CAMLreturn(res);
}

/* external median_3 : engine -> node -> node -> node -> node -> node
 * (own preliminary sets, parent's final sets, the two children's preliminary sets) -> final sets;
 * the reference's commented-out bv_CAML_fitch_median3(vb0, vb1, vb2, vb3) (lib/bitvector/bv.h:94) */
CAMLprim value nonadd_CAML_median3(value ve, value vprelim, value vparent, value vl, value vr)
{
  CAMLparam5(ve, vprelim, vparent, vl, vr);
  phylo_engine *e = Engine_val(ve);
  uint64_t gen = 0;
  int slot = slot_new(ve, 1, &gen), rc;
  rc = phylo_fitch_median_3(e, slot, Node_slot(vprelim), Node_slot(vparent), Node_slot(vl), Node_slot(vr));
  if (rc != PHYLO_OK) slot_drop(e, 1, slot, gen);
  check(e, rc);
  CAMLreturn(node_wrap(ve, 1, slot, 1, gen));
}

/* external distance : engine -> node -> node -> int   (bv_CAML_distance2, bv.c:455-461) */
CAMLprim value nonadd_CAML_distance(value ve, value va, value vb)
{
  CAMLparam3(ve, va, vb);
  phylo_engine *e = Engine_val(ve);
  uint64_t d = 0;
  check(e, phylo_fitch_distance(e, Node_slot(va), Node_slot(vb), &d));
  CAMLreturn(Val_long((intptr_t)d));
}

static phylo_op *ops_of_ids(value ids, int *n_out)
{
  int n = (int)Bigarray_val(ids)->dim[0], i;
  const int32_t *id = (const int32_t *)Data_bigarray_val(ids);
  phylo_op *ops = (phylo_op *)calloc(n > 0 ? n : 1, sizeof(phylo_op));
  for (i = 0; i < n; ++i) {
    ops[i].parent = id[3 * i]; ops[i].left = id[3 * i + 1]; ops[i].right = id[3 * i + 2];
  }
  *n_out = n;
  return ops;
}

/* external score_tree : engine -> ids (n_ops x 3 int32, numbering as in likelihood score_tree)
 *                       -> root_a:int -> root_b:int -> int * node array   (tree length, node values) */
CAMLprim value nonadd_CAML_score_tree(value ve, value ids, value va, value vb)
{
  CAMLparam4(ve, ids, va, vb);
  CAMLlocal3(res, arr, nd);
  phylo_engine *e = Engine_val(ve);
  int n = (int)Bigarray_val(ids)->dim[0], rc, a = Int_val(va), b = Int_val(vb), i, max_id = a > b ? a : b;
  const int32_t *id = (const int32_t *)Data_bigarray_val(ids);
  phylo_op *ops;
  uint64_t *gens = (uint64_t *)calloc(n > 0 ? n : 1, sizeof(uint64_t));
  int *map;
  uint64_t len = 0;
  for (i = 0; i < 3 * n; ++i) if (id[i] > max_id) max_id = id[i];
  map = (int *)malloc(sizeof(int) * (size_t)(max_id + 1));
  for (i = 0; i <= max_id; ++i) map[i] = i;
  for (i = 0; i < n; ++i) {
    int slot = slot_new(ve, 1, &gens[i]);
    id = (const int32_t *)Data_bigarray_val(ids);
    map[id[3 * i]] = slot;
  }
  ops = ops_of_ids(ids, &n);
  for (i = 0; i < n; ++i) { ops[i].parent = map[ops[i].parent]; ops[i].left = map[ops[i].left]; ops[i].right = map[ops[i].right]; }
  a = map[a];
  b = map[b];
  caml_release_runtime_system();
  rc = phylo_fitch_score_tree(e, ops, n, a, b, &len);
  caml_acquire_runtime_system();
  if (rc != PHYLO_OK)
    for (i = 0; i < n; ++i) slot_drop(e, 1, ops[i].parent, gens[i]);
  if (rc == PHYLO_OK) {
    arr = caml_alloc_tuple(n);
    for (i = 0; i < n; ++i) {
      nd = node_wrap(ve, 1, ops[i].parent, 1, gens[i]);
      Store_field(arr, i, nd);
    }
  }
  free(ops);
  free(gens);
  free(map);
  check(e, rc);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, Val_long((intptr_t)len));
  Store_field(res, 1, arr);
  CAMLreturn(res);
}

/* external uppass : engine -> ids (slots as held by the node values: phylo_CAML_node_slot) -> root_a:int
 *                   -> root_b:int -> unit
 * fills Node.final_states (lib/node.ml:260-268, TODO in the reference) for a tree scored by score_tree */
CAMLprim value nonadd_CAML_uppass(value ve, value ids, value va, value vb)
{
  CAMLparam4(ve, ids, va, vb);
  phylo_engine *e = Engine_val(ve);
  int n, rc, a = Int_val(va), b = Int_val(vb);
  phylo_op *ops = ops_of_ids(ids, &n);
  caml_release_runtime_system();
  rc = phylo_fitch_uppass(e, ops, n, a, b);
  caml_acquire_runtime_system();
  free(ops);
  check(e, rc);
  CAMLreturn(Val_unit);
}

/* external get_states : engine -> node -> final:bool -> (int, int8_unsigned_elt, c_layout) Array1.t -> unit */
CAMLprim value nonadd_CAML_get_states(value ve, value vnode, value vfinal, value out)
{
  CAMLparam4(ve, vnode, vfinal, out);
  phylo_engine *e = Engine_val(ve);
  check(e, phylo_fitch_get_states(e, Node_slot(vnode), Int_val(vfinal) != 0, Data_bigarray_val(out)));
  CAMLreturn(Val_unit);
}

/* Bitvector set algebra over node values: bv_CAML_union / inter / popcount / eltcount / saturation /
 * poly_saturation / compare (lib/bitvector/bv.c:405-453). union / inter return NEW node values, like
 * the reference's (bv_copy + custom block). */
static value bv_binop_stub(value ve, value va, value vb, int is_union)
{
  CAMLparam3(ve, va, vb);
  phylo_engine *e = Engine_val(ve);
  uint64_t gen = 0;
  int slot = slot_new(ve, 1, &gen), rc;
  rc = is_union ? phylo_bv_union(e, slot, Node_slot(va), Node_slot(vb)) : phylo_bv_inter(e, slot, Node_slot(va), Node_slot(vb));
  if (rc != PHYLO_OK) slot_drop(e, 1, slot, gen);
  check(e, rc);
  CAMLreturn(node_wrap(ve, 1, slot, 1, gen));
}
CAMLprim value nonadd_CAML_union(value ve, value va, value vb) { return bv_binop_stub(ve, va, vb, 1); }
CAMLprim value nonadd_CAML_inter(value ve, value va, value vb) { return bv_binop_stub(ve, va, vb, 0); }
CAMLprim value nonadd_CAML_popcount(value ve, value va)
{
  CAMLparam2(ve, va);
  phylo_engine *e = Engine_val(ve);
  uint64_t n = 0;
  check(e, phylo_bv_popcount(e, Node_slot(va), &n));
  CAMLreturn(Val_long((intptr_t)n)); /* the reference returns Int_val(i) here by mistake (bv.c:410) */
}
CAMLprim value nonadd_CAML_eltcount(value ve, value va, value vi)
{
  CAMLparam3(ve, va, vi);
  phylo_engine *e = Engine_val(ve);
  int n = 0;
  check(e, phylo_bv_eltcount(e, Node_slot(va), (int64_t)Long_val(vi), &n));
  CAMLreturn(Val_int(n));
}
CAMLprim value nonadd_CAML_saturation(value ve, value va, value vmask)
{
  CAMLparam3(ve, va, vmask);
  phylo_engine *e = Engine_val(ve);
  uint64_t n = 0;
  check(e, phylo_bv_saturation(e, Node_slot(va), (uint64_t)Long_val(vmask), &n));
  CAMLreturn(Val_long((intptr_t)n));
}
CAMLprim value nonadd_CAML_poly_saturation(value ve, value va, value vn)
{
  CAMLparam3(ve, va, vn);
  phylo_engine *e = Engine_val(ve);
  uint64_t n = 0;
  check(e, phylo_bv_poly_saturation(e, Node_slot(va), Int_val(vn), &n));
  CAMLreturn(Val_long((intptr_t)n));
}
CAMLprim value nonadd_CAML_compare(value ve, value va, value vb)
{
  CAMLparam3(ve, va, vb);
  phylo_engine *e = Engine_val(ve);
  int r = 0;
  check(e, phylo_bv_compare(e, Node_slot(va), Node_slot(vb), &r));
  CAMLreturn(Val_int(r));
}

/* ------------------------------------- several GPUs behind one handle (phylo_group_*) ---- */
/* The OCaml runtime is one process, so the whole-tree entry points also exist over a group
 * handle that owns one engine per GPU (include/phylo_engine.h, last section). Same argument
 * shapes as the single-engine stubs above; `group` is a second custom block. */
#define Group_val(v) (*((phylo_group **)Data_custom_val(v)))

static void group_finalize(value v)
{
  if (Group_val(v)) phylo_group_destroy(Group_val(v));
  Group_val(v) = NULL;
}

static struct custom_operations group_ops = {
    "AMNH/phylo_b200/group/0.1", group_finalize, custom_compare_default, custom_hash_default,
    custom_serialize_default, custom_deserialize_default, custom_compare_ext_default};

static void gcheck(phylo_group *g, int rc)
{
  if (rc != PHYLO_OK) caml_failwith(phylo_group_last_error(g));
}

/* external group_create : int array -> group = "phylo_CAML_group_create"   (device ids) */
CAMLprim value phylo_CAML_group_create(value vdevs)
{
  CAMLparam1(vdevs);
  CAMLlocal1(res);
  int n = (int)Wosize_val(vdevs), i, rc;
  int devs[64];
  phylo_group *g = NULL;
  if (n < 1 || n > 64) caml_failwith("group_create: need 1..64 device ids");
  for (i = 0; i < n; ++i) devs[i] = Int_val(Field(vdevs, i));
  rc = phylo_group_create(devs, n, &g);
  if (rc != PHYLO_OK) caml_failwith(phylo_group_last_error(NULL));
  res = caml_alloc_custom(&group_ops, sizeof(phylo_group *), 0, 1);
  Group_val(res) = g;
  CAMLreturn(res);
}

/* external group_set_model : group -> matrix -> matrix -> matrix option -> (vector * vector * vector * float option) -> unit */
CAMLprim value likelihood_CAML_group_set_model(value vg, value U, value D, value Uio, value rest)
{
  CAMLparam5(vg, U, D, Uio, rest);
  phylo_group *g = Group_val(vg);
  value pri = Field(rest, 0), rates = Field(rest, 1), probs = Field(rest, 2), pinv = Field(rest, 3);
  const double *ui = (Uio == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(Uio, 0));
  double pv = (pinv == Val_int(0)) ? -1.0 : Double_val(Field(pinv, 0));
  gcheck(g, phylo_group_lk_set_model(g, (int)Bigarray_val(U)->dim[0], (int)Bigarray_val(rates)->dim[0],
                                     (double *)Data_bigarray_val(U), (double *)Data_bigarray_val(D), ui,
                                     (double *)Data_bigarray_val(pri), (double *)Data_bigarray_val(rates),
                                     (double *)Data_bigarray_val(probs), pv));
  CAMLreturn(Val_unit);
}

/* external group_set_tips : group -> masks -> vector option -> int -> unit */
CAMLprim value likelihood_CAML_group_set_tips(value vg, value masks, value wo, value vcap)
{
  CAMLparam4(vg, masks, wo, vcap);
  phylo_group *g = Group_val(vg);
  struct caml_ba_array *b = Bigarray_val(masks);
  int kind = (int)(b->flags & 0xff), bytes = 1, rc;
  const double *w = (wo == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(wo, 0));
  if (kind == CAML_BA_INT32) bytes = 4;
  else if (kind == CAML_BA_INT64) bytes = 8;
  else if (kind != CAML_BA_UINT8) caml_failwith("likelihood group_set_tips: masks must be uint8, int32 or int64");
  caml_release_runtime_system();
  rc = phylo_group_lk_set_tips(g, (int)b->dim[0], (int64_t)b->dim[1], b->data, bytes, w, Int_val(vcap));
  caml_acquire_runtime_system();
  gcheck(g, rc);
  CAMLreturn(Val_unit);
}

/* external group_score_tree : group -> ids -> lens -> (int * int * float) -> float */
CAMLprim value likelihood_CAML_group_score_tree(value vg, value ids, value lens, value root)
{
  CAMLparam4(vg, ids, lens, root);
  phylo_group *g = Group_val(vg);
  int n = (int)Bigarray_val(ids)->dim[0], i, rc;
  const int32_t *id = (const int32_t *)Data_bigarray_val(ids);
  const double *tl = (const double *)Data_bigarray_val(lens);
  phylo_op *ops = (phylo_op *)malloc(sizeof(phylo_op) * (n > 0 ? n : 1));
  double lnl = 0.0;
  int ra = Int_val(Field(root, 0)), rb = Int_val(Field(root, 1));
  double rt = Double_val(Field(root, 2));
  for (i = 0; i < n; ++i) {
    ops[i].parent = id[3 * i]; ops[i].left = id[3 * i + 1]; ops[i].right = id[3 * i + 2];
    ops[i].pad_ = 0; ops[i].t_left = tl[2 * i]; ops[i].t_right = tl[2 * i + 1];
  }
  caml_release_runtime_system();
  rc = phylo_group_lk_score_tree(g, ops, n, ra, rb, rt, &lnl);
  caml_acquire_runtime_system();
  free(ops);
  gcheck(g, rc);
  CAMLreturn(caml_copy_double(lnl));
}

/* external group_optimize_branch : group -> int -> int -> (float * float * float * float) -> float * float */
CAMLprim value likelihood_CAML_group_optimize_branch(value vg, value va, value vb, value par)
{
  CAMLparam4(vg, va, vb, par);
  CAMLlocal1(res);
  phylo_group *g = Group_val(vg);
  int rc, a = Int_val(va), b = Int_val(vb), iters = 0;
  double t0 = Double_val(Field(par, 0)), tmin = Double_val(Field(par, 1)), tmax = Double_val(Field(par, 2)),
         tol = Double_val(Field(par, 3)), t = 0.0, l = 0.0;
  caml_release_runtime_system();
  rc = phylo_group_lk_optimize_branch(g, a, b, t0, tmin, tmax, tol, 50, &t, &l, &iters);
  caml_acquire_runtime_system();
  gcheck(g, rc);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(t));
  Store_field(res, 1, caml_copy_double(l));
  CAMLreturn(res);
}

/* external group_set_tips : group -> codes -> n_states:int -> vector option -> capacity:int -> unit */
CAMLprim value nonadd_CAML_group_set_tips(value vg, value codes, value vns, value wo, value vcap)
{
  CAMLparam5(vg, codes, vns, wo, vcap);
  phylo_group *g = Group_val(vg);
  struct caml_ba_array *b = Bigarray_val(codes);
  int kind = (int)(b->flags & 0xff), bytes = 1, rc;
  const double *w = (wo == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(wo, 0));
  if (kind == CAML_BA_INT32) bytes = 4;
  else if (kind == CAML_BA_INT64) bytes = 8;
  else if (kind != CAML_BA_UINT8) caml_failwith("nonadd group_set_tips: codes must be uint8, int32 or int64");
  caml_release_runtime_system();
  rc = phylo_group_fitch_set_tips(g, (int)b->dim[0], (int64_t)b->dim[1], bytes, Int_val(vns), b->data, w,
                                  Int_val(vcap));
  caml_acquire_runtime_system();
  gcheck(g, rc);
  CAMLreturn(Val_unit);
}

/* external group_score_tree : group -> ids -> root_a:int -> root_b:int -> int   (* tree length *) */
CAMLprim value nonadd_CAML_group_score_tree(value vg, value ids, value va, value vb)
{
  CAMLparam4(vg, ids, va, vb);
  phylo_group *g = Group_val(vg);
  int n, rc, a = Int_val(va), b = Int_val(vb);
  phylo_op *ops = ops_of_ids(ids, &n);
  uint64_t len = 0;
  caml_release_runtime_system();
  rc = phylo_group_fitch_score_tree(g, ops, n, a, b, &len);
  caml_acquire_runtime_system();
  free(ops);
  gcheck(g, rc);
  CAMLreturn(Val_long((intptr_t)len));
}

/* external set_symbol_table : engine -> (int64, int64_elt, c_layout) Array1.t option -> unit
 * 256 state masks indexed by symbol byte, read off Alphabet.t's name_code map for the
 * single-character names (lib/alphabet.ml:180-199); None restores state-mask input. */
CAMLprim value phylo_CAML_set_symbol_table(value ve, value vtab)
{
  CAMLparam2(ve, vtab);
  phylo_engine *e = Engine_val(ve);
  const uint64_t *t = NULL;
  if (vtab != Val_int(0)) {
    if (Bigarray_val(Field(vtab, 0))->dim[0] != 256) caml_failwith("set_symbol_table: need 256 entries");
    t = (const uint64_t *)Data_bigarray_val(Field(vtab, 0));
  }
  check(e, phylo_engine_set_symbol_table(e, t));
  CAMLreturn(Val_unit);
}

/* external gamma_rates : float -> int -> int -> vector   (alpha, classes, mode) -> rates
 * replaces MlModel.gamma_rates (lib/mlModel.ml:93-99; Pareto/GSL there). mode 0 = the reference's
 * literal quantiles, 1 = Yang (1994) class means. */
CAMLprim value likelihood_CAML_gamma_rates(value valpha, value vk, value vmode)
{
  CAMLparam3(valpha, vk, vmode);
  CAMLlocal1(res);
  intptr_t dims[1];
  dims[0] = Int_val(vk);
  if (dims[0] < 1) caml_failwith("gamma_rates: need at least one class");
  res = caml_ba_alloc(CAML_BA_FLOAT64 | CAML_BA_C_LAYOUT, 1, NULL, dims);
  if (phylo_gamma_rates(Double_val(valpha), Int_val(vk), Int_val(vmode), (double *)Data_bigarray_val(res), NULL) != PHYLO_OK)
    caml_failwith("gamma_rates: alpha must be positive and finite");
  CAMLreturn(res);
}

/* external integerize : matrix -> vector option -> int -> (int32, int32_elt, c_layout) Array2.t
 * MlModel.integerized_model's conversion (lib/mlModel.ml:639-660) of P(t) into integer costs. */
CAMLprim value likelihood_CAML_integerize(value P, value prio, value vsigma)
{
  CAMLparam3(P, prio, vsigma);
  CAMLlocal1(res);
  intptr_t dims[2];
  int n = (int)Bigarray_val(P)->dim[0];
  const double *pri = (prio == Val_int(0)) ? NULL : (const double *)Data_bigarray_val(Field(prio, 0));
  dims[0] = n;
  dims[1] = n;
  res = caml_ba_alloc(CAML_BA_INT32 | CAML_BA_C_LAYOUT, 2, NULL, dims);
  if (phylo_integerize_matrix((const double *)Data_bigarray_val(P), pri, n, Int_val(vsigma),
                              (int32_t *)Data_bigarray_val(res)) != PHYLO_OK)
    caml_failwith("integerize: an entry of P is not positive");
  CAMLreturn(res);
}
