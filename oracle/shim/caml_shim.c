/* TEST INFRASTRUCTURE ONLY -- runtime half of the fake OCaml headers in shim/caml/.
 * Just enough for the reference's `value` stubs to link; the parity oracle calls the
 * plain C functions underneath them (compose_gtr, bv_fitch, ...) and never these. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "caml/mlvalues.h"
#include "caml/alloc.h"
#include "caml/fail.h"
#include "caml/custom.h"
#include "caml/intext.h"
#include "caml/bigarray.h"

static char shim_failure[256];
static int shim_failed = 0;

void caml_failwith(const char *msg)
{
  /* the reference calls failwith() from LAPACK error paths; record, do not abort */
  strncpy(shim_failure, msg, sizeof(shim_failure) - 1);
  shim_failed = 1;
}
const char *shim_last_failure(void) { return shim_failed ? shim_failure : NULL; }
void shim_clear_failure(void) { shim_failed = 0; }

static value shim_block(mlsize_t wosize, int tag)
{
  uintptr_t *p = (uintptr_t *)calloc(wosize + 1, sizeof(uintptr_t));
  p[0] = ((uintptr_t)wosize << 10) | (uintptr_t)(tag & 0xff);
  return (value)(p + 1);
}
value caml_alloc(mlsize_t wosize, int tag) { return shim_block(wosize, tag); }
value caml_copy_int32(int32_t i) { value v = shim_block(2, 255); Int32_val(v) = i; return v; }
value caml_copy_int64(int64_t i) { value v = shim_block(2, 255); Int64_val(v) = i; return v; }
value caml_alloc_tuple(mlsize_t n) { return shim_block(n, 0); }
value caml_copy_double(double d) { value v = shim_block(1, 253); Double_val(v) = d; return v; }

/* A toy collector for the C-level lifetime test of stubs/phylo_stubs.c (tests/c/stub_lifetime.c):
 * custom blocks are registered here; the test says which values have become unreachable
 * (shim_mark_dead) and caml_gc_full_major -- which the stubs call when an engine's slot table is
 * full -- finalizes exactly those, as the real major GC would. */
struct shim_custom { value v; int dead; struct shim_custom *next; };
static struct shim_custom *shim_customs = NULL;
static long shim_n_live = 0, shim_n_finalized = 0, shim_n_gc = 0;

value caml_alloc_custom(struct custom_operations *ops, unsigned long size, mlsize_t mem, mlsize_t max)
{
  (void)mem; (void)max;
  value v = shim_block(1 + (size + sizeof(value) - 1) / sizeof(value), 255);
  struct shim_custom *c = (struct shim_custom *)malloc(sizeof(*c));
  Field(v, 0) = (value)ops;
  c->v = v; c->dead = 0; c->next = shim_customs;
  shim_customs = c;
  ++shim_n_live;
  return v;
}
void caml_register_custom_operations(struct custom_operations *ops) { (void)ops; }

void shim_mark_dead(value v)
{
  for (struct shim_custom *c = shim_customs; c; c = c->next)
    if (c->v == v) { c->dead = 1; return; }
}
value caml_gc_full_major(value unit)
{
  struct shim_custom **pp = &shim_customs;
  ++shim_n_gc;
  while (*pp) {
    struct shim_custom *c = *pp;
    if (c->dead) {
      struct custom_operations *ops = (struct custom_operations *)Field(c->v, 0);
      if (ops && ops->finalize) ops->finalize(c->v);
      free((uintptr_t *)c->v - 1);
      *pp = c->next;
      free(c);
      --shim_n_live;
      ++shim_n_finalized;
    } else {
      pp = &c->next;
    }
  }
  return unit;
}
long shim_live_custom_blocks(void) { return shim_n_live; }
long shim_finalized_custom_blocks(void) { return shim_n_finalized; }
long shim_gc_runs(void) { return shim_n_gc; }

void caml_serialize_int_1(int i) { (void)i; }
void caml_serialize_int_2(int i) { (void)i; }
void caml_serialize_int_4(int32_t i) { (void)i; }
void caml_serialize_int_8(int64_t i) { (void)i; }
int caml_deserialize_uint_1(void) { return 0; }
int caml_deserialize_uint_2(void) { return 0; }
uint32_t caml_deserialize_uint_4(void) { return 0; }
uint64_t caml_deserialize_uint_8(void) { return 0; }

value caml_ba_alloc(int flags, int num_dims, void *data, intptr_t *dim)
{
  value v = shim_block(1 + sizeof(struct caml_ba_array) / sizeof(value), 255);
  struct caml_ba_array *b = Caml_ba_array_val(v);
  b->num_dims = num_dims; b->flags = flags;
  size_t n = 1;
  for (int i = 0; i < num_dims && i < 4; ++i) { b->dim[i] = dim[i]; n *= (size_t)dim[i]; }
  b->data = data ? data : calloc(n ? n : 1, 8); /* data == NULL: runtime-owned storage (<= 8-byte elements) */
  return v;
}
