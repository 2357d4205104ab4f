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

value caml_alloc_custom(struct custom_operations *ops, unsigned long size, mlsize_t mem, mlsize_t max)
{
  (void)mem; (void)max;
  value v = shim_block(1 + (size + sizeof(value) - 1) / sizeof(value), 255);
  Field(v, 0) = (value)ops;
  return v;
}
void caml_register_custom_operations(struct custom_operations *ops) { (void)ops; }

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
  b->data = data; b->num_dims = num_dims; b->flags = flags;
  for (int i = 0; i < num_dims && i < 4; ++i) b->dim[i] = dim[i];
  return v;
}
