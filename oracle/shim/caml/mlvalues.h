/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for OCaml's <caml/mlvalues.h>.
 * Lets the reference's C sources (lib/mlmodel.c, lib/bitvector/bv.c, lib/phyloc.c)
 * compile unmodified without an OCaml toolchain so that their *non-stub* C
 * functions (compose_gtr, bv_fitch, ...) can be called as the parity oracle.
 * The `value` stubs compile but are never called through this shim. */
#ifndef SHIM_CAML_MLVALUES_H
#define SHIM_CAML_MLVALUES_H
#include <stdint.h>
#include <stddef.h>
typedef intptr_t value;
typedef uintptr_t mlsize_t;
#define Val_long(x)   ((value)(((intptr_t)(x) << 1) + 1))
#define Long_val(x)   ((intptr_t)(x) >> 1)
#define Val_int(x)    Val_long(x)
#define Int_val(x)    ((int)Long_val(x))
#define Unsigned_long_val(x) ((uintptr_t)(x) >> 1)
#define Unsigned_int_val(x)  ((unsigned int)Unsigned_long_val(x))
#define Val_unit      Val_int(0)
#define Val_emptylist Val_int(0)
#define Val_bool(x)   Val_int((x) != 0)
#define Field(v, i)   (((value *)(v))[i])
#define Store_field(b, i, v) (Field(b, i) = (v))
#define Wosize_val(v) ((mlsize_t)(((uintptr_t *)(v))[-1] >> 10))
#define Double_val(v) (*(double *)(v))
#define Int32_val(v)  (*(int32_t *)((value *)(v) + 1))
#define Int64_val(v)  (*(int64_t *)((value *)(v) + 1))
#define CAMLprim
#define CAMLextern extern
#endif
