/* TEST INFRASTRUCTURE ONLY -- stand-in for <caml/alloc.h>. */
#ifndef SHIM_CAML_ALLOC_H
#define SHIM_CAML_ALLOC_H
#include "mlvalues.h"
value caml_alloc(mlsize_t wosize, int tag);
value caml_copy_int32(int32_t i);
value caml_copy_int64(int64_t i);
value caml_copy_double(double d);
value caml_alloc_tuple(mlsize_t n);
#endif
