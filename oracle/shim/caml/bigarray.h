/* TEST INFRASTRUCTURE ONLY -- stand-in for <caml/bigarray.h>. */
#ifndef SHIM_CAML_BIGARRAY_H
#define SHIM_CAML_BIGARRAY_H
#include "mlvalues.h"
struct caml_ba_array {
  void *data;
  intptr_t num_dims;
  intptr_t flags;
  void *proxy;
  intptr_t dim[4];
};
#define Caml_ba_array_val(v) ((struct caml_ba_array *)&Field((v), 1))
#define Caml_ba_data_val(v)  (Caml_ba_array_val(v)->data)
#define Bigarray_val(v)      Caml_ba_array_val(v)
#define Data_bigarray_val(v) Caml_ba_data_val(v)
enum { CAML_BA_FLOAT32 = 0, CAML_BA_FLOAT64 = 1, CAML_BA_UINT8 = 3, CAML_BA_INT32 = 6, CAML_BA_INT64 = 7 };
#define CAML_BA_C_LAYOUT 0
#define CAML_BA_MANAGED  0x200
#define BIGARRAY_FLOAT64  CAML_BA_FLOAT64
#define BIGARRAY_C_LAYOUT CAML_BA_C_LAYOUT
#define BIGARRAY_MANAGED  CAML_BA_MANAGED
value caml_ba_alloc(int flags, int num_dims, void *data, intptr_t *dim);
#define alloc_bigarray(f, n, d, dims) caml_ba_alloc((f), (n), (d), (intptr_t *)(dims))
#endif
