/* TEST INFRASTRUCTURE ONLY -- stand-in for <caml/custom.h>. */
#ifndef SHIM_CAML_CUSTOM_H
#define SHIM_CAML_CUSTOM_H
#include "mlvalues.h"
struct custom_operations {
  const char *identifier;
  void (*finalize)(value v);
  int (*compare)(value v1, value v2);
  long (*hash)(value v);
  void (*serialize)(value v, unsigned long *wsize_32, unsigned long *wsize_64);
  unsigned long (*deserialize)(void *dst);
  int (*compare_ext)(value v1, value v2);
};
#define custom_finalize_default    NULL
#define custom_compare_default     NULL
#define custom_hash_default        NULL
#define custom_serialize_default   NULL
#define custom_deserialize_default NULL
#define custom_compare_ext_default NULL
#define Data_custom_val(v) ((void *)&Field((v), 1))
value caml_alloc_custom(struct custom_operations *ops, unsigned long size,
                        mlsize_t mem, mlsize_t max);
void caml_register_custom_operations(struct custom_operations *ops);
#define register_custom_operations caml_register_custom_operations
#endif
