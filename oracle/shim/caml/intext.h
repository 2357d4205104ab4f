/* TEST INFRASTRUCTURE ONLY -- stand-in for <caml/intext.h>. */
#ifndef SHIM_CAML_INTEXT_H
#define SHIM_CAML_INTEXT_H
#include <stdint.h>
void caml_serialize_int_1(int i);
void caml_serialize_int_2(int i);
void caml_serialize_int_4(int32_t i);
void caml_serialize_int_8(int64_t i);
int caml_deserialize_uint_1(void);
int caml_deserialize_uint_2(void);
uint32_t caml_deserialize_uint_4(void);
uint64_t caml_deserialize_uint_8(void);
#endif
