/* TEST INFRASTRUCTURE ONLY -- stand-in for <caml/fail.h>. */
#ifndef SHIM_CAML_FAIL_H
#define SHIM_CAML_FAIL_H
void caml_failwith(const char *msg);
#define failwith caml_failwith
#endif
