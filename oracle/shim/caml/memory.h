/* TEST INFRASTRUCTURE ONLY -- stand-in for <caml/memory.h>; GC roots are no-ops. */
#ifndef SHIM_CAML_MEMORY_H
#define SHIM_CAML_MEMORY_H
#include "mlvalues.h"
#define CAMLparam0()
#define CAMLparam1(a)             (void)(a)
#define CAMLparam2(a,b)           (void)(a);(void)(b)
#define CAMLparam3(a,b,c)         (void)(a);(void)(b);(void)(c)
#define CAMLparam4(a,b,c,d)       (void)(a);(void)(b);(void)(c);(void)(d)
#define CAMLparam5(a,b,c,d,e)     (void)(a);(void)(b);(void)(c);(void)(d);(void)(e)
#define CAMLxparam1(a)            (void)(a)
#define CAMLxparam2(a,b)          (void)(a);(void)(b)
#define CAMLxparam3(a,b,c)        (void)(a);(void)(b);(void)(c)
#define CAMLlocal1(a)             value a = 0
#define CAMLlocal2(a,b)           value a = 0, b = 0
#define CAMLlocal3(a,b,c)         value a = 0, b = 0, c = 0
#define CAMLreturn(x)             return (x)
#define CAMLreturn0               return
#define CAMLreturnT(type, x)      return (x)
#endif
