/* stand-in for R_ext/Rdynload.h (test infrastructure, see ../Rinternals.h) */
#ifndef MOCK_RDYNLOAD_H
#define MOCK_RDYNLOAD_H
#include "../Rinternals.h"
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct mock_dllinfo DllInfo;
int R_registerRoutines(DllInfo *, const void *c, const R_CallMethodDef *call, const void *f, const void *ext);
Rboolean R_useDynamicSymbols(DllInfo *, Rboolean);
#endif
