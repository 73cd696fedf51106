/* see ../R.h: prototypes only, test infrastructure */
#ifndef NTL_STUB_RDYNLOAD_H
#define NTL_STUB_RDYNLOAD_H
#include "../Rinternals.h"
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; void *types; } R_CMethodDef;
typedef R_CMethodDef R_FortranMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo *, const R_CMethodDef *, const R_CallMethodDef *, const R_FortranMethodDef *,
                       const R_ExternalMethodDef *);
Rboolean R_useDynamicSymbols(DllInfo *, Rboolean);
#endif
