#pragma once
typedef void* (*DL_FUNC)(); typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef; typedef struct _DllInfo DllInfo;
extern "C" { int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*); int R_useDynamicSymbols(DllInfo*, int); }
