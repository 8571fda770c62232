/* stub_mx.c - a minimal mxArray runtime, TEST INFRASTRUCTURE ONLY.
 *
 * MATLAB and Octave are absent from this image, so mex/sbd_mex.c could only be
 * compile-checked.  This file implements the ~25 functions of the MEX API that
 * the gateway uses (mex/stub/mex.h) on top of a plain C struct, so that the REAL
 * mexFunction of sbd_mex.c can be linked against libsbd.so and executed from the
 * tests (tests/mexrt/runtime.py drives it through ctypes).  Semantics follow the
 * documented MEX API: column-major doubles, separate real/imaginary storage,
 * 1x1 struct arrays with named fields, mexErrMsgIdAndTxt does not return.
 */
#define _POSIX_C_SOURCE 200809L
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

enum { K_DOUBLE = 0, K_CHAR = 1, K_STRUCT = 2 };
struct mxArray_tag {
    int kind;
    size_t m, n;
    double *pr, *pi;
    char* str;
    int nfields;
    char** names;
    mxArray** fields;
};

static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static char g_err_id[128], g_err_msg[1024];
static int g_locked = 0;
static void (*g_atexit)(void) = NULL;

void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    snprintf(g_err_id, sizeof g_err_id, "%s", id ? id : "");
    va_start(ap, fmt);
    vsnprintf(g_err_msg, sizeof g_err_msg, fmt, ap);
    va_end(ap);
    if (g_jmp_armed) longjmp(g_jmp, 1);
    fprintf(stderr, "mexErrMsgIdAndTxt outside a call: %s: %s\n", g_err_id, g_err_msg);
    abort();
}
void mexLock(void) { g_locked++; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }

static mxArray* mk(int kind, size_t m, size_t n) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->kind = kind; a->m = m; a->n = n;
    return a;
}
int mxIsDouble(const mxArray* a) { return a && a->kind == K_DOUBLE; }
int mxIsComplex(const mxArray* a) { return a && a->kind == K_DOUBLE && a->pi != NULL; }
int mxIsChar(const mxArray* a) { return a && a->kind == K_CHAR; }
int mxIsStruct(const mxArray* a) { return a && a->kind == K_STRUCT; }
int mxIsEmpty(const mxArray* a) { return !a || a->m * a->n == 0; }
mwSize mxGetNumberOfDimensions(const mxArray* a) { (void)a; return 2; }
size_t mxGetM(const mxArray* a) { return a->m; }
size_t mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
double* mxGetPr(const mxArray* a) { return a->pr; }
double* mxGetPi(const mxArray* a) { return a->pi; }
double mxGetScalar(const mxArray* a) {
    if (!a || a->kind != K_DOUBLE || a->m * a->n == 0) mexErrMsgIdAndTxt("stub:scalar", "mxGetScalar of an empty or non-numeric array");
    return a->pr[0];
}
int mxGetString(const mxArray* a, char* buf, mwSize len) {
    if (!a || a->kind != K_CHAR || strlen(a->str) + 1 > len) return 1;
    strcpy(buf, a->str);
    return 0;
}
mxArray* mxGetField(const mxArray* a, mwIndex i, const char* name) {
    int f;
    if (!a || a->kind != K_STRUCT || i != 0) return NULL;
    for (f = 0; f < a->nfields; ++f)
        if (!strcmp(a->names[f], name)) return a->fields[f];
    return NULL;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    mxArray* a = mk(K_DOUBLE, m, n);
    a->pr = (double*)calloc((m * n) != 0 ? m * n : 1, sizeof(double));
    if (c == mxCOMPLEX) a->pi = (double*)calloc((m * n) != 0 ? m * n : 1, sizeof(double));
    return a;
}
mxArray* mxCreateDoubleScalar(double v) {
    mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->pr[0] = v;
    return a;
}
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names) {
    int f;
    mxArray* a = mk(K_STRUCT, m, n);
    a->nfields = nfields;
    a->names = (char**)calloc(nfields ? nfields : 1, sizeof(char*));
    a->fields = (mxArray**)calloc(nfields ? nfields : 1, sizeof(mxArray*));
    for (f = 0; f < nfields; ++f) a->names[f] = strdup(names[f]);
    return a;
}
void mxSetField(mxArray* a, mwIndex i, const char* name, mxArray* v) {
    int f;
    if (!a || a->kind != K_STRUCT || i != 0) mexErrMsgIdAndTxt("stub:field", "mxSetField on a non-struct");
    for (f = 0; f < a->nfields; ++f)
        if (!strcmp(a->names[f], name)) { a->fields[f] = v; return; }
    mexErrMsgIdAndTxt("stub:field", "mxSetField: no field '%s'", name);
}
void* mxCalloc(size_t n, size_t sz) { return calloc(n, sz); }
void mxFree(void* p) { free(p); }

/* ---- helpers for the Python driver (not part of the MEX API) -------------------------------- */
mxArray* stub_string(const char* s) {
    mxArray* a = mk(K_CHAR, 1, strlen(s));
    a->str = strdup(s);
    return a;
}
/* struct with room for `nfields` fields added one by one */
mxArray* stub_struct(int nfields) {
    mxArray* a = mk(K_STRUCT, 1, 1);
    a->names = (char**)calloc(nfields ? nfields : 1, sizeof(char*));
    a->fields = (mxArray**)calloc(nfields ? nfields : 1, sizeof(mxArray*));
    return a;
}
void stub_add_field(mxArray* a, const char* name, mxArray* v) {
    a->names[a->nfields] = strdup(name);
    a->fields[a->nfields] = v;
    a->nfields++;
}
int stub_kind(const mxArray* a) { return a ? a->kind : -1; }
int stub_nfields(const mxArray* a) { return a->nfields; }
const char* stub_field_name(const mxArray* a, int f) { return a->names[f]; }
mxArray* stub_field_value(const mxArray* a, int f) { return a->fields[f]; }
const char* stub_last_error(void) { return g_err_msg; }
const char* stub_last_error_id(void) { return g_err_id; }
int stub_lock_count(void) { return g_locked; }
void stub_run_atexit(void) { if (g_atexit) g_atexit(); }
void stub_free(mxArray* a) {
    int f;
    if (!a) return;
    free(a->pr); free(a->pi); free(a->str);
    for (f = 0; f < a->nfields; ++f) { free(a->names[f]); stub_free(a->fields[f]); }
    free(a->names); free(a->fields); free(a);
}
/* calls the gateway; returns 0, or 1 if it raised through mexErrMsgIdAndTxt */
int stub_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    int rc = 0;
    g_err_id[0] = g_err_msg[0] = 0;
    g_jmp_armed = 1;
    if (setjmp(g_jmp) == 0) mexFunction(nlhs, plhs, nrhs, prhs);
    else rc = 1;
    g_jmp_armed = 0;
    return rc;
}
