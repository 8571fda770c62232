/* Minimal stand-in for MATLAB's / Octave's mex.h, used ONLY to compile-check
 * sbd_mex.c in an image that has neither MATLAB nor Octave.
 * A real build uses the real header:  mex -I../../include sbd_mex.c -L../lib -lsbd
 * or  mkoctfile --mex ...  */
#ifndef SBD_STUB_MEX_H
#define SBD_STUB_MEX_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
void mexLock(void);
int mexAtExit(void (*fn)(void));
int mxIsDouble(const mxArray*);
int mxIsComplex(const mxArray*);
int mxIsChar(const mxArray*);
int mxIsStruct(const mxArray*);
int mxIsEmpty(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
double* mxGetPr(const mxArray*);
double* mxGetPi(const mxArray*);
double mxGetScalar(const mxArray*);
int mxGetString(const mxArray*, char* buf, mwSize len);
mxArray* mxGetField(const mxArray*, mwIndex i, const char* name);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names);
void mxSetField(mxArray*, mwIndex i, const char* name, mxArray* v);
void* mxCalloc(size_t n, size_t sz);
void mxFree(void*);
#ifdef __cplusplus
}
#endif
#endif
