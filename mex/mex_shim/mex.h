/*
 * mex.h -- stand-in for the MATLAB / Octave MEX header, for native tests only.
 *
 * The build image has neither MATLAB nor Octave, so ssfm_mex.c cannot be compiled against
 * the real header here.  This shim declares the handful of pre-R2018a "separate complex"
 * API calls the gateway uses (the same ones fastexp.c and cmaadaptivefilter.c use) over a
 * plain C struct, so that the very same mexFunction can be linked into a test driver.
 */
#ifndef PMX_MEX_SHIM_H
#define PMX_MEX_SHIM_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef size_t mwSize;

typedef struct mxArray_tag {
    size_t m, n;
    double *pr; /* real plane, column-major */
    double *pi; /* imaginary plane or NULL */
    char *str;  /* char row vector (1 x n) or NULL for a numeric array */
} mxArray;

mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity flag);
void mxDestroyArray(mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
double *mxGetPr(const mxArray *a);
double *mxGetPi(const mxArray *a);
double mxGetScalar(const mxArray *a);
int mxIsChar(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen); /* 0 = ok, 1 = truncated / not a string */
mxArray *mxCreateString(const char *s);
void mexErrMsgTxt(const char *msg); /* longjmps back to the driver */
int mexAtExit(void (*fn)(void));

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

/* driver side */
int mex_shim_call(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]); /* 0 ok, 1 mexErrMsgTxt */
const char *mex_shim_last_error(void);
void mex_shim_run_at_exit(void);

#ifdef __cplusplus
}
#endif
#endif
