/* Implementation of the mex.h stand-in (see mex.h). */
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

static jmp_buf g_jmp;
static char g_err[1024];
static void (*g_at_exit)(void) = NULL;

mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity flag)
{
    mxArray *a = (mxArray *)calloc(1, sizeof *a);
    size_t cnt = m * n != 0 ? m * n : 1;
    a->m = m;
    a->n = n;
    a->pr = (double *)calloc(cnt, sizeof(double));
    a->pi = flag == mxCOMPLEX ? (double *)calloc(cnt, sizeof(double)) : NULL;
    return a;
}
void mxDestroyArray(mxArray *a)
{
    if (!a) return;
    free(a->pr);
    free(a->pi);
    free(a->str);
    free(a);
}
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
double *mxGetPr(const mxArray *a) { return a->pr; }
double *mxGetPi(const mxArray *a) { return a->pi; }
double mxGetScalar(const mxArray *a) { return (a->m * a->n != 0) ? a->pr[0] : 0.0; }
int mxIsChar(const mxArray *a) { return a && a->str != NULL; }
int mxGetString(const mxArray *a, char *buf, mwSize buflen)
{
    if (!a || !a->str || buflen == 0) return 1;
    strncpy(buf, a->str, buflen - 1);
    buf[buflen - 1] = 0;
    return strlen(a->str) >= buflen;
}
mxArray *mxCreateString(const char *s)
{
    mxArray *a = (mxArray *)calloc(1, sizeof *a);
    a->m = 1;
    a->n = strlen(s);
    a->str = (char *)malloc(a->n + 1);
    memcpy(a->str, s, a->n + 1);
    return a;
}
void mexErrMsgTxt(const char *msg)
{
    strncpy(g_err, msg, sizeof g_err - 1);
    g_err[sizeof g_err - 1] = 0;
    longjmp(g_jmp, 1);
}
int mexAtExit(void (*fn)(void))
{
    g_at_exit = fn;
    return 0;
}
int mex_shim_call(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    g_err[0] = 0;
    if (setjmp(g_jmp)) return 1;
    mexFunction(nlhs, plhs, nrhs, prhs);
    return 0;
}
const char *mex_shim_last_error(void) { return g_err; }
void mex_shim_run_at_exit(void)
{
    if (g_at_exit) g_at_exit();
    g_at_exit = NULL;
}
