/*
 * Native driver for the MEX gateway: builds the prhs[] a patched fiber.m would pass, calls
 * mexFunction through the shim, writes the outputs.  Used by tests/test_mex_gateway.py.
 *
 *   test_ssfm_mex in.bin out.bin
 *
 * in.bin : int64 header {nfft, nfc, nplates, manakov, fls[4], has_uy_imag, use_scal, nscal, nspan, namp}
 *          then doubles: dzmaxt dphimaxt alphalin Lf, gam[nfc], uxr uxi uyr uyi [nfft*nfc each],
 *          betat db1 [nfft*nfc each], db0 theta epsilon [nplates*nspan each, span after span], scal[nscal], amp[namp]
 * out.bin: doubles {status, firstdz, ncycle(1)}, uxr uxi uyr uyi, ncycle[nspan]
 *          (status 1 = mexErrMsgTxt, message on stderr)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

static mxArray *vec(FILE *f, size_t m, size_t n, int cplx)
{
    mxArray *a = mxCreateDoubleMatrix(m, n, cplx ? mxCOMPLEX : mxREAL);
    if (m * n != 0 && fread(a->pr, sizeof(double), m * n, f) != m * n) exit(3);
    if (cplx && m * n != 0 && fread(a->pi, sizeof(double), m * n, f) != m * n) exit(3);
    return a;
}
static mxArray *scal(double v)
{
    mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->pr[0] = v;
    return a;
}

int main(int argc, char **argv)
{
    int64_t h[13];
    double s4[4];
    const mxArray *prhs[18];
    mxArray *plhs[4] = {0, 0, 0, 0};
    FILE *f, *o;
    size_t nfft, nfc, n, np, nspan;
    int rc, nrhs, k;
    double hdr[3];
    if (argc != 3) return 2;
    f = fopen(argv[1], "rb");
    if (!f || fread(h, sizeof(int64_t), 13, f) != 13) return 3;
    nfft = (size_t)h[0], nfc = (size_t)h[1], np = (size_t)h[2];
    nspan = (size_t)h[11];
    n = nfft * nfc;
    if (fread(s4, sizeof(double), 4, f) != 4) return 3;
    {
        mxArray *gam = vec(f, 1, nfc, 0);
        mxArray *ux = vec(f, nfft, nfc, 1);
        mxArray *uy = vec(f, nfft, nfc, 1);
        mxArray *betat = vec(f, nfft, nfc, 0);
        mxArray *db1 = vec(f, nfft, nfc, 0);
        mxArray *db0 = vec(f, np, nspan, 0), *theta = vec(f, np, nspan, 0), *eps = vec(f, np, nspan, 0);
        mxArray *fls = mxCreateDoubleMatrix(1, 4, mxREAL);
        mxArray *sc = vec(f, 1, (size_t)h[10], 0);
        mxArray *amp = vec(f, 1, (size_t)h[12], 0);
        if (!h[8]) { /* exercise the "purely real array has no imaginary plane" branch */
            free(uy->pi);
            uy->pi = NULL;
        }
        for (k = 0; k < 4; k++) fls->pr[k] = (double)h[4 + k];
        prhs[0] = ux; prhs[1] = uy; prhs[2] = betat; prhs[3] = db1;
        prhs[4] = scal(s4[0]); prhs[5] = scal(s4[1]); prhs[6] = gam; prhs[7] = scal(s4[2]);
        prhs[8] = scal((double)nfc); prhs[9] = scal(s4[3]); prhs[10] = scal((double)np);
        prhs[11] = scal((double)h[3]); prhs[12] = fls; prhs[13] = db0; prhs[14] = theta; prhs[15] = eps;
        prhs[16] = sc;
        prhs[17] = amp;
        nrhs = h[12] ? 18 : (h[9] ? 17 : 16);
    }
    fclose(f);
    rc = mex_shim_call(4, plhs, nrhs, prhs);
    o = fopen(argv[2], "wb");
    if (!o) return 4;
    hdr[0] = (double)rc;
    hdr[1] = rc ? 0.0 : plhs[2]->pr[0];
    hdr[2] = rc ? 0.0 : plhs[3]->pr[0];
    fwrite(hdr, sizeof(double), 3, o);
    if (rc) {
        fprintf(stderr, "%s\n", mex_shim_last_error());
    } else {
        fwrite(plhs[0]->pr, sizeof(double), n, o);
        fwrite(plhs[0]->pi, sizeof(double), n, o);
        fwrite(plhs[1]->pr, sizeof(double), n, o);
        fwrite(plhs[1]->pi, sizeof(double), n, o);
        fwrite(plhs[3]->pr, sizeof(double), nspan, o);
    }
    fclose(o);
    mex_shim_run_at_exit();
    return 0;
}
