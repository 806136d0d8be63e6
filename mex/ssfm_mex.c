/*
 * ssfm_mex.c -- MEX gateway to the B200 split-step Fourier library (libpolmux_ssfm.so).
 *
 * Written in the style of the reference's own gateways (fastexp.c:46-68,
 * cmaadaptivefilter.c:93-174): split real/imaginary storage through mxGetPr/mxGetPi, outputs
 * from mxCreateDoubleMatrix, failures through mexErrMsgTxt, no toolbox / gpuArray dependency.
 *
 * It replaces the body of matrix_ssfm (fiber.m:459-555) and takes that function's argument
 * list, with the three brf vectors passed separately:
 *
 *   [ux,uy,firstdz,ncycle] = ssfm_mex(ux,uy,betat,db1,dzmaxt,dphimaxt,gam,alphalin, ...
 *                                     nfc,Lf,nplates,manakov,fls,db0,theta,epsilon[,scal])
 *
 *   ux, uy     Nfft x nfc complex (uy may be real or empty -> zeros, fiber.m:286)
 *   betat,db1  Nfft x nfc real                        (fiber.m:350-362)
 *   gam        1 x nfc (or scalar) [1/mW/m], before the Manakov 8/9
 *   manakov    logical / 0-1 scalar: strcmp(x.manakov,'yes')
 *   fls        1 x 4 flag vector                      (fiber.m:157)
 *   db0,theta,epsilon   nplates x 1                   (brf.*, fiber.m:266-276)
 *   scal       optional [symbolrate nsymb nt b30 dgdrms beta1(1:nfc) beta2(1:nfc)]: the library
 *              regenerates betat/db1 on the device and the two vectors are not uploaded
 *
 * Build (not verifiable in the image this was written in -- it has no mex.h):
 *   mex ssfm_mex.c -I../include -L../polmux_b200/lib -lpolmux_ssfm
 *   mkoctfile --mex ssfm_mex.c -I../include -L../polmux_b200/lib -lpolmux_ssfm
 * The native test driver mex/test_ssfm_mex.c links this very file against mex_shim/mex.h.
 */
#include <math.h>
#include <string.h>
#include "mex.h"
#include "polmux_ssfm.h"

static pmx_ctx *g_ctx = NULL; /* one device context for the life of the MEX file */

static void ssfm_at_exit(void)
{
    if (g_ctx) {
        pmx_ctx_destroy(g_ctx);
        g_ctx = NULL;
    }
}

static void fail(const char *what)
{
    static char msg[768];
    const char *detail = pmx_last_error(g_ctx);
    strncpy(msg, what, sizeof msg - 1);
    msg[sizeof msg - 1] = 0;
    if (detail && detail[0]) {
        strncat(msg, ": ", sizeof msg - strlen(msg) - 1);
        strncat(msg, detail, sizeof msg - strlen(msg) - 1);
    }
    mexErrMsgTxt(msg);
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    pmx_fiber_desc d;
    pmx_field io;
    pmx_fiber_result res;
    double firstdz = 0.0, gam_buf[16], beta_buf[32];
    int32_t ncycle = 0, ntot = 0, status = 0;
    size_t nfft, nfc, n, k;
    mxArray *uy_out;
    int rc;

    if (nrhs != 16 && nrhs != 17)
        mexErrMsgTxt("ssfm_mex: 16 or 17 input arguments expected (see the header of ssfm_mex.c).");
    if (nlhs > 4)
        mexErrMsgTxt("ssfm_mex: at most 4 outputs [ux,uy,firstdz,ncycle].");

    nfft = mxGetM(prhs[0]);
    nfc = mxGetN(prhs[0]);
    n = nfft * nfc;
    if (n == 0)
        mexErrMsgTxt("ssfm_mex: empty x field.");
    if (nfc > 16)
        mexErrMsgTxt("ssfm_mex: at most 16 field columns.");
    if (mxGetNumberOfElements(prhs[1]) != 0 && (mxGetM(prhs[1]) != nfft || mxGetN(prhs[1]) != nfc))
        mexErrMsgTxt("ssfm_mex: ux and uy must have the same size.");
    if ((size_t)mxGetScalar(prhs[8]) != nfc)
        mexErrMsgTxt("ssfm_mex: nfc does not match the number of columns of ux.");
    if (mxGetNumberOfElements(prhs[12]) != 4)
        mexErrMsgTxt("ssfm_mex: fls must have 4 elements.");

    memset(&d, 0, sizeof d);
    d.nfft = (int64_t)nfft;
    d.nfc = (int32_t)nfc;
    d.batch = 1;
    d.precision = PMX_F64;
    d.dzmaxt = mxGetScalar(prhs[4]);
    d.dphimaxt = mxGetScalar(prhs[5]);
    d.alphalin = mxGetScalar(prhs[7]);
    d.length = mxGetScalar(prhs[9]);
    d.nplates = (int32_t)mxGetScalar(prhs[10]);
    d.manakov = mxGetScalar(prhs[11]) != 0.0;
    for (k = 0; k < 4; k++)
        d.fls[k] = mxGetPr(prhs[12])[k] != 0.0;
    /* gam: scalar (nfc == 1) or one value per column */
    for (k = 0; k < nfc; k++)
        gam_buf[k] = mxGetPr(prhs[6])[mxGetNumberOfElements(prhs[6]) == 1 ? 0 : k];
    d.gam = gam_buf;
    if ((size_t)d.nplates != mxGetNumberOfElements(prhs[13]) || (size_t)d.nplates != mxGetNumberOfElements(prhs[14]) ||
        (size_t)d.nplates != mxGetNumberOfElements(prhs[15]))
        mexErrMsgTxt("ssfm_mex: db0, theta and epsilon must have nplates elements.");
    d.plate_sets = 1;
    d.db0 = mxGetPr(prhs[13]);
    d.theta = mxGetPr(prhs[14]);
    d.epsilon = mxGetPr(prhs[15]);
    if (nrhs == 17 && mxGetNumberOfElements(prhs[16]) != 0) {
        const double *s = mxGetPr(prhs[16]);
        if (mxGetNumberOfElements(prhs[16]) != 5 + 2 * nfc)
            mexErrMsgTxt("ssfm_mex: scal must be [symbolrate nsymb nt b30 dgdrms beta1(1:nfc) beta2(1:nfc)].");
        d.disp_mode = PMX_DISP_SCALAR;
        d.symbolrate = s[0];
        d.nsymb = (int32_t)s[1];
        d.nt = (int32_t)s[2];
        d.b30 = s[3];
        d.dgdrms = s[4];
        for (k = 0; k < 2 * nfc; k++)
            beta_buf[k] = s[5 + k];
        d.beta1 = beta_buf;
        d.beta2 = beta_buf + nfc;
    } else {
        if (mxGetM(prhs[2]) != nfft || mxGetN(prhs[2]) != nfc)
            mexErrMsgTxt("ssfm_mex: betat must be Nfft x nfc.");
        d.disp_mode = PMX_DISP_VECTOR;
        d.betat = mxGetPr(prhs[2]); /* column-major Nfft x nfc == [nfc][nfft] */
        d.db1 = (mxGetNumberOfElements(prhs[3]) == n) ? mxGetPr(prhs[3]) : NULL;
    }

    /* outputs are allocated first and the propagation runs in place on them */
    plhs[0] = mxCreateDoubleMatrix(nfft, nfc, mxCOMPLEX);
    uy_out = mxCreateDoubleMatrix(nfft, nfc, mxCOMPLEX);
    memcpy(mxGetPr(plhs[0]), mxGetPr(prhs[0]), n * sizeof(double));
    if (mxGetPi(prhs[0])) /* purely real arrays carry no imaginary plane (cmaadaptivefilter.c:136-140) */
        memcpy(mxGetPi(plhs[0]), mxGetPi(prhs[0]), n * sizeof(double));
    if (mxGetNumberOfElements(prhs[1]) == n) {
        memcpy(mxGetPr(uy_out), mxGetPr(prhs[1]), n * sizeof(double));
        if (mxGetPi(prhs[1]))
            memcpy(mxGetPi(uy_out), mxGetPi(prhs[1]), n * sizeof(double));
    }
    io.layout = PMX_PLANAR;
    io.reserved = 0;
    io.xr = mxGetPr(plhs[0]);
    io.xi = mxGetPi(plhs[0]);
    io.yr = mxGetPr(uy_out);
    io.yi = mxGetPi(uy_out);

    if (!g_ctx) {
        if (pmx_ctx_create(&g_ctx, 0) != PMX_OK)
            fail("ssfm_mex: no usable B200 (there is no CPU fallback)");
        mexAtExit(ssfm_at_exit);
    }
    memset(&res, 0, sizeof res);
    res.firstdz = &firstdz;
    res.ncycle = &ncycle;
    res.ntot = &ntot;
    res.status = &status;
    rc = pmx_fiber_run(g_ctx, &d, &io, &res);
    if (nlhs > 1)
        plhs[1] = uy_out;
    else
        mxDestroyArray(uy_out);
    if (rc == PMX_ERR_PLATE_INDEX)
        fail("ssfm_mex: index out of bound; value out of bound nplates (fiber.m:910)");
    if (rc != PMX_OK)
        fail("ssfm_mex: propagation failed");

    if (nlhs > 2) {
        plhs[2] = mxCreateDoubleMatrix(1, 1, mxREAL);
        *mxGetPr(plhs[2]) = firstdz;
    }
    if (nlhs > 3) {
        plhs[3] = mxCreateDoubleMatrix(1, 1, mxREAL);
        *mxGetPr(plhs[3]) = (double)ncycle;
    }
}
