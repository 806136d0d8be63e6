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
 *              regenerates betat/db1 on the device and the two vectors are not uploaded ([] to skip)
 *
 * Span loop in one call (the loop  for k=1:Nspan, fiber(x,flag); ampliflat(G,'gain',opt); end  of
 * ex06_ber.m:110-115; the field crosses PCIe once in and once out, pmx_link_run):
 *
 *   [ux,uy,firstdz,ncycle] = ssfm_mex(..., db0,theta,epsilon, scal, amp)
 *
 *   db0,theta,epsilon   nplates x Nspan: column k holds the waveplates fiber.m:274-276 draws for span k
 *   amp        [gain sigma(1:nfc) seed]: linear power gain 10^(G/10) of the amplifier after every span
 *              (ampliflat.m:62), ASE sigma per column (ampliflat.m:91-106, 0 = noiseless) and the seed of the
 *              device noise generator (span k uses seed+k-1); [] or absent: fibers only
 *   firstdz    first step of span 1;  ncycle  1 x Nspan
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
    pmx_link_desc lk;
    double gam_buf[16], beta_buf[32], sigma_buf[16];
    double *firstdz;
    int32_t *ncycle, *ntot, *status;
    uint64_t *seeds;
    size_t nfft, nfc, n, k, nspan;
    mxArray *uy_out, *work;
    int rc, is_link;

    if (nrhs < 16 || nrhs > 18)
        mexErrMsgTxt("ssfm_mex: 16 to 18 input arguments expected (see the header of ssfm_mex.c).");
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
    /* one column of plates per span: a vector (either orientation) is one span */
    nspan = (d.nplates > 0) ? mxGetNumberOfElements(prhs[14]) / (size_t)d.nplates : 0;
    if (nspan < 1 || nspan * (size_t)d.nplates != mxGetNumberOfElements(prhs[14]) ||
        mxGetNumberOfElements(prhs[13]) != mxGetNumberOfElements(prhs[14]) ||
        mxGetNumberOfElements(prhs[15]) != mxGetNumberOfElements(prhs[14]))
        mexErrMsgTxt("ssfm_mex: db0, theta and epsilon must have nplates elements (nplates x Nspan for a span loop).");
    if (nspan > 1 && mxGetM(prhs[14]) != (size_t)d.nplates)
        mexErrMsgTxt("ssfm_mex: span loop: db0, theta and epsilon must be nplates x Nspan.");
    is_link = nspan > 1 || (nrhs == 18 && mxGetNumberOfElements(prhs[17]) != 0);
    d.plate_sets = 1;
    d.db0 = mxGetPr(prhs[13]);
    d.theta = mxGetPr(prhs[14]);
    d.epsilon = mxGetPr(prhs[15]);
    if (nrhs >= 17 && mxGetNumberOfElements(prhs[16]) != 0) {
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
    /* per-span results live in one scratch matrix owned by the interpreter (freed on error, too) */
    work = mxCreateDoubleMatrix(4 * nspan, 1, mxREAL);
    firstdz = mxGetPr(work);
    ncycle = (int32_t *)(firstdz + nspan);
    ntot = ncycle + nspan;
    status = ntot + nspan;
    seeds = (uint64_t *)(firstdz + 3 * nspan);
    memset(&res, 0, sizeof res);
    res.firstdz = firstdz;
    res.ncycle = ncycle;
    res.ntot = ntot;
    res.status = status;
    if (is_link) {
        memset(&lk, 0, sizeof lk);
        lk.nspan = (int32_t)nspan;
        lk.plate_sets = 1;
        lk.db0 = d.db0; /* nplates x Nspan column-major == [nspan][1][nplates] */
        lk.theta = d.theta;
        lk.epsilon = d.epsilon;
        if (nrhs == 18 && mxGetNumberOfElements(prhs[17]) != 0) {
            const double *a = mxGetPr(prhs[17]);
            if (mxGetNumberOfElements(prhs[17]) != 2 + nfc)
                mexErrMsgTxt("ssfm_mex: amp must be [gain sigma(1:nfc) seed].");
            if (!(a[0] > 0.0))
                mexErrMsgTxt("ssfm_mex: amp(1) is the linear power gain 10^(G/10) and must be positive.");
            lk.gain = a[0];
            for (k = 0; k < nfc; k++)
                sigma_buf[k] = a[1 + k];
            lk.sigma = sigma_buf;
            for (k = 0; k < nspan; k++)
                seeds[k] = (uint64_t)a[1 + nfc] + (uint64_t)k;
            lk.seeds = seeds;
        }
        rc = pmx_link_run(g_ctx, &d, &lk, &io, &res);
    } else {
        rc = pmx_fiber_run(g_ctx, &d, &io, &res);
    }
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
        *mxGetPr(plhs[2]) = firstdz[0];
    }
    if (nlhs > 3) {
        plhs[3] = mxCreateDoubleMatrix(1, nspan, mxREAL);
        for (k = 0; k < nspan; k++)
            mxGetPr(plhs[3])[k] = (double)ncycle[k];
    }
    mxDestroyArray(work);
}
