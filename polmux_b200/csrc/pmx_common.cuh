// Shared device-side types of the SSFM kernels (sm_100a).
//
// Field layout in HBM: one "Sa" = (xr, xi, yr, yi) = 4 doubles = 32 B, stored as
// two consecutive double2 (X then Y).  field[((b*nfc + c)*N + n)*2 + pol].
// Four-step index split: time n = n1*N2 + n2, frequency k = k1 + N1*k2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cpx;

#define PMX_MAX_NFC 16

enum { PMX_ST_RUN = 0, PMX_ST_LAST = 1, PMX_ST_DONE = 2, PMX_ST_ERROR = 3 };
// Basis bookkeeping of the PMD product.  The reference goes laboratory -> PSP basis of the first trunk at the
// start of every linear step and back at its end (fiber.m:920-921,931-932).  When the nonlinear step is a
// scalar phase (Manakov) or absent, those constant unitary matrices commute with everything between two
// linear steps (phase rotation, FFTs, attenuation; |ux|^2+|uy|^2 is invariant), so the field can stay in
// the PSP basis of the trunk it is in: R(last)*...*R(first)^H of consecutive steps collapses to nothing
// (same trunk) or to the boundary matrix of the plate just left.
enum { PMX_BM_ENTRY_R = 1, PMX_BM_ENTRY_C = 2, PMX_BM_EXIT_R = 4 };

// Per-realization propagation state + the schedule of the step about to run.
// Written by the last CTA of pass C (or by the init kernel), read by passes A/B/C.
struct __align__(16) StepCtl {
    // running state of matrix_ssfm (fiber.m:506-517)
    double zprop;      // end coordinate of the step about to be taken
    double dz;         // dz returned by the last nextstep
    double dz_miss;    // fiber.m:508
    double firstdz;    // fiber.m:516
    int ntot;          // fiber.m:515
    int ncycle;        // fiber.m:506
    int state;         // PMX_ST_*
    int err;           // pmx_status when state == PMX_ST_ERROR
    // schedule of the current step
    double dz_cur;     // dz (or last_step) used by NL, linear and attenuation
    double leff;       // fiber.m:827-831
    double scale;      // exp(-alpha/2*dz_cur) / N   (attenuation fused with the ifft 1/N)
    double dzb_first;  // dzb(1)
    double dzb_last;   // dzb(ntrunk)
    int ntrunk;        // fiber.m:743,749
    int nmem;          // fiber.m:742,748
    int n_first;       // 0-based plate index of trunk k=1:  ntot + 1 - nmem - 1
    int bmode;         // PMX_BM_* : which constant basis changes pass B applies around the trunk product
    // scalar dispersion mode: per-bin-step factors exp(-i*0.5*dgdrms*domega*dzb/lcorr) of the
    // first / last (partial) trunk of the step, domega = spacing of a thread's bins
    double gpf_r, gpf_i, gpl_r, gpl_i;
    // reduction scratch for nextstep
    unsigned long long umax_bits[PMX_MAX_NFC];  // max over n of |ux|^2+|uy|^2, per column
    unsigned int pad_ticket;
    unsigned int pad1;
};

// Per-plate constants, precomputed on the host in IEEE double.
struct __align__(16) PlateConst {
    // matR = Rtheta*Repsilon (fiber.m:910-912), row-major complex
    double r11r, r11i, r12r, r12i, r21r, r21i, r22r, r22i;
    // C = matR(next)^H * matR(this): change of PSP basis at the boundary to the next plate
    double c11r, c11i, c12r, c12i, c21r, c21i, c22r, c22i;
    double db0;       // brf.db0(n)
    double h0r, h0i;  // exp(-i*db0/2): interior-plate phase factor
    double pad;
};

// What the pass kernels need to know about the step in flight, per realization: written by the step-control
// kernel (pmx_k_ctl) once per step, fetched by every tile with ONE bulk copy (cp.async.bulk, same mbarrier as the
// tile itself), so no pass thread ever waits on a dependent global load of step state.
#define PMX_PKG_PLATES 16
struct __align__(16) StepPkg {
    double dz_cur, leff, scale, dzb_first, dzb_last, gpf_r, gpf_i, gpl_r, gpl_i, db0_last;
    int ntrunk, n_first, bmode, state;  // -- 96-byte header: all passes
    double E[8];   // pass B entry matrix (row-major re,im): R(first)^H, or the boundary matrix of the plate before
    double X[8];   // pass B exit matrix R(last)
    PlateConst plates[PMX_PKG_PLATES];  // the first trunks of the step (more are read from the plate array)
};
#define PMX_PKG_HEAD 96

// Constants of one fiber() call (kernel parameter, by value).
struct FiberConst {
    double Lf, alphalin, halfalpha, dzmax, phimax, lcorr, invN;
    double gam[PMX_MAX_NFC];  // after the Manakov 8/9 (fiber.m:500)
    int nplates, nfc, spm, manakov, pmd, gvd_any, plate_sets, trace_cap;
    int keep_basis, pad_;  // keep_basis: pmd && (manakov || !spm), see PMX_BM_*
    // scalar dispersion mode (fiber.m:350-362 regenerated per bin instead of read from HBM):
    //   omega = w0*fn, fn = kk/NSYMB (kk = signed FFT bin), betat = omega*beta1 + 0.5*omega^2*beta2
    //   + omega^3*b30/6, db1 = dgdrms*omega
    int disp_scalar;
    unsigned nfc_magic;   // ceil(2^32/nfc): bc / nfc = umulhi(bc, nfc_magic) for nfc > 1
    double w0, inv_nsymb, b30_6, dgdrms, domega;  // domega = w0*NT/8: spacing of a thread's bins
    double g1r, g1i;                              // exp(-i*0.5*dgdrms*domega)
    double beta1[PMX_MAX_NFC], beta2[PMX_MAX_NFC];
};

struct PassParams {
    cpx* field;             // [batch*nfc][N][2]
    StepCtl* ctl;           // [batch]
    const cpx* tw_stage;    // in-CTA FFT stage twiddles for this L
    const double* betat_p;  // [nfc][N1][N2] permuted so that bin k1 + N1*k2 sits at k1*N2 + k2
    const double* db1_p;    // same layout
    const PlateConst* plates;  // [plate_sets][nplates]
    StepPkg* pkg;           // [batch]
    const cpx* tw4;         // four-step twiddle rows for this pass: [rows][PmxTw4<L>::PER], see pmx_kernels.cuh
    double* trace_dz;       // [batch][trace_cap] or null
    int* trace_ntrunk;
    int N1, N2, log2N1, log2N2;
    int batch;
    int bc0;                // first realization-column of this launch inside the field (tensor-map coordinate offset)
    int stagger;            // cycles by which the CTAs sharing an SM start apart (de-phases their load / exchange / math phases)
    int reverse;            // walk the tile list backwards (alternates from pass to pass: the tiles the previous
                            // pass wrote last are still in L2 when this pass reads them first)
    long long* dbg;         // optional per-CTA phase cycle counters (PMX_TIMING builds only)
};

__device__ __forceinline__ cpx cmul(cpx a, cpx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cpx cmulc(cpx a, cpx b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cpx cscale(cpx a, double s) { return make_double2(a.x * s, a.y * s); }

// One Sa (both polarizations, 32 B) per 256-bit global access (LDG.E.256 / STG.E.256 on sm_100a).
__device__ __forceinline__ void ld_sa(const cpx* p, cpx& x, cpx& y) {
    asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x.x), "=d"(x.y), "=d"(y.x), "=d"(y.y) : "l"(p));
}
__device__ __forceinline__ void st_sa(cpx* p, cpx x, cpx y) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x.x), "d"(x.y), "d"(y.x), "d"(y.y) : "memory");
}

// Branch-free sin/cos: three-term Cody-Waite reduction by pi/2 carried out with FMAs (each product q*c is
// exact inside the FMA, so the reduction stays accurate far beyond the |x| < 105615 range of the library's own
// fast path: the absolute error of the reduced argument is about 2^-53 * |x| * 6e-17, i.e. below one ulp of
// the result for |x| up to ~1e9 rad and never worse than the spacing of the doubles around x), then the fdlibm
// kernel polynomials on [-pi/4, pi/4]; <= ~1 ulp like the libm calls the reference makes in fastexp.c:41-42.
// No slow path and no branch: the compiler interleaves the evaluations a thread needs.
__device__ __forceinline__ void pmx_sincos_fast(double x, double* sp, double* cp) {
    const double q = rint(x * 6.3661977236758138e-01);
    double r = fma(q, -1.5707963267948966e+00, x);
    r = fma(q, -6.1232339957367574e-17, r);
    r = fma(q, -1.4973849048591698e-33, r);
    const int n = (int)(long long)q;
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sn = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double a = (n & 1) ? cs : sn, b = (n & 1) ? sn : cs;
    *sp = (n & 2) ? -a : a;
    *cp = ((n + 1) & 2) ? -b : b;
}
__device__ __forceinline__ void pmx_sincos(double x, double* sp, double* cp) { pmx_sincos_fast(x, sp, cp); }
__device__ __forceinline__ void pmx_sincos8(const double (&x)[8], double (&s)[8], double (&c)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) pmx_sincos_fast(x[q], &s[q], &c[q]);
}

// |ux|^2+|uy|^2 in the reference's order, no FMA contraction (fiber.m:694)
__device__ __forceinline__ double power_ref(cpx x, cpx y) {
    double p = __dadd_rn(__dmul_rn(x.x, x.x), __dmul_rn(x.y, x.y));
    p = __dadd_rn(p, __dmul_rn(y.x, y.x));
    p = __dadd_rn(p, __dmul_rn(y.y, y.y));
    return p;
}
