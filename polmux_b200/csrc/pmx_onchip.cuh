// Small fields (Nfft = L <= 4096 samples per column): the whole fiber in ONE kernel launch, the field never
// leaves the SM.  One CTA per realization-column keeps its column in registers (8 Sa per thread, L/8 threads), runs
// the complete matrix_ssfm / scalar_ssfm loop (fiber.m:459-555, 557-636) -- nextstep + checkstep, nonlinear step,
// in-CTA Stockham FFT of the full length, the step's trunks, inverse FFT, attenuation, max |u|^2 -- until the fiber is
// done, and writes the column back once.  HBM traffic: 64 B per Sa and fiber() call instead of 192 B per Sa and step.
//
// Columns of one realization ('sepfields', nfc > 1) couple through the step control (Pmax = max over the columns,
// fiber.m:694-698) and, on the scalar path with the 'x' flag, through sum_j |u_j|^2 of every sample (nl_step,
// fiber.m:793-799): the nfc CTAs of a realization form a THREAD-BLOCK CLUSTER and read each other's maxima / powers
// through distributed shared memory; every CTA runs the (deterministic) scalar step control redundantly, so no
// broadcast is needed and all CTAs leave the loop in the same iteration.
#pragma once
#include <cooperative_groups.h>
#include "pmx_kernels.cuh"

namespace cg = cooperative_groups;

#define PMX_ONCHIP_MAX_L 4096
#define PMX_ONCHIP_MAX_NFC 8    // portable cluster size

template <int L>
struct OnchipSmem {
    static constexpr int T = L / 8;
    static constexpr int R16(int v) { return (v + 15) / 16 * 16; }
    static constexpr int WORK_BYTES = 2 * L * (int)sizeof(cpx);   // exchange buffer (and, before a step, XPM powers)
    static constexpr int TW_OFF = WORK_BYTES;
    static constexpr int SCR_OFF = TW_OFF + R16(pmx_tw_total(L) * (int)sizeof(cpx));
    static constexpr int PKG_OFF = SCR_OFF + PMX_B_SCR * T * (int)sizeof(double2);
    static constexpr int PLATE_OFF = PKG_OFF + (int)sizeof(StepPkg);
    static constexpr int CTL_OFF = PLATE_OFF + PMX_PKG_PLATES * (int)sizeof(PlateConst);
    static constexpr int RED_OFF = CTL_OFF + R16((int)sizeof(StepCtl));
    static constexpr int XCH_OFF = RED_OFF + 32 * 8;              // [2] column maxima, by step parity (read by the cluster)
    static constexpr int GO_OFF = XCH_OFF + 16;
    static constexpr int TOTAL = GO_OFF + 16;
};

// block-wide max of an order-preserving key; the result is valid in thread 0
__device__ __forceinline__ unsigned long long pmx_block_max_t0(unsigned long long key, void* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    unsigned long long* red = reinterpret_cast<unsigned long long*>(scratch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarp = (blockDim.x + 31) >> 5;
    if (lane == 0) red[warp] = key;
    __syncthreads();
    unsigned long long m = 0ull;
    if (threadIdx.x == 0)
        for (int w = 0; w < nwarp; ++w) m = red[w] > m ? red[w] : m;
    return m;
}

template <typename R, int L, bool SC>
__global__ void __launch_bounds__((L / 8) < 32 ? 32 : (L / 8), 1) pmx_k_onchip(PassParams p, FiberConst f) {
    using S = OnchipSmem<L>;
    constexpr int T = L / 8;
    constexpr bool PRE = SC;
    extern __shared__ __align__(16) unsigned char smo[];
    cpx* work = reinterpret_cast<cpx*>(smo);
    cpx* stw = reinterpret_cast<cpx*>(smo + S::TW_OFF);
    StepPkg* st = reinterpret_cast<StepPkg*>(smo + S::PKG_OFF);
    PlateConst* schunk = reinterpret_cast<PlateConst*>(smo + S::PLATE_OFF);
    StepCtl* c = reinterpret_cast<StepCtl*>(smo + S::CTL_OFF);
    void* sred = smo + S::RED_OFF;
    unsigned long long* xch = reinterpret_cast<unsigned long long*>(smo + S::XCH_OFF);
    int* s_go = reinterpret_cast<int*>(smo + S::GO_OFF);
    const int col = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const bool live = t < T;   // (L = 64, 128: the CTA is padded to one warp)
    const int tt = live ? t : 0;   // (padding threads shadow thread 0: same loads, same values, same stores)
    dcpx* scr = reinterpret_cast<dcpx*>(smo + S::SCR_OFF) + tt;
    cg::cluster_group cluster = cg::this_cluster();
    const size_t N = L;
    cpx* fld = reinterpret_cast<cpx*>(p.field) + (size_t)(b * f.nfc + col) * N * 2;
    // position of time sample n in the resident column (transposed four-step layout for 2^12, natural order below)
    auto mem = [&](int n) { return (size_t)(((n & ((1 << p.log2N2) - 1)) << p.log2N1) + (n >> p.log2N2)); };
    PMX_ASSERT(col < f.nfc && b < p.batch && mem(L - 1) < N && (int)blockDim.x >= T && (int)gridDim.x == f.nfc);

    pmx_load_stage_tw<L>(stw, p.tw_stage);
    for (int i = t; i < (int)(sizeof(StepCtl) / 4); i += blockDim.x) reinterpret_cast<int*>(c)[i] = 0;
    cpx x[8], y[8];
    unsigned long long vmax = 0ull;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        ld_sa(fld + 2 * mem(tt + q * T), x[q], y[q]);
        const unsigned long long key = pmx_pow_key((double)power_ref(x[q], y[q]));
        if (live) vmax = key > vmax ? key : vmax;
    }
    __syncthreads();
    int first = 1, parity = 0;
    // a thread's bins after the full-length transform: k = t + q*T, q < 4 positive frequencies, q >= 4 negative
    const double dfn = (double)T * f.inv_nsymb;
    const double fn0 = (double)tt * f.inv_nsymb;
    const double fn4 = (double)(tt + 4 * T - L) * f.inv_nsymb;
    for (;;) {
        // ---- max |u|^2 of every column of the realization (nextstep, fiber.m:693-698)
        const unsigned long long m = pmx_block_max_t0(vmax, sred);
        if (f.nfc > 1) {
            if (t == 0) xch[parity] = m;
            cluster.sync();
            if (t < f.nfc) c->umax_bits[t] = *cluster.map_shared_rank(&xch[parity], t);
            parity ^= 1;
        } else if (t == 0) {
            c->umax_bits[0] = m;
        }
        __syncthreads();
        // ---- nextstep + checkstep + step package (all CTAs of the cluster compute the same schedule)
        const bool go = pmx_ctl_step(c, st, p, f, first, b, s_go);
        __syncthreads();
        if (!go) break;
        first = 0;
        const int ntrunk = st->ntrunk;
        const bool any_full = (ntrunk > 2) || (st->dzb_first == f.lcorr) || (st->dzb_last == f.lcorr);
        if constexpr (PRE) {
            if (ntrunk > 0) pmx_b_pre(scr, T, st, f, col, fn4, fn0, any_full);
        }
        // ---- scalar path with the 'x' flag: sum over the columns of |u|^2 per sample, left to right (fiber.m:793-799)
        if (f.xpm) {
            real* pw = reinterpret_cast<real*>(work);
            if (live) {
#pragma unroll
                for (int q = 0; q < 8; ++q) pw[t + q * T] = R_ADD(R_MUL(x[q].x, x[q].x), R_MUL(x[q].y, x[q].y));
            }
            cluster.sync();
            real sum[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) sum[q] = (real)0;
            for (int k = 0; k < f.nfc; ++k) {
                const real* rp = cluster.map_shared_rank(pw, k);
#pragma unroll
                for (int q = 0; q < 8; ++q) sum[q] = R_ADD(sum[q], rp[tt + q * T]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) y[q] = mkc(sum[q], (real)0);
            cluster.sync();   // everybody has read this CTA's powers: the buffer goes back to the transforms
        }
        pmx_nl_step(x, y, st, f, col);
        // ---- linear step: full-length transform, the step's trunks, inverse transform (= conj o forward o conj)
#pragma unroll 1
        for (int dir = 0; dir < 2; ++dir) {
            CtaFFT<R, L>::run(x, y, work, work + L, tt, stw);
            if (dir == 1) break;
            if (ntrunk > 0)
                pmx_linear_bins<SC, PRE>(x, y, st, f, p, scr, T, schunk, b, col, 0, tt, T, N, fn0, fn4, dfn, any_full);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                x[q] = cconj(x[q]);
                y[q] = cconj(y[q]);
            }
        }
        // ---- 1/N and attenuation exp(-alpha/2*dz) (fiber.m:531-532), running max for the next step
        const real sc = (real)st->scale, nsc = -sc;
        vmax = 0ull;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            x[q] = mkc(x[q].x * sc, x[q].y * nsc);
            y[q] = mkc(y[q].x * sc, y[q].y * nsc);
            const unsigned long long key = pmx_pow_key((double)power_ref(x[q], y[q]));
            if (live) vmax = key > vmax ? key : vmax;
        }
        __syncthreads();   // the package and the exchange buffer are rewritten by the next step
    }
    if (live) {
#pragma unroll
        for (int q = 0; q < 8; ++q) st_sa(fld + 2 * mem(t + q * T), x[q], y[q]);
    }
    if (col == 0)
        for (int i = t; i < (int)(sizeof(StepCtl) / 4); i += blockDim.x)
            reinterpret_cast<int*>(&p.ctl[b])[i] = reinterpret_cast<const int*>(c)[i];
    if (f.nfc > 1) cluster.sync();   // no CTA of the cluster exits while another may still read its shared memory
}
