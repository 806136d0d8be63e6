// In-CTA Stockham FFT of length L (power of two, 64..4096) on a two-polarization
// row, FP64.  L/8 threads cooperate on one row; each thread keeps 8 samples of
// BOTH polarizations in registers (element q <-> index t + q*T, T = L/8) through
// every stage, so the first stage can be fed straight from the landed tile, the
// last stage leaves the spectrum in registers for the per-bin Jones product,
// and the inverse transform starts from those same registers.  The first stage
// is radix 2 or 4 when log2(L) is not a multiple of 3 (it needs no twiddles),
// every other stage is radix 8; between stages the data is exchanged through
// shared memory with an XOR swizzle (index i -> i ^ ((i>>3)&7), bank-conflict
// free for the scattered stage writes and the strided reads, no padding).
//
// Stage (radix R, Ns = product of earlier radices), butterfly j in [0, L/R):
//   in : v[r] = x[j + r*L/R] * W_{Ns*R}^{(j mod Ns)*r}
//   out: x'[(j - j mod Ns)*R + (j mod Ns) + r*Ns] = DFT_R(v)[r]
// (natural order in, natural order out after the last stage).  Stages: first R0 = 2^(log2 L mod 3) (or 8), then radix 8.
//
// Twiddles of a radix-8 stage: W^k, W^2k, W^4k come from a shared-memory table
// ([3][Ns] per stage, conflict-free), W^3k, W^5k, W^6k, W^7k are four products.
#pragma once
#include "pmx_common.cuh"

#define PMX_SQRT1_2 ((real)0.70710678118654752440)

__host__ __device__ constexpr int pmx_ilog2(int v) { return v <= 1 ? 0 : 1 + pmx_ilog2(v >> 1); }
__host__ __device__ constexpr int pmx_sw(int i) { return i ^ ((i >> 3) & 7); }

// radix of the stage that starts with Ns already done
__host__ __device__ constexpr int pmx_stage_radix(int L, int Ns) {
    return (Ns == 1 && (pmx_ilog2(L) % 3) != 0) ? (1 << (pmx_ilog2(L) % 3)) : 8;
}
// offset (in cpx) of the twiddle block of the stage starting at Ns within the per-L table
__host__ __device__ constexpr int pmx_tw_offset(int L, int Ns) {
    int off = 0;
    int ns = 1;
    while (ns < Ns) {
        int r = pmx_stage_radix(L, ns);
        if (ns > 1) off += 3 * ns;
        ns *= r;
    }
    return off;
}
// L = 1024 in FP64 runs as 8 * 16 * 8 (two exchanges instead of the three of 2 * 8 * 8 * 8, see CtaFFT<R, 1024>):
// its table holds [15][8] twiddles W_128^(k*r) of the radix-16 stage, then [128] W_1024^k of the last radix-8 stage (its
// W^2k and W^4k are two squarings: 4 KB of shared memory and two 128-bit loads per thread and transform less).
#ifndef PMX_F32
#define PMX_FFT_8_16_8 1
#else
#define PMX_FFT_8_16_8 0
#endif
__host__ __device__ constexpr int pmx_tw_layout(int L) { return (PMX_FFT_8_16_8 && L == 1024) ? 1 : 0; }
__host__ __device__ constexpr int pmx_tw_total(int L) {
    return pmx_tw_layout(L) == 1 ? 15 * 8 + 128 : (pmx_tw_offset(L, L) > 0 ? pmx_tw_offset(L, L) : 1);
}

template <bool INV>
__device__ __forceinline__ cpx mul_mj(cpx a) {  // forward: *(-i); inverse: *(+i)
    return INV ? mkc(-a.y, a.x) : mkc(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ void dft2(cpx& a, cpx& b) {
    cpx t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <bool INV>
__device__ __forceinline__ void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
    cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mj<INV>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

template <bool INV>
__device__ __forceinline__ void dft8(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4, cpx& a5, cpx& a6,
                                     cpx& a7) {
    dft4<INV>(a0, a2, a4, a6);  // even -> E0..E3 in a0,a2,a4,a6
    dft4<INV>(a1, a3, a5, a7);  // odd  -> O0..O3 in a1,a3,a5,a7
    // W8^k * O[k]
    cpx o0 = a1;
    cpx o1, o2, o3;
    if (!INV) {
        o1 = mkc((a3.x + a3.y) * PMX_SQRT1_2, (a3.y - a3.x) * PMX_SQRT1_2);
        o2 = mkc(a5.y, -a5.x);
        o3 = mkc((a7.y - a7.x) * PMX_SQRT1_2, -(a7.x + a7.y) * PMX_SQRT1_2);
    } else {
        o1 = mkc((a3.x - a3.y) * PMX_SQRT1_2, (a3.x + a3.y) * PMX_SQRT1_2);
        o2 = mkc(-a5.y, a5.x);
        o3 = mkc(-(a7.x + a7.y) * PMX_SQRT1_2, (a7.x - a7.y) * PMX_SQRT1_2);
    }
    cpx e0 = a0, e1 = a2, e2 = a4, e3 = a6;
    a0 = cadd(e0, o0);
    a1 = cadd(e1, o1);
    a2 = cadd(e2, o2);
    a3 = cadd(e3, o3);
    a4 = csub(e0, o0);
    a5 = csub(e1, o1);
    a6 = csub(e2, o2);
    a7 = csub(e3, o3);
}

// Forward transform only: the inverse is conj o forward o conj, and the callers fold the two
// conjugations into operations they perform anyway (sign of an operand), so one copy of the
// butterfly code serves both directions.  The radix-8 stages after the first run as a loop over
// the stage size (one copy of the stage body; pass B calls the transform from a two-iteration loop),
// which keeps the pass kernels within the instruction cache.
// a + w*b as two FMAs per component, and a - w*b = 2a - (a + w*b) as one more: a twiddled radix-2 butterfly in
// 6 instructions instead of 8 (complex product, add, subtract)
__device__ __forceinline__ void bfly_tw(cpx a, cpx w, cpx b, cpx& s, cpx& d) {
    s = mkc(fma(-w.y, b.y, fma(w.x, b.x, a.x)), fma(w.y, b.x, fma(w.x, b.y, a.y)));
    d = mkc(fma((real)2, a.x, -s.x), fma((real)2, a.y, -s.y));
}

// Forward radix-8 butterfly with the stage twiddles w1..w7 folded into its first layer: inputs a1..a7 are the
// UNtwiddled points, the result is DFT_8(a0, w1*a1, ..., w7*a7).
__device__ __forceinline__ void dft8_tw(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4, cpx& a5, cpx& a6, cpx& a7, cpx w1,
                                        cpx w2, cpx w3, cpx w4, cpx w5, cpx w6, cpx w7) {
    // first layer: pairs (0,4) (2,6) (1,5) (3,7)
    cpx s04, d04, s26, d26, s15, d15, s37, d37;
    bfly_tw(a0, w4, a4, s04, d04);
    bfly_tw(cmul(w2, a2), w6, a6, s26, d26);
    bfly_tw(cmul(w1, a1), w5, a5, s15, d15);
    bfly_tw(cmul(w3, a3), w7, a7, s37, d37);
    // rest of the two radix-4 butterflies (forward: * -i)
    const cpx d26r = mkc(d26.y, -d26.x), d37r = mkc(d37.y, -d37.x);
    const cpx e0 = cadd(s04, s26), e2 = csub(s04, s26), e1 = cadd(d04, d26r), e3 = csub(d04, d26r);
    const cpx o0 = cadd(s15, s37), o2r = csub(s15, s37), o1r = cadd(d15, d37r), o3r = csub(d15, d37r);
    // W8^k * O[k]
    const cpx o1 = mkc((o1r.x + o1r.y) * PMX_SQRT1_2, (o1r.y - o1r.x) * PMX_SQRT1_2);
    const cpx o2 = mkc(o2r.y, -o2r.x);
    const cpx o3 = mkc((o3r.y - o3r.x) * PMX_SQRT1_2, -(o3r.x + o3r.y) * PMX_SQRT1_2);
    a0 = cadd(e0, o0);
    a1 = cadd(e1, o1);
    a2 = cadd(e2, o2);
    a3 = cadd(e3, o3);
    a4 = csub(e0, o0);
    a5 = csub(e1, o1);
    a6 = csub(e2, o2);
    a7 = csub(e3, o3);
}

// one point of both polarizations to / from the exchange buffers.  FP64: two arrays (sx, sy) of 16-byte complex
// numbers; FP32: ONE array of float4 (x, y of a point side by side) at sx, so that an exchange moves 16 bytes per
// shared-memory instruction in both precisions.
__device__ __forceinline__ void pmx_ex_st(cpx* sx, cpx* sy, int o, cpx x, cpx y) {
#ifdef PMX_F32
    reinterpret_cast<float4*>(sx)[o] = make_float4(x.x, x.y, y.x, y.y);
#else
    sx[o] = x;
    sy[o] = y;
#endif
}
__device__ __forceinline__ void pmx_ex_ld(const cpx* sx, const cpx* sy, int o, cpx& x, cpx& y) {
#ifdef PMX_F32
    const float4 v = reinterpret_cast<const float4*>(sx)[o];
    x = make_float2(v.x, v.y);
    y = make_float2(v.z, v.w);
#else
    x = sx[o];
    y = sy[o];
#endif
}

template <typename R, int L>
struct CtaFFT {
    static constexpr int T = L / 8;
    static constexpr int R0 = pmx_stage_radix(L, 1);  // radix of the first stage (2, 4 or 8), no twiddles

    // x[q], y[q] hold element t + q*T on entry and on exit (natural order).  tw: shared-memory
    // copy of the per-L stage table.
    __device__ __forceinline__ static void run(cpx (&x)[8], cpx (&y)[8], cpx* sx, cpx* sy, int t, const cpx* tw) {
        run(x, y, sx, sy, t, tw, [] {});
    }
    // on_free(): called by every thread as soon as the exchange buffer has been read for the last time (before the last
    // butterfly stage), so that a caller whose next tile lands in that buffer can issue its load that much earlier
    template <typename F>
    __device__ __forceinline__ static void run(cpx (&x)[8], cpx (&y)[8], cpx* sx, cpx* sy, int t, const cpx* tw, F&& on_free) {
        constexpr int NB = 8 / R0;
        // ---- first stage
        if constexpr (R0 == 8) {
            dft8<false>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
            dft8<false>(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7]);
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if constexpr (R0 == 4) {
                    dft4<false>(x[b], x[b + 2], x[b + 4], x[b + 6]);
                    dft4<false>(y[b], y[b + 2], y[b + 4], y[b + 6]);
                } else {
                    dft2<false>(x[b], x[b + 4]);
                    dft2<false>(y[b], y[b + 4]);
                }
            }
        }
        if constexpr (R0 == L) {
            on_free();
            return;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j0 = (t + b * T) * R0;
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                pmx_ex_st(sx, sy, pmx_sw(j0 + r), x[b + NB * r], y[b + NB * r]);
            }
        }
        int ns = R0, two = 0;
#pragma unroll 1
        for (;;) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                pmx_ex_ld(sx, sy, pmx_sw(t + q * T), x[q], y[q]);
            }
            __syncthreads();  // everyone has read the exchange buffer: free for the next stage / the caller
            if (ns * 8 >= L) on_free();
            // ---- radix-8 stage with ns sub-transforms done
            const int k = t & (ns - 1);
            {
                const cpx w1 = tw[two + k], w2 = tw[two + ns + k], w4 = tw[two + 2 * ns + k];
                const cpx w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2);
                const cpx w7 = cmul(w4, w3);
                dft8_tw(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], w1, w2, w3, w4, w5, w6, w7);
                dft8_tw(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7], w1, w2, w3, w4, w5, w6, w7);
            }
            if (ns * 8 >= L) break;
            const int j0 = (t - k) * 8 + k;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                pmx_ex_st(sx, sy, pmx_sw(j0 + r * ns), x[r], y[r]);
            }
            two += 3 * ns;
            ns *= 8;
        }
    }
};


#if PMX_FFT_8_16_8
// ---------------------------------------------------------------------------
// L = 1024 as 8 * 16 * 8.  The middle stage is a radix-16 butterfly on 16 points of ONE polarization: the first
// exchange is read back with threads 0..63 taking the X polarization and threads 64..127 the Y polarization
// (16 points = the same 64 data registers as 8 points of both), the second exchange returns to 8 points x both
// polarizations.  Two exchanges and four barriers per transform instead of three and six, 7 % fewer FP64
// instructions (no separate radix-2 stage; the 15 twiddles of the middle stage come from a small table because only
// 8 distinct sets exist).
#define PMX_C16 ((real)0.92387953251128675613)  // cos(pi/8)
#define PMX_S16 ((real)0.38268343236508977173)  // sin(pi/8)

// DFT_16 of (a0, w1*a1, ..., w15*a15), forward; w = table of the 15 twiddles of this thread (w[r-1], stride ws)
__device__ __forceinline__ void dft16_tw(cpx (&a)[16], const cpx* w, int ws) {
    cpx B[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {  // inner DFT_4 over n1 of a[n2 + 4*n1], twiddles folded into its first layer
        cpx s02, d02, s13, d13;
        const cpx e0 = (n2 == 0) ? a[0] : cmul(w[(n2 - 1) * ws], a[n2]);
        bfly_tw(e0, w[(n2 + 8 - 1) * ws], a[n2 + 8], s02, d02);
        bfly_tw(cmul(w[(n2 + 4 - 1) * ws], a[n2 + 4]), w[(n2 + 12 - 1) * ws], a[n2 + 12], s13, d13);
        const cpx d13r = mkc(d13.y, -d13.x);  // * (-i)
        B[n2][0] = cadd(s02, s13);
        B[n2][2] = csub(s02, s13);
        B[n2][1] = cadd(d02, d13r);
        B[n2][3] = csub(d02, d13r);
    }
    // internal twiddles W_16^(n2*k1)
    {
        const cpx t11 = B[1][1], t12 = B[1][2], t13 = B[1][3], t21 = B[2][1], t22 = B[2][2], t23 = B[2][3];
        const cpx t31 = B[3][1], t32 = B[3][2], t33 = B[3][3];
        B[1][1] = mkc(t11.x * PMX_C16 + t11.y * PMX_S16, t11.y * PMX_C16 - t11.x * PMX_S16);        // W^1 = (c, -s)
        B[1][2] = mkc((t12.x + t12.y) * PMX_SQRT1_2, (t12.y - t12.x) * PMX_SQRT1_2);                // W^2
        B[1][3] = mkc(t13.x * PMX_S16 + t13.y * PMX_C16, t13.y * PMX_S16 - t13.x * PMX_C16);        // W^3 = (s, -c)
        B[2][1] = mkc((t21.x + t21.y) * PMX_SQRT1_2, (t21.y - t21.x) * PMX_SQRT1_2);                // W^2
        B[2][2] = mkc(t22.y, -t22.x);                                                               // W^4 = -i
        B[2][3] = mkc((t23.y - t23.x) * PMX_SQRT1_2, -(t23.x + t23.y) * PMX_SQRT1_2);               // W^6
        B[3][1] = mkc(t31.x * PMX_S16 + t31.y * PMX_C16, t31.y * PMX_S16 - t31.x * PMX_C16);        // W^3
        B[3][2] = mkc((t32.y - t32.x) * PMX_SQRT1_2, -(t32.x + t32.y) * PMX_SQRT1_2);               // W^6
        B[3][3] = mkc(-(t33.x * PMX_C16 + t33.y * PMX_S16), t33.x * PMX_S16 - t33.y * PMX_C16);     // W^9 = (-c, s)
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {  // outer DFT_4 over n2: X[k1 + 4*k2]
        cpx b0 = B[0][k1], b1 = B[1][k1], b2 = B[2][k1], b3 = B[3][k1];
        dft4<false>(b0, b1, b2, b3);
        a[k1] = b0;
        a[k1 + 4] = b1;
        a[k1 + 8] = b2;
        a[k1 + 12] = b3;
    }
}

template <typename R>
struct CtaFFT<R, 1024> {
    static constexpr int L = 1024, T = 128;
    __device__ __forceinline__ static void run(cpx (&x)[8], cpx (&y)[8], cpx* sx, cpx* sy, int t, const cpx* tw) {
        run(x, y, sx, sy, t, tw, [] {});
    }
    template <typename F>
    __device__ __forceinline__ static void run(cpx (&x)[8], cpx (&y)[8], cpx* sx, cpx* sy, int t, const cpx* tw, F&& on_free) {
        // ---- stage A: radix 8, no twiddles; butterfly j = t, inputs t + r*128, outputs 8*t + r
        dft8<false>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
        dft8<false>(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int o = pmx_sw(8 * t + r);
            sx[o] = x[r];
            sy[o] = y[r];
        }
        __syncthreads();
        // ---- stage B: radix 16 on one polarization; butterfly j = t & 63, inputs j + r*64, Ns = 8
        const int j = t & 63, k = j & 7;
        cpx* sp = (t < 64) ? sx : sy;
        cpx a[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) a[r] = sp[pmx_sw(j + r * 64)];
        __syncthreads();  // the first exchange is read out
        dft16_tw(a, tw + k, 8);
        {
            const int j0 = (j - k) * 16 + k;  // outputs (j - k)*16 + k + r*8
#pragma unroll
            for (int r = 0; r < 16; ++r) sp[pmx_sw(j0 + r * 8)] = a[r];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int o = pmx_sw(t + q * T);
            x[q] = sx[o];
            y[q] = sy[o];
        }
        __syncthreads();  // everyone has read the exchange buffer: free for the caller
        on_free();
        // ---- stage C: radix 8, Ns = 128: twiddles W_1024^(t*r), natural-order output t + r*128
        {
            const cpx* tc = tw + 15 * 8;
            const cpx w1 = tc[t];
            const cpx w2 = mkc(fma(w1.x, w1.x, -w1.y * w1.y), (real)2 * w1.x * w1.y);
            const cpx w4 = mkc(fma(w2.x, w2.x, -w2.y * w2.y), (real)2 * w2.x * w2.y);
            const cpx w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2);
            const cpx w7 = cmul(w4, w3);
            dft8_tw(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], w1, w2, w3, w4, w5, w6, w7);
            dft8_tw(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7], w1, w2, w3, w4, w5, w6, w7);
        }
    }
};
#endif
