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
__host__ __device__ constexpr int pmx_tw_total(int L) { return pmx_tw_offset(L, L) > 0 ? pmx_tw_offset(L, L) : 1; }

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
        if constexpr (R0 == L) return;
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

