// Instantiates the three pass kernels for one in-CTA FFT length (-DPMX_L=<L>).
#include "pmx_kernels.cuh"
#include "pmx_launch.h"

#ifndef PMX_L
#error "compile with -DPMX_L=<fft length>"
#endif

namespace {
constexpr int L = PMX_L;
// Tile shapes.  Rows/columns per tile are chosen so that a tile is <= 32 KiB (1024 Sa) when
// possible: landing buffer + exchange buffer then fit three CTAs per SM with the next tile
// prefetched.  Longer transforms land in the exchange buffer itself (no separate prefetch).
#ifndef PMX_GAC
constexpr int GAC = (L >= 1024) ? 1 : (L == 512 ? 2 : 4);
#else
constexpr int GAC = PMX_GAC;
#endif
constexpr int GB = (L >= 1024) ? 1 : (L == 512 ? 2 : 4);
#ifndef PMX_PFAC
constexpr bool PFAC = (GAC * L <= 1024);
#else
constexpr bool PFAC = (PMX_PFAC != 0) && (GAC * L <= 1024);
#endif
#ifndef PMX_PFB
// pass B is compute-bound: it prefers a fourth resident CTA to a separate prefetch buffer
constexpr bool PFB = (GB * L <= 512);
#else
constexpr bool PFB = (PMX_PFB != 0) && (GB * L <= 1024);
#endif
using SA = PassSmem<L, GAC, PFAC, 0>;
using SB = PassSmem<L, GB, PFB, 1>;
using SC = PassSmem<L, GAC, PFAC, 2>;

cudaError_t setup(int* ctasA, int* ctasB, int* ctasC) {
    cudaError_t e;
    auto prep = [](auto kern, int smem, int threads, int* ctas) -> cudaError_t {
        cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (r != cudaSuccess) return r;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, kern, threads, smem);
    };
    e = prep(pmx_k_passA<L, GAC, PFAC>, SA::TOTAL, SA::THREADS, ctasA);
    if (e != cudaSuccess) return e;
    e = prep(pmx_k_passB<L, GB, PFB, false>, SB::TOTAL, SB::THREADS, ctasB);
    if (e != cudaSuccess) return e;
    e = prep(pmx_k_passB<L, GB, PFB, true>, SB::TOTAL, SB::THREADS, ctasB);
    if (e != cudaSuccess) return e;
    return prep(pmx_k_passC<L, GAC, PFAC>, SC::TOTAL, SC::THREADS, ctasC);
}
void passA(int gx, cudaStream_t s, const PassParams& p, const FiberConst& f, const CUtensorMap& m) {
    pmx_k_passA<L, GAC, PFAC><<<gx, SA::THREADS, SA::TOTAL, s>>>(p, f, m);
}
void passB(int gx, cudaStream_t s, const PassParams& p, const FiberConst& f, const CUtensorMap& m) {
    if (f.disp_scalar)
        pmx_k_passB<L, GB, PFB, true><<<gx, SB::THREADS, SB::TOTAL, s>>>(p, f, m);
    else
        pmx_k_passB<L, GB, PFB, false><<<gx, SB::THREADS, SB::TOTAL, s>>>(p, f, m);
}
void passC(int gx, cudaStream_t s, const PassParams& p, const FiberConst& f, const CUtensorMap& m) {
    pmx_k_passC<L, GAC, PFAC><<<gx, SC::THREADS, SC::TOTAL, s>>>(p, f, m);
}
}  // namespace

#define PMX_CAT2(a, b) a##b
#define PMX_CAT(a, b) PMX_CAT2(a, b)
extern const PmxLaunchTable PMX_CAT(pmx_table_, PMX_L) = {
    L, GAC, GB, PFAC ? 1 : 0, PFB ? 1 : 0, SA::THREADS, SB::THREADS, (size_t)SA::TOTAL, (size_t)SB::TOTAL,
    pmx_tw_total(L), PmxTw4<L>::LO, PmxTw4<L>::PER, setup, passA, passB, passC};
