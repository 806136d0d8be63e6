// Instantiates the three pass kernels for one in-CTA FFT length (-DPMX_L=<L>).
#include "pmx_kernels.cuh"
#include "pmx_launch.h"

#ifndef PMX_L
#error "compile with -DPMX_L=<fft length>"
#endif

namespace {
constexpr int L = PMX_L;
constexpr int T = L / 8;
// CTA shapes: >=128 threads, columns grouped for 64..512 B contiguous runs in passes A/C
constexpr int CPC = (L <= 256) ? (128 / T) : (L == 512 ? 4 : (L <= 2048 ? 2 : 1));
constexpr int RPC = (T >= 128) ? 1 : (128 / T);

cudaError_t setup() {
    cudaError_t e;
    // ask for the largest shared-memory carveout so that several CTAs fit per SM
    cudaFuncSetAttribute(pmx_k_passA<L, CPC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(pmx_k_passB<L, RPC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(pmx_k_passC<L, CPC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    e = cudaFuncSetAttribute(pmx_k_passA<L, CPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)PmxSmem<L, CPC>::bytes(CPC));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pmx_k_passC<L, CPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)PmxSmem<L, CPC>::bytes(CPC));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pmx_k_passB<L, RPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)PmxSmem<L, RPC>::bytes(RPC));
    return e;
}
void passA(dim3 grid, cudaStream_t s, const PassParams& p, const FiberConst& f) {
    pmx_k_passA<L, CPC><<<grid, CPC * T, PmxSmem<L, CPC>::bytes(CPC), s>>>(p, f);
}
void passB(dim3 grid, cudaStream_t s, const PassParams& p, const FiberConst& f) {
    pmx_k_passB<L, RPC><<<grid, RPC * T, PmxSmem<L, RPC>::bytes(RPC), s>>>(p, f);
}
void passC(dim3 grid, cudaStream_t s, const PassParams& p, const FiberConst& f) {
    pmx_k_passC<L, CPC><<<grid, CPC * T, PmxSmem<L, CPC>::bytes(CPC), s>>>(p, f);
}
}  // namespace

#define PMX_CAT2(a, b) a##b
#define PMX_CAT(a, b) PMX_CAT2(a, b)
extern const PmxLaunchTable PMX_CAT(pmx_table_, PMX_L) = {
    L, CPC, RPC, CPC * T, RPC * T, PmxSmem<L, CPC>::bytes(CPC), PmxSmem<L, RPC>::bytes(RPC),
    pmx_tw_total(L), setup, passA, passB, passC};
