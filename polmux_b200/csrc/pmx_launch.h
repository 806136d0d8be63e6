// Per-L launch table: one translation unit per in-CTA FFT length (pmx_passes_inst.cu
// compiled with -DPMX_L=<L>) exports its three pass launchers through this struct.
#pragma once
#include "pmx_common.cuh"

struct PmxLaunchTable {
    int L;
    int cpc;       // columns per CTA in passes A / C
    int rpc;       // rows per CTA in pass B
    int threadsAC, threadsB;
    size_t smemAC, smemB;
    int tw_total;  // cpx entries of the stage-twiddle table for this L
    // grid.x = tiles, grid.y = batch*nfc
    cudaError_t (*setup)();  // opt in to > 48 KB dynamic shared memory
    void (*passA)(dim3 grid, cudaStream_t s, const PassParams& p, const FiberConst& f);
    void (*passB)(dim3 grid, cudaStream_t s, const PassParams& p, const FiberConst& f);
    void (*passC)(dim3 grid, cudaStream_t s, const PassParams& p, const FiberConst& f);
};

const PmxLaunchTable* pmx_get_table(int L);  // nullptr if L is not built

// Host: fill the stage-twiddle table of length L (layout documented in pmx_fft.cuh).
void pmx_fill_stage_twiddles(int L, cpx* out);
