// Column tiles without TMA: (a) cp.async 16 B per thread into shared memory (double-buffered), store with
// st.global.v4.f64 from shared memory; (b) LDG.256 / STG.256 straight through registers.  Same matrix and tile walk
// as tma_cols.cu (1024 x 1024 Sa of 32 B, 8 matrices, one column per tile).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k_cpasync(double4* m, int total) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int N2 = 1024;
    auto issue = [&](int tl, int buf) {
        const int bc = tl >> 10, c = tl & 1023;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(m + (size_t)bc * 1024 * N2 + c);
        for (int i = threadIdx.x; i < 2048; i += 128) {  // 16-byte pieces: row = i >> 1, half = i & 1
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + buf * 32768 + i * 16);
            const unsigned char* s = src + (size_t)(i >> 1) * N2 * 32 + (i & 1) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(s) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int tile = blockIdx.x, it = 0;
    if (tile < total) issue(tile, 0);
    while (tile < total) {
        const int next = tile + gridDim.x, buf = it & 1;
        if (next < total) {
            issue(next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int bc = tile >> 10, c = tile & 1023;
        double4* dst = m + (size_t)bc * 1024 * N2 + c;
        for (int r = threadIdx.x; r < 1024; r += 128) dst[(size_t)r * N2] = *reinterpret_cast<const double4*>(sm + buf * 32768 + r * 32);
        __syncthreads();
        tile = next;
        ++it;
    }
}
__global__ void __launch_bounds__(128) k_ldg(double4* m, int total) {
    const int N2 = 1024;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int bc = tile >> 10, c = tile & 1023;
        double4* p = m + (size_t)bc * 1024 * N2 + c;
        double4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = p[(size_t)(threadIdx.x + q * 128) * N2];
#pragma unroll
        for (int q = 0; q < 8; ++q) p[(size_t)(threadIdx.x + q * 128) * N2] = v[q];
    }
}
int main() {
    const size_t bytes = (size_t)8 * 1024 * 1024 * 32;
    double4* d;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int cps : {3, 6}) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_cpasync<<<148 * cps, 128, 65536>>>(d, 8192);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("cp.async 16 B pieces, %d CTAs/SM: %7.1f us %6.0f GB/s  %s\n", cps, ms * 1e3, 2.0 * bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    for (int cps : {3, 6, 12}) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_ldg<<<148 * cps, 128>>>(d, 8192);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("LDG.256/STG.256 registers, %d CTAs/SM: %7.1f us %6.0f GB/s  %s\n", cps, ms * 1e3, 2.0 * bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
