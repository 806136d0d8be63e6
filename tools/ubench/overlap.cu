// Do FP64 FMAs and shared-memory traffic overlap on a B200 SM?  Even warps run DFMA chains, odd warps
// run LDS.128 / STS.128 streams; compare against each running alone (same warp counts).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mix(double* out, int iters, long long* cyc, int mode_even, int mode_odd) {
    extern __shared__ unsigned char sm[];
    unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int mode = (warp & 1) ? mode_odd : mode_even;   // 0 idle, 1 dfma, 2 lds128, 3 sts128, 4 lds+dfma in one warp
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x + k;
    unsigned acc = 0;
    long long t0 = clock64();
    if (mode == 1) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], 1.0000001, 0.5);
    } else if (mode == 2) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                unsigned v0, v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                             : "r"(base + (((threadIdx.x + 32 * k + it) * 16) & 32767)));
                acc ^= v0 ^ v1 ^ v2 ^ v3;
            }
    } else if (mode == 3) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + (((threadIdx.x + 32 * k + it) * 16) & 32767)), "r"(it) : "memory");
    } else if (mode == 4) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                unsigned v0, v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                             : "r"(base + (((threadIdx.x + 32 * k + it) * 16) & 32767)));
                acc ^= v0 ^ v1 ^ v2 ^ v3;
                a[k] = fma(a[k], 1.0000001, 0.5);
                a[(k + 1) & 7] = fma(a[(k + 1) & 7], 1.0000001, 0.25);
            }
    }
    long long t1 = clock64();
    __shared__ long long tmax;
    if (threadIdx.x == 0) tmax = 0;
    __syncthreads();
    atomicMax((unsigned long long*)&tmax, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = tmax;
    double s = acc;
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 1.2345) out[0] = s;
}
int main() {
    long long* cyc; double* out; long long h;
    cudaMalloc(&cyc, 64); cudaMalloc(&out, 64);
    const int iters = 4000, threads = 1024;
    const char* names[] = {"idle", "dfma", "lds128", "sts128", "lds128+2dfma same warp"};
    int combos[][2] = {{1, 0}, {2, 0}, {3, 0}, {1, 1}, {2, 2}, {1, 2}, {1, 3}, {2, 3}, {4, 4}};
    for (auto& c : combos) {
        mix<<<1, threads, 32768>>>(out, iters, cyc, c[0], c[1]);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("even warps: %-8s odd warps: %-8s  %8lld cycles  (%.2f cyc per iteration-of-8 per warp-pair)\n", names[c[0]], names[c[1]], h,
               (double)h / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
