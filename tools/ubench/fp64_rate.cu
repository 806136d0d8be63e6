// FP64 pipe on a B200 SM: DFMA/clk/SM as a function of resident warps and per-warp ILP, and the
// dependent-issue latency (1 warp, ILP 1).   nvcc -arch=sm_100a -O3 fp64_rate.cu -o fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], 1.0000001, 0.5);
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int warps, double* out, long long* cyc) {
    const int iters = 4096;
    k<ILP><<<148, warps * 32>>>(out, iters, cyc);
    k<ILP><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("warps/SM %2d ILP %d: %.1f cycles/iter  -> %.1f DFMA/clk/SM\n", warps, ILP, avg / iters,
           (double)warps * 32 * ILP * iters / avg);
}
int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&cyc, 148 * 8);
    for (int w : {1, 4, 8, 12, 16, 32}) {
        run<1>(w, out, cyc);
        run<2>(w, out, cyc);
        run<4>(w, out, cyc);
        run<8>(w, out, cyc);
    }
    return 0;
}
