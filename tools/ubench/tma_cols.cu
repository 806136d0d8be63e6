// TMA throughput on column tiles: copy (load + store in place) a [rows=1024][cols] matrix of 32-byte
// elements by column groups of G elements (box = G*32 B x 256 rows, 4 boxes per tile), persistent CTAs,
// double-buffered.  Compares G = 1, 2, 4 (32/64/128-byte box rows) and a row-contiguous copy.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "../../polmux_b200/csrc/pmx_tma.cuh"

template <int G, int NBUF>
__global__ void __launch_bounds__(128) k_cols(const __grid_constant__ CUtensorMap tmap, int tiles_per_bc, int total) {
    extern __shared__ __align__(1024) unsigned char sm[];
    constexpr int TILE = G * 32 * 1024;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + NBUF * TILE);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) pmx_mbar_init(&mbar[i], 1);
        pmx_fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto issue = [&](int tl, int buf) {
        pmx_mbar_expect_tx(&mbar[buf], TILE);
        const int bc = tl / tiles_per_bc, c0 = (tl % tiles_per_bc) * G;
        for (int r0 = 0; r0 < 1024; r0 += 256) pmx_tma_load_3d(sm + buf * TILE + r0 * G * 32, &tmap, c0 * 4, r0, bc, &mbar[buf]);
    };
    int tile = blockIdx.x, it = 0;
    uint32_t ph[NBUF] = {};
    if (tile < total) issue(tile, 0);
    while (tile < total) {
        const int buf = it % NBUF, next = tile + gridDim.x;
        if (NBUF > 1 && next < total) {
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 2) : "memory");
            issue(next, (it + 1) % NBUF);
        }
        pmx_mbar_wait(&mbar[buf], ph[buf]);
        ph[buf] ^= 1u;
        pmx_fence_proxy_async();
        const int bc = tile / tiles_per_bc, c0 = (tile % tiles_per_bc) * G;
        for (int r0 = 0; r0 < 1024; r0 += 256) pmx_tma_store_3d(&tmap, c0 * 4, r0, bc, sm + buf * TILE + r0 * G * 32);
        pmx_tma_commit();
        if (NBUF == 1) {
            pmx_tma_wait_read();
            if (next < total) issue(next, 0);
        }
        tile = next;
        ++it;
    }
    pmx_tma_wait_read();
}

template <int MODE>  // 0: narrow load -> contiguous store, 1: contiguous load -> narrow store
__global__ void __launch_bounds__(128) k_mixed(const __grid_constant__ CUtensorMap tmap, unsigned char* alt, int tiles_per_bc, int total) {
    extern __shared__ __align__(1024) unsigned char sm[];
    constexpr int TILE = 32 * 1024, NBUF = 2;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + NBUF * TILE);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) pmx_mbar_init(&mbar[i], 1);
        pmx_fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto issue = [&](int tl, int buf) {
        pmx_mbar_expect_tx(&mbar[buf], TILE);
        const int bc = tl / tiles_per_bc, c0 = tl % tiles_per_bc;
        if (MODE == 0)
            for (int r0 = 0; r0 < 1024; r0 += 256) pmx_tma_load_3d(sm + buf * TILE + r0 * 32, &tmap, c0 * 4, r0, bc, &mbar[buf]);
        else
            pmx_bulk_load(sm + buf * TILE, alt + (size_t)tl * TILE, TILE, &mbar[buf]);
    };
    int tile = blockIdx.x, it = 0;
    uint32_t ph[NBUF] = {};
    if (tile < total) issue(tile, 0);
    while (tile < total) {
        const int buf = it % NBUF, next = tile + gridDim.x;
        if (next < total) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue(next, (it + 1) % NBUF);
        }
        pmx_mbar_wait(&mbar[buf], ph[buf]);
        ph[buf] ^= 1u;
        pmx_fence_proxy_async();
        const int bc = tile / tiles_per_bc, c0 = tile % tiles_per_bc;
        if (MODE == 1)
            for (int r0 = 0; r0 < 1024; r0 += 256) pmx_tma_store_3d(&tmap, c0 * 4, r0, bc, sm + buf * TILE + r0 * 32);
        else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(alt + (size_t)tile * TILE),
                         "r"(pmx_smem_u32(sm + buf * TILE)), "r"(TILE) : "memory");
        pmx_tma_commit();
        tile = next;
        ++it;
    }
    pmx_tma_wait_read();
}

int main() {
    const int N1 = 1024, N2 = 1024, BC = 8;
    const size_t bytes = (size_t)BC * N1 * N2 * 32;
    void* d;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 1, bytes);
    PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](auto kern, int G, int nbuf, int ctas_per_sm, const char* name) {
        CUtensorMap m;
        cuuint64_t dims[3] = {(cuuint64_t)N2 * 4, (cuuint64_t)N1, (cuuint64_t)BC};
        cuuint64_t strides[2] = {(cuuint64_t)N2 * 32, (cuuint64_t)N1 * N2 * 32};
        cuuint32_t box[3] = {(cuuint32_t)G * 4, 256, 1}, ones[3] = {1, 1, 1};
        CUtensorMapSwizzle sw = G == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : (G == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode failed %d\n", (int)r); return; }
        const int smem = nbuf * G * 32 * 1024 + 64;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int tpb = N2 / G, total = tpb * BC;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            kern<<<148 * ctas_per_sm, 128, smem>>>(m, tpb, total);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-34s CTAs/SM %d: %7.1f us  %6.0f GB/s (read+write)  %s\n", name, ctas_per_sm, ms * 1e3, 2.0 * bytes / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    };
    {
        unsigned char* alt;
        cudaMalloc(&alt, bytes);
        CUtensorMap m;
        cuuint64_t dims[3] = {(cuuint64_t)N2 * 4, (cuuint64_t)N1, (cuuint64_t)BC};
        cuuint64_t strides[2] = {(cuuint64_t)N2 * 32, (cuuint64_t)N1 * N2 * 32};
        cuuint32_t box[3] = {4, 256, 1}, ones[3] = {1, 1, 1};
        enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int smem = 2 * 32 * 1024 + 64;
        cudaFuncSetAttribute(k_mixed<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_mixed<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int mode = 0; mode < 2; ++mode)
            for (int cps : {2, 3}) {
                float ms = 0;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0) k_mixed<0><<<148 * cps, 128, smem>>>(m, alt, N2, N2 * BC);
                    else k_mixed<1><<<148 * cps, 128, smem>>>(m, alt, N2, N2 * BC);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                printf("%-34s CTAs/SM %d: %7.1f us  %6.0f GB/s (read+write)  %s\n", mode == 0 ? "narrow load -> contiguous store" : "contiguous load -> narrow store",
                       cps, ms * 1e3, 2.0 * bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
    }
    for (int c : {3}) run(k_cols<1, 2>, 1, 2, c, "G=1 (32 B rows), 2 buffers");
    return 0;
}
