// Column tiles through a CTA pair (cluster of 2 or 4): the cluster loads a slab of C adjacent columns
// (C*32-byte box rows) -- CTA r takes rows [r*1024/C, (r+1)*1024/C) -- then every CTA pulls ITS column out of all
// the slab parts through distributed shared memory, and the way back for the store.  Data movement only:
// is a wide-row TMA + DSMEM transpose faster than the 32-byte-row TMA the column passes use today?
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cooperative_groups.h>
#include "../../polmux_b200/csrc/pmx_tma.cuh"
namespace cg = cooperative_groups;

template <int C>
__global__ void __launch_bounds__(128) k_cluster(const __grid_constant__ CUtensorMap tmap, int slabs_per_bc, int total) {
    extern __shared__ __align__(1024) unsigned char sm[];
    constexpr int ROWS = 1024 / C, PART = ROWS * C * 32;  // bytes of this CTA's part of the slab (= 32 KB)
    unsigned char* land = sm;                // [ROWS][C*32 B]  (no swizzle in this test)
    unsigned char* col = sm + PART;          // [1024][32 B] this CTA's column, gathered
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + 2 * PART);
    cg::cluster_group cl = cg::this_cluster();
    const unsigned r = cl.block_rank();
    if (threadIdx.x == 0) {
        pmx_mbar_init(mbar, 1);
        pmx_fence_mbar_init();
    }
    cl.sync();
    uint32_t ph = 0;
    for (int slab = blockIdx.x / C; slab < total; slab += gridDim.x / C) {
        const int bc = slab / slabs_per_bc, c0 = (slab % slabs_per_bc) * C;
        if (threadIdx.x == 0) {
            pmx_fence_proxy_async();
            pmx_mbar_expect_tx(mbar, PART);
            for (int r0 = 0; r0 < ROWS; r0 += 256) pmx_tma_load_3d(land + r0 * C * 32, &tmap, c0 * 4, r * ROWS + r0, bc, mbar);
        }
        pmx_mbar_wait(mbar, ph);
        ph ^= 1u;
        cl.sync();  // every part has landed
        // gather column r: rows of part p live in CTA p
        for (int p = 0; p < C; ++p) {
            const unsigned char* src = (const unsigned char*)cl.map_shared_rank(land, p);
            for (int i = threadIdx.x; i < ROWS * 2; i += blockDim.x) {  // 16-byte pieces of the 32-byte Sa
                const int row = i >> 1, h = i & 1;
                *reinterpret_cast<double2*>(col + ((p * ROWS + row) * 32 + h * 16)) =
                    *reinterpret_cast<const double2*>(src + (row * C * 32 + r * 32 + h * 16));
            }
        }
        cl.sync();  // every CTA has read the parts
        // scatter the column back into the parts
        for (int p = 0; p < C; ++p) {
            unsigned char* dst = (unsigned char*)cl.map_shared_rank(land, p);
            for (int i = threadIdx.x; i < ROWS * 2; i += blockDim.x) {
                const int row = i >> 1, h = i & 1;
                *reinterpret_cast<double2*>(dst + (row * C * 32 + r * 32 + h * 16)) =
                    *reinterpret_cast<const double2*>(col + ((p * ROWS + row) * 32 + h * 16));
            }
        }
        cl.sync();  // parts complete
        if (threadIdx.x == 0) {
            pmx_fence_proxy_async();
            for (int r0 = 0; r0 < ROWS; r0 += 256) pmx_tma_store_3d(&tmap, c0 * 4, r * ROWS + r0, bc, land + r0 * C * 32);
            pmx_tma_commit();
            pmx_tma_wait_read();
        }
        __syncthreads();
    }
    cl.sync();
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
    auto run = [&](auto kern, int C, int ctas_per_sm) {
        CUtensorMap m;
        cuuint64_t dims[3] = {(cuuint64_t)N2 * 4, (cuuint64_t)N1, (cuuint64_t)BC};
        cuuint64_t strides[2] = {(cuuint64_t)N2 * 32, (cuuint64_t)N1 * N2 * 32};
        cuuint32_t box[3] = {(cuuint32_t)C * 4, 256, 1}, ones[3] = {1, 1, 1};
        CUresult rr = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr) { printf("encode failed %d\n", (int)rr); return; }
        const int smem = 2 * 32 * 1024 + 64;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((148 / C) * C * ctas_per_sm);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const int spb = N2 / C, total = spb * BC;
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            cudaLaunchKernelEx(&cfg, kern, m, spb, total);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("cluster of %d (%3d-byte box rows), %d CTAs/SM: %7.1f us  %6.0f GB/s (read+write)  %s\n", C, C * 32, ctas_per_sm, ms * 1e3,
               2.0 * bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int c : {2, 3}) run(k_cluster<2>, 2, c);
    for (int c : {2, 3}) run(k_cluster<4>, 4, c);
    return 0;
}
