// Minimal TMA / mbarrier wrappers (inline PTX, sm_100a) used to stage field tiles
// from HBM into shared memory ahead of the compute that consumes them.
//
// Tiles are described by 3-D tensor maps over the resident field buffer
// (cuTensorMapEncodeTiled, built on the host in pmx_api.cu):
// (the field is an [N2][N1] matrix of Sa per realization-column, see pmx_kernels.cuh; FP32 fields: floats, 16-byte Sa)
//   rows map   {128-byte line, N*SA/128 lines, batch*nfc}, SWIZZLE_128B            passes A and C (contiguous rows)
//   cols map   {N1*4 reals, N2 rows (pitch N1*SA bytes), batch*nfc}                 pass B (columns k1),
//              box {G*4 reals, <=256 rows, 1}, swizzle = box pitch (32/64/128 B)
// The swizzle makes the thread-per-Sa reads of the landed tile bank-conflict free:
// physical offset = off ^ (((off >> 7) & (pitch/16 - 1)) << 4)  (CuTe Swizzle<B,4,3>).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t pmx_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void pmx_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pmx_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pmx_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pmx_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pmx_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pmx_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pmx_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(pmx_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one box of a 3-D tiled tensor map -> shared memory, completion on an mbarrier
__device__ __forceinline__ void pmx_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(pmx_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(pmx_smem_u32(bar))
        : "memory");
}
// Programmatic dependent launch: let the next kernel of the stream be scheduled early / wait for the previous one.
__device__ __forceinline__ void pmx_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pmx_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// contiguous global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void pmx_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pmx_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(pmx_smem_u32(bar))
                 : "memory");
}
// shared memory -> one box of a 3-D tiled tensor map (bulk async-group completion)
__device__ __forceinline__ void pmx_tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(pmx_smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void pmx_tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores have finished READING shared memory
__device__ __forceinline__ void pmx_tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// swizzled byte offset inside a landed tile; MASK = pitch/16 - 1 (1, 3 or 7), 0 = no swizzle
template <int MASK>
__device__ __forceinline__ uint32_t pmx_swz(uint32_t off) {
    return off ^ (((off >> 7) & MASK) << 4);
}
