// Shared device helpers for the despawn_b200 kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "despawn_b200 kernels are written for sm_100a (B200) only"
#endif

namespace dsp {

constexpr int kChunkDim = 16;                       // chunk.rs:7
constexpr int kChunkVoxels = 4096;                  // 16^3
constexpr int kChunkBytes = kChunkVoxels * 2;       // u16 voxels
constexpr int kRows = 256;                          // x-rows per chunk: r = y + 16*z
constexpr int kFaceWords = 30;                      // 6 vertices x 5 f32
constexpr int kFaceBytes = 120;
constexpr int kSlotAlign = 128;
constexpr int kMaxFaces = 12288;                    // 3-D checkerboard

// ---- mbarrier / bulk-copy (TMA 1-D) PTX wrappers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// plain arrival: with an arrival count of 1 and no bytes expected the phase completes at once
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP), completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy, tracked by the per-thread bulk async-group.
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// global -> L2 bulk prefetch (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // <= N groups still reading their smem source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (bulk store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// four consecutive u64 (32-byte aligned group) with two 128-bit volatile loads
__device__ __forceinline__ void ld_volatile_u64x4(const uint64_t* p, uint64_t (&v)[4]) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "l"(p) : "memory");
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v[2]), "=l"(v[3]) : "l"(p + 2) : "memory");
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Programmatic dependent launch: when the host launches a kernel with the "programmatic stream serialization" attribute, its
// CTAs may be scheduled while the previous kernel of the stream is still draining; pdl_wait() then blocks until that kernel has
// completed and its memory is visible.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// system-scope acquire / release on 32-bit flags: region-seam hand-shake between GPUs (the writer sits across NVLink)
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// device-scope acquire / release: a word one kernel publishes for another kernel running beside it on the same GPU
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy writes to global memory (observed through an acquire) before an async-proxy (bulk copy) read of them
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// L2-coherent 128-bit load (no L1, no read-only path): data another GPU wrote into this GPU's memory during the kernel
__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u16(const uint16_t* p) {
    uint16_t v;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
    return v;
}

// ---- builder-defined terrain spec (DESIGN.md "Terrain spec"); integer-only ---------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t hash2(uint32_t seed, int32_t a, int32_t b) {
    return mix32(seed ^ mix32(static_cast<uint32_t>(a) * 0x9E3779B1U ^ mix32(static_cast<uint32_t>(b) * 0x85EBCA77U + 0xC2B2AE3DU)));
}
__host__ __device__ __forceinline__ uint32_t hash3(uint32_t seed, int32_t a, int32_t b, int32_t c) {
    return mix32(hash2(seed, a, b) + static_cast<uint32_t>(c) * 0x27D4EB2FU);
}

}  // namespace dsp
