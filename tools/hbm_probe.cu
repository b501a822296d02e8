// HBM ceiling probe (sm_100a): pure-write, pure-read and copy bandwidth with plain 128-bit accesses and
// with bulk (TMA) shared->global stores, to know what "100 %" means for a write-dominated kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hbm_probe tools/hbm_probe.cu && tools/hbm_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_write(uint4* dst, size_t n16) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    uint4 v = make_uint4(1, 2, 3, 4);
    for (; i < n16; i += stride) dst[i] = v;
}
__global__ void k_read(const uint4* src, size_t n16, uint32_t* sink) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (; i < n16; i += stride) { uint4 v = src[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void k_copy(const uint4* src, uint4* dst, size_t n16) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) dst[i] = src[i];
}
// each warp: fill a private smem slot, then one bulk store of SLOT bytes; loop
template <int SLOT>
__global__ void k_bulk_write(uint8_t* dst, size_t bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    uint8_t* slot = sm + warp * SLOT;
    const size_t gw = blockIdx.x * (size_t)nwarp + warp, tw = (size_t)gridDim.x * nwarp;
    for (size_t off = gw * SLOT; off + SLOT <= bytes; off += tw * SLOT) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int j = lane * 16; j < SLOT; j += 512) *reinterpret_cast<uint4*>(slot + j) = make_uint4(off, j, 3, 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(SLOT) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Mesh-like pattern: a CTA takes region tickets (REGION contiguous bytes each, regions follow each other in the
// buffer), optionally reads 8 KiB per region, and writes the region with per-warp 3840-byte bulk stores.
template <bool READ>
__global__ void k_region_write(uint8_t* dst, const uint4* src, size_t region, size_t nregions, unsigned* ticket, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ unsigned s_t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* slot = sm + warp * 3840;
    uint32_t acc = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
        __syncthreads();
        const size_t t = s_t;
        if (t >= nregions) break;
        if (READ) { const uint4* q = src + t * 512; uint4 v = q[threadIdx.x]; uint4 w = q[threadIdx.x + 256]; acc ^= v.x ^ w.y; }
        uint8_t* base = dst + t * region;
        for (size_t off = (size_t)warp * 3840; off + 3840 <= region; off += 8 * 3840) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            for (int j = lane * 16; j < 3840; j += 512) *reinterpret_cast<uint4*>(slot + j) = make_uint4(off, j, acc, 4);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(3840) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
float time_ms(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}

int main() {
    const size_t bytes = 1ull << 30, n16 = bytes / 16;
    uint4 *a, *b; uint32_t* sink;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int mult : {4, 8, 16, 32}) {
        const int grid = sms * mult;
        float w = time_ms([&] { k_write<<<grid, 256>>>(a, n16); });
        float r = time_ms([&] { k_read<<<grid, 256>>>(a, n16, sink); });
        float c = time_ms([&] { k_copy<<<grid, 256>>>(a, b, n16); });
        printf("grid %4d x256: write %7.1f GB/s   read %7.1f GB/s   copy %7.1f GB/s (r+w bytes)\n", grid, bytes / w / 1e6, bytes / r / 1e6, 2.0 * bytes / c / 1e6);
    }
    float ms_memset = time_ms([&] { cudaMemsetAsync(a, 0, bytes); });
    printf("cudaMemsetAsync: %7.1f GB/s\n", bytes / ms_memset / 1e6);
    cudaFuncSetAttribute(k_bulk_write<3840>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3840);
    cudaFuncSetAttribute(k_bulk_write<15360>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 15360);
    cudaFuncSetAttribute(k_bulk_write<30720>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 30720);
    for (int cps : {2, 3, 4, 6}) {
        float t1 = time_ms([&] { k_bulk_write<3840><<<sms * cps, 256, 8 * 3840>>>(reinterpret_cast<uint8_t*>(a), bytes); });
        printf("bulk store 3840 B/warp, 8 warps/CTA, %d CTA/SM: %7.1f GB/s\n", cps, bytes / t1 / 1e6);
    }
    for (int cps : {2, 3}) {
        float t2 = time_ms([&] { k_bulk_write<15360><<<sms * cps, 128, 4 * 15360>>>(reinterpret_cast<uint8_t*>(a), bytes); });
        printf("bulk store 15360 B/warp, 4 warps/CTA, %d CTA/SM: %7.1f GB/s\n", cps, bytes / t2 / 1e6);
        float t3 = time_ms([&] { k_bulk_write<30720><<<sms * cps, 64, 2 * 30720>>>(reinterpret_cast<uint8_t*>(a), bytes); });
        printf("bulk store 30720 B/warp, 2 warps/CTA, %d CTA/SM: %7.1f GB/s\n", cps, bytes / t3 / 1e6);
    }
    {
        unsigned* ticket; cudaMalloc(&ticket, 4);
        cudaFuncSetAttribute(k_region_write<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3840);
        cudaFuncSetAttribute(k_region_write<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3840);
        for (size_t region : {(size_t)30720 * 4, (size_t)30720 * 6, (size_t)30720 * 16}) {
            const size_t nreg = bytes / region;
            for (int cps : {3, 4, 6}) {
                float tw = time_ms([&] { cudaMemsetAsync(ticket, 0, 4); k_region_write<false><<<sms * cps, 256, 8 * 3840>>>(reinterpret_cast<uint8_t*>(a), b, region, nreg, ticket, sink); });
                float tr = time_ms([&] { cudaMemsetAsync(ticket, 0, 4); k_region_write<true><<<sms * cps, 256, 8 * 3840>>>(reinterpret_cast<uint8_t*>(a), b, region, nreg, ticket, sink); });
                printf("region %7zu B, ticketed, %d CTA/SM: write-only %7.1f GB/s   with 8 KiB read/region %7.1f GB/s (write bytes only)\n", region, cps,
                       nreg * region / tw / 1e6, nreg * region / tr / 1e6);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
