// Probe: how fast can SMs pull 8 KiB tiles straight out of pinned host memory with 1-D bulk (TMA) copies, against cudaMemcpyAsync
// of the same bytes?  (Would a mesh kernel that reads host tiles directly beat the copy-engine pipeline of dsp_mesh_host_chunks?)
//   nvcc -arch=sm_100a -O3 -o tools/pcie_pull_probe tools/pcie_pull_probe.cu && tools/pcie_pull_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../despawnengine_b200/csrc/dsp_common.cuh"
using namespace dsp;

template <int NBUF>
__global__ void __launch_bounds__(128) pull_kernel(const uint8_t* __restrict__ host, uint8_t* __restrict__ dev, uint32_t n_tiles, uint32_t* ticket) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[NBUF];
    __shared__ uint32_t s_t[NBUF];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
        for (int b = 0; b < NBUF; ++b) {
            const uint32_t t = atomicAdd(ticket, 1u);
            s_t[b] = t;
            if (t < n_tiles) { mbar_arrive_expect_tx(&bar[b], 8192); bulk_load(smem + b * 8192, host + (size_t)t * 8192, 8192, &bar[b]); }
        }
    }
    __syncthreads();
    if (tid != 0) return;
    for (uint32_t it = 0;; ++it) {
        const int b = it % NBUF;
        const uint32_t t = s_t[b];
        if (t >= n_tiles) break;
        mbar_wait(&bar[b], (it / NBUF) & 1);
        bulk_store(dev + (size_t)t * 8192, smem + b * 8192, 8192);
        bulk_commit();
        bulk_wait_read<0>();
        const uint32_t t2 = atomicAdd(ticket, 1u);
        s_t[b] = t2;
        if (t2 < n_tiles) { mbar_arrive_expect_tx(&bar[b], 8192); bulk_load(smem + b * 8192, host + (size_t)t2 * 8192, 8192, &bar[b]); }
    }
    bulk_wait<0>();
}

int main() {
    const size_t bytes = 32u << 20;
    const uint32_t n_tiles = bytes / 8192;
    uint8_t *h, *d; uint32_t* ticket;
    cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    for (size_t i = 0; i < bytes; ++i) h[i] = (uint8_t)(i * 2654435761u >> 24);
    cudaMalloc(&d, bytes); cudaMalloc(&ticket, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        for (int k = 0; k < 10; ++k) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("cudaMemcpyAsync H2D 32 MiB: %.3f ms  %.1f GB/s\n", ms / 10, bytes / (ms / 10) / 1e6);
    }
    auto run = [&](auto kern, int nbuf, int grid, const char* name) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nbuf * 8192);
        for (int rep = 0; rep < 3; ++rep) {
            float tot = 0;
            for (int k = 0; k < 10; ++k) {
                cudaMemsetAsync(ticket, 0, 4);
                cudaEventRecord(e0);
                kern<<<grid, 128, nbuf * 8192>>>(h, d, n_tiles, ticket);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); tot += ms;
            }
            printf("%s grid %d: %.3f ms  %.1f GB/s  (%s)\n", name, grid, tot / 10, bytes / (tot / 10) / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    };
    run(pull_kernel<2>, 2, 148, "pull 2 x 8 KiB per CTA");
    run(pull_kernel<2>, 2, 444, "pull 2 x 8 KiB per CTA");
    run(pull_kernel<4>, 4, 444, "pull 4 x 8 KiB per CTA");
    run(pull_kernel<4>, 4, 148 * 8, "pull 4 x 8 KiB per CTA");
    run(pull_kernel<2>, 2, 32, "pull 2 x 8 KiB per CTA");
    // verify
    uint8_t* back = (uint8_t*)malloc(bytes); cudaMemcpy(back, d, bytes, cudaMemcpyDeviceToHost);
    size_t bad = 0; for (size_t i = 0; i < bytes; ++i) bad += back[i] != h[i];
    printf("mismatches: %zu\n", bad);
    return 0;
}
