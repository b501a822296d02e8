// Handing the vertex arena to the consumer without a copy (the reference's consumer is a Vulkan vertex
// buffer: chunk_mesh.rs:111-124 `Buffer::from_iter(.. VERTEX_BUFFER ..)`, device selection with only khr_swapchain enabled
// at engine/rendering/vulkan.rs:34-37).  With DSP_CREATE_EXPORTABLE_ARENA the arena is a CUDA virtual-memory allocation
// (cuMemCreate, POSIX file-descriptor handle type) instead of a cudaMalloc block: dsp_arena_export_fd returns a file
// descriptor that another API or process imports -- Vulkan through VK_KHR_external_memory_fd
// (VkImportMemoryFdInfoKHR, handle type OPAQUE_FD), another CUDA process through cuMemImportFromShareableHandle; the
// importer sees the same HBM bytes the mesh kernels write.  dsp_arena_import_* is the CUDA-side importer (used by the tests
// and by CUDA consumers).  Driver entry points are fetched through cudaGetDriverEntryPoint, so the library still has no
// link-time dependency on libcuda (it loads, and reports DSP_ERR_NO_DEVICE, on machines without a driver).
#include <cuda.h>

namespace {

struct DriverApi {
    CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*memExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*memImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*getErrorString)(CUresult, const char**) = nullptr;
    bool ok = false;
};

const DriverApi& driver_api() {
    static DriverApi api = [] {
        DriverApi a;
        bool all = true;
        auto get = [&](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) { all = false; cudaGetLastError(); }
        };
        get("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&a.memGetAllocationGranularity));
        get("cuMemCreate", reinterpret_cast<void**>(&a.memCreate));
        get("cuMemRelease", reinterpret_cast<void**>(&a.memRelease));
        get("cuMemAddressReserve", reinterpret_cast<void**>(&a.memAddressReserve));
        get("cuMemAddressFree", reinterpret_cast<void**>(&a.memAddressFree));
        get("cuMemMap", reinterpret_cast<void**>(&a.memMap));
        get("cuMemUnmap", reinterpret_cast<void**>(&a.memUnmap));
        get("cuMemSetAccess", reinterpret_cast<void**>(&a.memSetAccess));
        get("cuMemExportToShareableHandle", reinterpret_cast<void**>(&a.memExportToShareableHandle));
        get("cuMemImportFromShareableHandle", reinterpret_cast<void**>(&a.memImportFromShareableHandle));
        get("cuGetErrorString", reinterpret_cast<void**>(&a.getErrorString));
        a.ok = all;
        return a;
    }();
    return api;
}

std::string cu_error(CUresult r) {
    const char* s = nullptr;
    if (driver_api().getErrorString) driver_api().getErrorString(r, &s);
    return s ? s : "unknown driver error";
}

CUmemAllocationProp arena_prop(int device) {
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    return prop;
}

// maps `handle` (size bytes) into this process' address space on `device`, read + write
int map_allocation(int device, CUmemGenericAllocationHandle handle, size_t size, CUdeviceptr* out, std::string* err) {
    const DriverApi& d = driver_api();
    CUdeviceptr ptr = 0;
    CUresult r = d.memAddressReserve(&ptr, size, 0, 0, 0);
    if (r != CUDA_SUCCESS) { *err = "cuMemAddressReserve: " + cu_error(r); return DSP_ERR_CUDA; }
    if ((r = d.memMap(ptr, size, 0, handle, 0)) != CUDA_SUCCESS) { d.memAddressFree(ptr, size); *err = "cuMemMap: " + cu_error(r); return DSP_ERR_CUDA; }
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if ((r = d.memSetAccess(ptr, size, &acc, 1)) != CUDA_SUCCESS) { d.memUnmap(ptr, size); d.memAddressFree(ptr, size); *err = "cuMemSetAccess: " + cu_error(r); return DSP_ERR_CUDA; }
    *out = ptr;
    return DSP_OK;
}

// the arena of an exportable context; on failure *err says why and nothing is left allocated
int alloc_exportable_arena(dsp_ctx* ctx, uint64_t bytes, std::string* err) {
    const DriverApi& d = driver_api();
    if (!d.ok) { *err = "the CUDA driver does not offer the virtual-memory entry points (cuMemCreate ...)"; return DSP_ERR_CUDA; }
    if (cudaFree(nullptr) != cudaSuccess) { *err = "no CUDA context"; return DSP_ERR_CUDA; }  // make sure the primary context is current
    const CUmemAllocationProp prop = arena_prop(ctx->device);
    size_t gran = 0;
    CUresult r = d.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || gran == 0) { *err = "cuMemGetAllocationGranularity: " + cu_error(r); return DSP_ERR_CUDA; }
    const size_t size = (std::max<uint64_t>(bytes, 128) + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h = 0;
    if ((r = d.memCreate(&h, size, &prop, 0)) != CUDA_SUCCESS) { *err = "cuMemCreate(" + std::to_string(size) + " bytes, POSIX fd handle type): " + cu_error(r); return DSP_ERR_CUDA; }
    CUdeviceptr ptr = 0;
    const int rc = map_allocation(ctx->device, h, size, &ptr, err);
    if (rc != DSP_OK) { d.memRelease(h); return rc; }
    ctx->d_arena = reinterpret_cast<uint8_t*>(ptr);
    ctx->arena_vmm_handle = h;
    ctx->arena_vmm_size = size;
    return DSP_OK;
}

void free_exportable_arena(dsp_ctx* ctx) {
    if (!ctx->arena_vmm_size) return;
    const DriverApi& d = driver_api();
    d.memUnmap(reinterpret_cast<CUdeviceptr>(ctx->d_arena), ctx->arena_vmm_size);
    d.memAddressFree(reinterpret_cast<CUdeviceptr>(ctx->d_arena), ctx->arena_vmm_size);
    d.memRelease(ctx->arena_vmm_handle);
    ctx->d_arena = nullptr;
    ctx->arena_vmm_size = 0;
}

}  // namespace

struct dsp_arena_import {
    int device = 0;
    CUmemGenericAllocationHandle handle = 0;
    CUdeviceptr ptr = 0;
    size_t size = 0;
};

extern "C" {

int dsp_arena_export_fd(dsp_ctx* ctx, int* out_fd, uint64_t* out_allocation_bytes) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (!out_fd) return fail(ctx, DSP_ERR_BAD_ARG, "out_fd is NULL");
    if (!ctx->arena_vmm_size) return fail(ctx, DSP_ERR_BAD_ARG, "this context's arena is not exportable: create it with dsp_create_ex(..., DSP_CREATE_EXPORTABLE_ARENA, ...)");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int fd = -1;
    const CUresult r = driver_api().memExportToShareableHandle(&fd, ctx->arena_vmm_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) return fail(ctx, DSP_ERR_CUDA, "cuMemExportToShareableHandle: %s", cu_error(r).c_str());
    *out_fd = fd;
    if (out_allocation_bytes) *out_allocation_bytes = ctx->arena_vmm_size;
    return DSP_OK;
}

int dsp_arena_import_fd(int device, int fd, uint64_t allocation_bytes, dsp_arena_import** out_import, void** out_device_ptr) {
    if (!out_import || !out_device_ptr || fd < 0 || allocation_bytes == 0) return fail(nullptr, DSP_ERR_BAD_ARG, "bad import arguments");
    *out_import = nullptr;
    *out_device_ptr = nullptr;
    const DriverApi& d = driver_api();
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DSP_ERR_NO_DEVICE, "cannot use CUDA device %d", device); }
    if (!d.ok) return fail(nullptr, DSP_ERR_CUDA, "the CUDA driver does not offer the virtual-memory entry points");
    CUmemGenericAllocationHandle h = 0;
    CUresult r = d.memImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) return fail(nullptr, DSP_ERR_CUDA, "cuMemImportFromShareableHandle: %s", cu_error(r).c_str());
    CUdeviceptr ptr = 0;
    std::string err;
    const int rc = map_allocation(device, h, allocation_bytes, &ptr, &err);
    if (rc != DSP_OK) { d.memRelease(h); return fail(nullptr, rc, "%s", err.c_str()); }
    dsp_arena_import* imp = new dsp_arena_import();
    imp->device = device; imp->handle = h; imp->ptr = ptr; imp->size = allocation_bytes;
    *out_import = imp;
    *out_device_ptr = reinterpret_cast<void*>(ptr);
    return DSP_OK;
}

int dsp_arena_import_read(dsp_arena_import* imp, uint64_t offset_bytes, uint64_t bytes, void* host_dst) {
    if (!imp || (bytes && !host_dst) || offset_bytes + bytes > imp->size) return fail(nullptr, DSP_ERR_BAD_ARG, "bad arguments to dsp_arena_import_read");
    if (bytes == 0) return DSP_OK;
    cudaError_t e = cudaSetDevice(imp->device);
    if (e == cudaSuccess) e = cudaMemcpy(host_dst, reinterpret_cast<const void*>(imp->ptr + offset_bytes), bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(nullptr, DSP_ERR_CUDA, "cudaMemcpy from the imported arena: %s", cudaGetErrorString(e));
    return DSP_OK;
}

int dsp_arena_import_close(dsp_arena_import* imp) {
    if (!imp) return DSP_OK;
    const DriverApi& d = driver_api();
    cudaSetDevice(imp->device);
    d.memUnmap(imp->ptr, imp->size);
    d.memAddressFree(imp->ptr, imp->size);
    d.memRelease(imp->handle);
    delete imp;
    return DSP_OK;
}

}  // extern "C"
