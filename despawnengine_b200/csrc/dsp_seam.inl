// Region seams between GPUs, inside the library (BASELINE.json "border-slab halos move between
// neighbouring regions by NVLink peer copies ... only at region seams").  Included at the end of dsp_api.cu.
//
// The reference is one process on one thread and its mesher never leaves the chunk (chunk_mesh.rs:46-51), so there is
// nothing there to mirror; the exchange exists for DSP_MESH_HALO (cross-chunk culling) when the world is cut into one
// region per GPU.  Design (B200 / NVSwitch: every GPU reaches every peer's memory at full NVLink bandwidth):
//   * each side of a seam owns an INBOX in its own HBM: {flag, ack, data[2][n][256] u16}; the neighbour maps it -- peer
//     access inside one process, CUDA IPC between the one-process-per-GPU ranks -- and is the only writer;
//   * dsp_seam_exchange_async = one small kernel per seam on the context's stream: it packs this GPU's border layers
//     (512 B per chunk face, not the 8 KiB chunk) and stores them STRAIGHT INTO THE NEIGHBOUR'S INBOX through the mapping,
//     then sets the inbox flag to the epoch with a system-scope release.  No host synchronisation, no staging buffer, no
//     copy on the receiving side, no collective library;
//   * the halo-mode mesh call that follows is still ONE launch: its pre-pass puts chunks that touch a seam at the END of
//     the work list, and a CTA that reaches one spins on the (local) inbox flag first.  By then the interior of the region
//     has been meshed, so the transfer (8 MiB per seam for the 512 x 512 x 32 world: ~10 us of NVLink) is hidden;
//   * two data buffers per inbox alternate by epoch and an ack word travels back with the next push, so a sender can
//     never overwrite layers its neighbour is still reading, and nobody waits for anybody in steady state.
// Exchanges are collective over a seam: both sides call dsp_seam_exchange_async once per epoch.

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is resolved at run time (dlopen), there is no link-time dependency

namespace {

// ---- NCCL transport (fallback where a neighbour's memory cannot be mapped: no P2P path, separate PID namespaces) ------------
// The product path above needs nothing but CUDA.  Where peer mapping is not available, a seam can be given a NCCL route instead:
// the same push kernel packs the border layers into a LOCAL send buffer, and one grouped ncclSend / ncclRecv per seam on the
// context's stream moves them into the inbox; a one-thread kernel then raises the inbox flag, so the mesh kernel's wait is the
// same as on the peer-memory route.  libnccl.so.2 is looked up with dlopen when dsp_comm_init is called (the copy torch has
// already loaded, else the system one); without it the call fails with a clear message and nothing else is affected.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};
const NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the one already in the process (torch's), if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { a.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return a; }
        bool all = true;
        auto get = [&](const char* name, void** fn) { *fn = dlsym(h, name); if (!*fn) { all = false; a.why = std::string("libnccl.so.2 lacks ") + name; } };
        get("ncclGetUniqueId", reinterpret_cast<void**>(&a.GetUniqueId));
        get("ncclCommInitRank", reinterpret_cast<void**>(&a.CommInitRank));
        get("ncclCommDestroy", reinterpret_cast<void**>(&a.CommDestroy));
        get("ncclGroupStart", reinterpret_cast<void**>(&a.GroupStart));
        get("ncclGroupEnd", reinterpret_cast<void**>(&a.GroupEnd));
        get("ncclSend", reinterpret_cast<void**>(&a.Send));
        get("ncclRecv", reinterpret_cast<void**>(&a.Recv));
        get("ncclGetErrorString", reinterpret_cast<void**>(&a.GetErrorString));
        a.ok = all;
        return a;
    }();
    return api;
}
#define DSP_NCCL(ctx, call)                                                                                                   \
    do {                                                                                                                      \
        ncclResult_t r_ = (call);                                                                                             \
        if (r_ != ncclSuccess) return fail(ctx, DSP_ERR_CUDA, "%s failed: %s", #call, nccl_api().GetErrorString(r_));         \
    } while (0)

constexpr uint32_t kSeamMagic = 0x4D535344u;  // "DSSM"
struct SeamWire {  // what dsp_seam_handle carries (128 bytes)
    uint32_t magic, version;
    int32_t pid, device;
    uint32_t n;
    int32_t face;
    uint64_t raw_ptr, inbox_bytes;
    cudaIpcMemHandle_t ipc;
    uint32_t ipc_ok;
    uint32_t reserved[5];
};
static_assert(sizeof(SeamWire) <= sizeof(dsp_seam_handle), "dsp_seam_handle too small");

int apply_seam_batches(dsp_ctx* ctx) {
    if (ctx->seam_epoch == 0) return DSP_OK;  // nothing has been received yet: the seam faces stay "no neighbour"
    for (size_t s = 0; s < ctx->seams.size(); ++s) {
        const dsp_ctx::Seam& sm = ctx->seams[s];
        if (!sm.connected || sm.n == 0) continue;
        dsp::apply_halo_batch_kernel<<<(sm.n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_nbr, sm.d_slots, sm.face, static_cast<uint32_t>(s) + 1u, 0u, sm.n);
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return DSP_OK;
}

void fill_seam_params(const dsp_ctx* ctx, dsp::MeshParams& p) {
    p.ext_base[0] = ctx->d_halo;
    p.n_seams = 0;
    p.seam_epoch = ctx->seam_epoch;
    if (ctx->seam_epoch == 0) return;
    uint32_t k = 0;
    for (size_t s = 0; s < ctx->seams.size(); ++s) {
        const dsp_ctx::Seam& sm = ctx->seams[s];
        if (!sm.connected) continue;
        p.ext_base[1 + s] = reinterpret_cast<const uint16_t*>(sm.inbox + dsp::kSeamInboxHeader) + static_cast<size_t>(ctx->seam_epoch & 1u) * sm.n * 256;
        p.seam_flag[k++] = reinterpret_cast<const uint32_t*>(sm.inbox);  // SeamInbox::flag
    }
    p.n_seams = k;
}

void destroy_seams(dsp_ctx* ctx) {
    if (ctx->nccl_comm) { nccl_api().CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm)); ctx->nccl_comm = nullptr; }
    for (auto& sm : ctx->seams) {
        if (sm.peer && sm.peer_ipc) cudaIpcCloseMemHandle(sm.peer);
        cudaFree(sm.outbox);
        cudaFree(sm.inbox);
        cudaFree(sm.d_slots);
        cudaFree(sm.d_done);
    }
    ctx->seams.clear();
}

}  // namespace

extern "C" {

int dsp_seam_create(dsp_ctx* ctx, int face, uint64_t n, const uint32_t* slots, uint32_t* out_seam, dsp_seam_handle* out_handle) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (face < 0 || face > 5) return fail(ctx, DSP_ERR_BAD_ARG, "face %d out of range", face);
    if (n == 0 || n > dsp::kNbrIndexMask || !slots || !out_seam || !out_handle) return fail(ctx, DSP_ERR_BAD_ARG, "bad seam arguments");
    if (ctx->seams.size() >= 7) return fail(ctx, DSP_ERR_BAD_ARG, "a context holds at most 7 seams");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = validate_slots(ctx, n, slots);
    if (rc != DSP_OK) return rc;
    dsp_ctx::Seam sm{};
    sm.face = face;
    sm.n = static_cast<uint32_t>(n);
    sm.slots.assign(slots, slots + n);
    sm.inbox_bytes = dsp::kSeamInboxHeader + 2 * n * 512;
    DSP_CUDA(ctx, cudaMalloc(&sm.inbox, sm.inbox_bytes));
    DSP_CUDA(ctx, cudaMemset(sm.inbox, 0, sm.inbox_bytes));
    DSP_CUDA(ctx, cudaMalloc(&sm.d_slots, n * sizeof(uint32_t)));
    DSP_CUDA(ctx, cudaMemcpy(sm.d_slots, slots, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    DSP_CUDA(ctx, cudaMalloc(&sm.d_done, sizeof(uint32_t)));
    DSP_CUDA(ctx, cudaMemset(sm.d_done, 0, sizeof(uint32_t)));
    SeamWire w{};
    w.magic = kSeamMagic; w.version = 1; w.pid = static_cast<int32_t>(getpid()); w.device = ctx->device;
    w.n = sm.n; w.face = face; w.raw_ptr = reinterpret_cast<uint64_t>(sm.inbox); w.inbox_bytes = sm.inbox_bytes;
    w.ipc_ok = cudaIpcGetMemHandle(&w.ipc, sm.inbox) == cudaSuccess ? 1u : 0u;  // (not fatal: a same-process peer needs no IPC)
    if (!w.ipc_ok) cudaGetLastError();
    std::memset(out_handle, 0, sizeof *out_handle);
    std::memcpy(out_handle, &w, sizeof w);
    ctx->seam_slots.insert(slots, slots + n);
    *out_seam = static_cast<uint32_t>(ctx->seams.size());
    ctx->seams.push_back(std::move(sm));
    return DSP_OK;
}

int dsp_seam_connect(dsp_ctx* ctx, uint32_t seam, const dsp_seam_handle* peer_handle) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (seam >= ctx->seams.size() || !peer_handle) return fail(ctx, DSP_ERR_BAD_ARG, "no such seam");
    dsp_ctx::Seam& sm = ctx->seams[seam];
    if (sm.connected) return fail(ctx, DSP_ERR_BAD_ARG, "seam %u is connected already", seam);
    SeamWire w;
    std::memcpy(&w, peer_handle, sizeof w);
    if (w.magic != kSeamMagic || w.version != 1) return fail(ctx, DSP_ERR_BAD_ARG, "not a dsp_seam_handle");
    if (w.n != sm.n) return fail(ctx, DSP_ERR_BAD_ARG, "the two sides of a seam list different numbers of chunks (%u here, %u there)", sm.n, w.n);
    if (w.face != (sm.face ^ 1)) return fail(ctx, DSP_ERR_BAD_ARG, "seam faces do not oppose each other (%d here, %d there)", sm.face, w.face);
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (w.pid == static_cast<int32_t>(getpid())) {
        if (w.device != ctx->device) {  // two contexts of one process on two GPUs: plain peer access
            int can = 0;
            DSP_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, w.device));
            if (!can) return fail(ctx, DSP_ERR_CUDA, "device %d cannot access device %d (no NVLink / P2P path)", ctx->device, w.device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(w.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, DSP_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        sm.peer = reinterpret_cast<uint8_t*>(w.raw_ptr);
        sm.peer_ipc = false;
    } else {
        if (!w.ipc_ok) return fail(ctx, DSP_ERR_CUDA, "the peer could not export its inbox (cudaIpcGetMemHandle failed there)");
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, w.ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, DSP_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
        sm.peer = static_cast<uint8_t*>(ptr);
        sm.peer_ipc = true;
    }
    sm.connected = true;
    return DSP_OK;
}

int dsp_seam_exchange_async(dsp_ctx* ctx) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (ctx->seams.empty()) return DSP_OK;
    for (size_t s = 0; s < ctx->seams.size(); ++s)
        if (!ctx->seams[s].connected) return fail(ctx, DSP_ERR_BAD_ARG, "seam %zu has no peer yet (dsp_seam_connect / dsp_seam_connect_nccl)", s);
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t epoch = ++ctx->seam_epoch;
    if (epoch == 1) ctx->nbr_dirty = true;  // from now on the seam faces have a neighbour
    bool any_nccl = false;
    for (auto& sm : ctx->seams) {
        dsp::SeamPushParams p{};
        p.voxels = ctx->d_voxels;
        p.slots = sm.d_slots;
        p.n = sm.n;
        p.face = sm.face;
        p.epoch = epoch;
        // peer-memory route: straight into the neighbour's inbox; NCCL route: into our own outbox (same layout), shipped below
        p.peer = reinterpret_cast<dsp::SeamInbox*>(sm.nccl_peer >= 0 ? sm.outbox : sm.peer);
        p.mine = reinterpret_cast<const dsp::SeamInbox*>(sm.nccl_peer >= 0 ? sm.outbox : sm.inbox);  // (NCCL is stream-ordered: no ack to wait for)
        if (sm.nccl_peer >= 0) p.epoch_for_ack_wait_off = 1u;
        p.done = sm.d_done;
        p.status = reinterpret_cast<uint32_t*>(ctx->d_ctl) + 1;
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(sm.n, static_cast<uint64_t>(ctx->num_sms) * 8));
        dsp::seam_push_kernel<<<grid, 256, 0, ctx->stream>>>(p);
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        any_nccl = any_nccl || sm.nccl_peer >= 0;
    }
    if (any_nccl) {
        const NcclApi& nc = nccl_api();
        ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
        DSP_NCCL(ctx, nc.GroupStart());
        for (auto& sm : ctx->seams) {
            if (sm.nccl_peer < 0) continue;
            const size_t bytes = static_cast<size_t>(sm.n) * 512, off = dsp::kSeamInboxHeader + static_cast<size_t>(epoch & 1u) * bytes;
            DSP_NCCL(ctx, nc.Send(sm.outbox + off, bytes, ncclUint8, sm.nccl_peer, comm, ctx->stream));
            DSP_NCCL(ctx, nc.Recv(sm.inbox + off, bytes, ncclUint8, sm.nccl_peer, comm, ctx->stream));
        }
        DSP_NCCL(ctx, nc.GroupEnd());
        for (auto& sm : ctx->seams) {
            if (sm.nccl_peer < 0) continue;
            dsp::seam_flag_kernel<<<1, 1, 0, ctx->stream>>>(reinterpret_cast<dsp::SeamInbox*>(sm.inbox), epoch);
            DSP_CUDA(ctx, cudaGetLastError());
            ctx->launches++;
        }
    }
    return DSP_OK;
}

int dsp_comm_unique_id(uint8_t* out_id128) {
    if (!out_id128) return fail(nullptr, DSP_ERR_BAD_ARG, "out_id128 is NULL");
    const NcclApi& nc = nccl_api();
    if (!nc.ok) return fail(nullptr, DSP_ERR_CUDA, "NCCL is not available: %s", nc.why.c_str());
    ncclUniqueId id;
    DSP_NCCL(nullptr, nc.GetUniqueId(&id));
    static_assert(sizeof id == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out_id128, &id, sizeof id);
    return DSP_OK;
}

int dsp_comm_init(dsp_ctx* ctx, int rank, int n_ranks, const uint8_t* id128) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (!id128 || rank < 0 || rank >= n_ranks) return fail(ctx, DSP_ERR_BAD_ARG, "bad communicator arguments");
    if (ctx->nccl_comm) return fail(ctx, DSP_ERR_BAD_ARG, "the context has a communicator already");
    const NcclApi& nc = nccl_api();
    if (!nc.ok) return fail(ctx, DSP_ERR_CUDA, "NCCL is not available: %s", nc.why.c_str());
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    DSP_NCCL(ctx, nc.CommInitRank(&comm, n_ranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->nccl_rank = rank;
    return DSP_OK;
}

int dsp_seam_connect_nccl(dsp_ctx* ctx, uint32_t seam, int peer_rank) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (seam >= ctx->seams.size()) return fail(ctx, DSP_ERR_BAD_ARG, "no such seam");
    if (!ctx->nccl_comm) return fail(ctx, DSP_ERR_BAD_ARG, "dsp_comm_init has not been called on this context");
    dsp_ctx::Seam& sm = ctx->seams[seam];
    if (sm.connected) return fail(ctx, DSP_ERR_BAD_ARG, "seam %u is connected already", seam);
    if (peer_rank < 0 || peer_rank == ctx->nccl_rank) return fail(ctx, DSP_ERR_BAD_ARG, "bad peer rank %d", peer_rank);
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    DSP_CUDA(ctx, cudaMalloc(&sm.outbox, sm.inbox_bytes));
    DSP_CUDA(ctx, cudaMemset(sm.outbox, 0, sm.inbox_bytes));
    sm.nccl_peer = peer_rank;
    sm.connected = true;
    return DSP_OK;
}

uint32_t dsp_seam_epoch(const dsp_ctx* ctx) { return ctx ? ctx->seam_epoch : 0u; }

}  // extern "C"
