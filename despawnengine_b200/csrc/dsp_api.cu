// libdespawn_b200.so -- C ABI (include/despawn_b200.h) over the sm_100a kernels.
//
// Host side of the context: the chunk map that World keeps (world.rs:10-14: pos -> chunk) becomes
// pos -> slot of a device-resident voxel store; everything that touches voxels or vertices runs on
// the GPU.  There is deliberately no CPU implementation of any operation in this file.
#include "../../include/despawn_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <map>
#include <vector>
#include <unistd.h>
#include <sched.h>
#include <thread>

#include "mesh_kernels.cuh"
#include "quad_kernels.cuh"
#include "terrain_kernels.cuh"
#include "host_pack.hpp"

namespace {

struct PosKey {
    int32_t x, y, z;
    bool operator==(const PosKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct PosHash {
    size_t operator()(const PosKey& k) const {
        uint64_t h = static_cast<uint32_t>(k.x) * 0x9E3779B97F4A7C15ull;
        h ^= (static_cast<uint64_t>(static_cast<uint32_t>(k.y)) + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        h ^= (static_cast<uint64_t>(static_cast<uint32_t>(k.z)) + 0x165667B1ull) * 0x9FB21C651E98DF25ull;
        h ^= h >> 29;
        return static_cast<size_t>(h * 0xBF58476D1CE4E5B9ull);
    }
};
struct Region {
    int32_t origin[3];
    uint32_t dims[3];
    uint32_t first_slot;
};

thread_local std::string g_create_error = "";

}  // namespace

struct dsp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of dsp_mesh_host_chunks, overlapped with meshing on `stream`
    cudaStream_t copy_stream2 = nullptr; // second side stream: the unpack kernels of a persistent-launch call alternate between the two
    cudaStream_t d2h_stream = nullptr;   // vertices -> host of dsp_mesh_host_chunks_to_host (created on first use)
    cudaEvent_t ev_meshed[64] = {};
    static constexpr int kMaxSub = 64;   // sub-batches of one pipelined call
    cudaEvent_t ev_sub[kMaxSub] = {};
    cudaEvent_t ev_done = nullptr;
    uint32_t* d_tickets = nullptr;       // kMaxSub ticket counters
    uint64_t max_chunks = 0, arena_bytes = 0;
    int num_sms = 0;
    int mesh_variant = 0;
    bool mesh_variant_forced = false;  // DSP_MESH_VARIANT was given
    unsigned mesh_prefetch_mult = 2;  // L2 prefetch distance in units of the grid size (0 = off)
    int mesh_ctas_per_sm[4][4] = {};  // [variant][halo], filled on first launch
    bool mesh_live_ready[4] = {};   // the LIVE instantiation of a kernel shape has its shared-memory attribute set
    int quad_ctas_per_sm[8] = {};     // [halo + 2 greedy + 4 indexed]

    uint16_t* d_voxels = nullptr;
    int4* d_pos = nullptr;
    uint32_t* d_nbr = nullptr;  // [max_chunks][6], allocated on first halo mesh
    uint16_t* d_halo = nullptr; // [halo_cap][256] border slabs of neighbours that live on other devices
    uint64_t halo_cap = 0, halo_n = 0;
    // received border slabs, one batch per dsp_set_halo_slabs call: slab first_index + i stands in for the neighbour of
    // slots[i] across `face` wherever no resident neighbour exists (applied to the neighbour table on the device)
    struct HaloBatch { int face; uint32_t first_index; uint32_t n; std::vector<uint32_t> slots; };  // device copy of slots: d_halo_slots + first_index
    uint32_t* d_halo_slots = nullptr;  // [halo_cap], parallel to d_halo
    std::vector<HaloBatch> halo_batches;
    // region seams (dsp_seam.inl): inboxes in this GPU's memory that neighbour GPUs write into
    struct Seam {
        int face = 0; uint32_t n = 0;
        std::vector<uint32_t> slots; uint32_t* d_slots = nullptr;   // this context's boundary chunks at the seam
        uint8_t* inbox = nullptr; size_t inbox_bytes = 0;           // ours: {flag, ack, data[2][n][256]}
        uint8_t* peer = nullptr; bool peer_ipc = false, connected = false;  // the neighbour's inbox, mapped
        uint32_t* d_done = nullptr;
        int nccl_peer = -1; uint8_t* outbox = nullptr;                       // NCCL route (fallback): peer rank and the local send buffer
    };
    void* nccl_comm = nullptr;  // ncclComm_t of dsp_comm_init
    int nccl_rank = -1;
    std::vector<Seam> seams;
    std::unordered_set<uint32_t> seam_slots;  // every slot listed by a seam (they may not be unloaded)
    uint32_t seam_epoch = 0;
    uint8_t* d_arena = nullptr;
    unsigned long long arena_vmm_handle = 0;  // DSP_CREATE_EXPORTABLE_ARENA: CUmemGenericAllocationHandle of the arena (dsp_export.inl)
    size_t arena_vmm_size = 0;                // ... and its size rounded to the allocation granularity (0: plain cudaMalloc arena)
    dsp::MeshEntry* d_entries = nullptr;
    uint8_t* d_ctl = nullptr;  // [ctrl 16 B][total 16 B][face counts of the DSP_MESH_ORDERED count pass: 8 B x max_chunks reserved, u32 used][result 16 B]
    uint32_t* d_slots = nullptr;
    uint8_t* d_summary = nullptr;     // per slot: dsp::kSummary* (all air / uniformly solid / unknown)
    uint16_t* d_uid = nullptr;        // per slot, valid while the summary says "no air": the one block id of the chunk, or kUidMixed
    uint16_t* d_bmask = nullptr;      // per slot: 96 border occupancy masks (6 layers x 16), what cross-chunk culling reads of a neighbour
    uint8_t* d_bvalid = nullptr;      // per slot: d_bmask up to date (bit 0) / queued for a refresh (bit 1)
    bool bmask_stale = false;         // some resident chunk's masks may be out of date: halo-mode calls check (ensure_border_masks)
    uint32_t* d_stale_list = nullptr; // [max_chunks] chunks whose border masks are being refreshed + [1] their count (allocated on first use)
    uint32_t* d_work_slots = nullptr; // device-built work list of halo-mode mesh calls (allocated on first use)
    uint32_t* d_work_entry = nullptr;
    float4* d_uv_rows = nullptr;
    float4* d_uvmap = nullptr;
    uint32_t n_blocks = 0, uv_cap = 0;
    uint8_t* d_scratch = nullptr;
    // packed upload of host-resident chunks (csrc/host_pack.hpp): worker threads, pinned staging (8 KiB per chunk of the largest call so far)
    dsp_host::PackPool* pack_pool = nullptr;
    uint8_t* h_pack = nullptr;
    size_t h_pack_bytes = 0;
    int e2e_pack = 1;         // DSP_E2E_PACK=0: host-chunk calls never pack (kernel pull / copy-engine pipeline as before)
    int host_threads = -1;    // DSP_HOST_THREADS: worker threads of the packer (-1: the CPUs this thread may run on, minus one; at most 15)
    uint64_t upload_stats[8] = {};  // dsp_host_upload_stats
    int e2e_pack_keep = 0;    // DSP_E2E_PACK_KEEP=1..16: of every 16 chunks, how many the host packs (0: from the number of workers)
    int e2e_pack_pieces = 3;  // DSP_E2E_PACK_PIECES: launches per call (pack / unpack + mesh pipeline)
    int e2e_pack_persist = 1; // parity-mode calls use ONE mesh launch that takes chunks as the unpack kernels release them (DSP_E2E_PACK_PERSIST=0: a mesh launch per piece, as the other modes)
    int e2e_pack_persist_pieces = 6;  // DSP_E2E_PACK_PERSIST_PIECES: unpack launches of such a call
    int quad_ctas_cap = 0;  // DSP_QUAD_CTAS: upper bound on the quad kernels' CTAs per SM (0: whatever fits)
    unsigned long long* d_phase = nullptr;  // debug phase timers (16 counters)
    size_t scratch_bytes = 0;
    uint8_t* h_pinned = nullptr;  // 4 KiB of pinned host memory for small read-backs
    uint8_t* h_stage = nullptr;   // pinned staging for positions / slot lists (async H2D without a pageable bounce)
    size_t h_stage_bytes = 0;
    cudaEvent_t ev_stage = nullptr;  // recorded after the last copy out of h_stage
    bool stage_pending = false;
    // CUDA events around every mesh kernel launch (ring of kEvRing pairs, drained lazily) so that the
    // kernel's own device time can be reported separately from the memset / copies around it.
    static constexpr int kEvRing = 16;
    cudaEvent_t ev_k0[kEvRing] = {}, ev_k1[kEvRing] = {};
    bool ev_pending[kEvRing] = {};
    int ev_next = 0;
    double kernel_ms_total = 0.0;
    float kernel_ms_last = 0.0f;
    uint64_t kernel_count = 0;

    std::unordered_map<PosKey, uint32_t, PosHash> map;
    std::vector<Region> regions;
    std::vector<uint32_t> free_slots;
    std::vector<PosKey> slot_pos;  // position of every slot handed out through the map (not region slots)
    std::vector<uint8_t> resident;
    uint64_t next_unused = 0, n_resident = 0;
    std::vector<uint32_t> slot_stamp;  // per-call duplicate detection (stamp_slot)
    uint32_t last_acquired = 0xFFFFFFFEu;  // acquire_slot's previous answer (its successor is tried first), its position and its box region (-1: a map slot)
    int last_region = -1;
    uint32_t last_local[3] = {0, 0, 0};  // ... its coordinates inside that box
    uint32_t stamp = 0;
    bool nbr_dirty = true;
    bool region_holes = false;  // a slot inside a box region has been unloaded (device-side edit addressing needs full boxes)

    uint64_t last_mesh_n = 0;
    uint64_t launches = 0;
    uint64_t e2e_pull_first_div = 5; // kernel pull: the first launch takes n / this chunks (DSP_E2E_PULL_DIV, for A/B runs)
    int e2e_pull_pieces = 2;         // (DSP_E2E_PULL_PIECES=3: a smaller launch in front, all on one stream; 1: two launches on one stream -- for A/B runs)
    int e2e_pull = 1;            // host-chunk calls on pinned memory: the mesh kernel pulls the tiles itself (DSP_E2E_PULL=0: copy-engine pipeline, for A/B runs)
    uint64_t e2e_head_chunks = 128;  // first sub-batch of the pipelined host-chunk calls (DSP_E2E_HEAD=0 switches it off for A/B runs)
    int summary_skip = 3;       // parity mesh calls use the chunk summaries: bit 0 all-air chunks end at the ticket, bit 1 chunks of one known solid id skip their tile (DSP_SUMMARY_SKIP=0..3 for A/B runs)
    bool use_pdl = true;        // programmatic dependent launch of the mesh kernels (DSP_PDL=0 switches it off for A/B runs)
    bool kernel_timing = true;  // record an event pair around every mesh launch (dsp_mesh_kernel_stats)
    bool timing_open = false;   // the start event of the current call's pair has been recorded
    bool ctl_clean = false;    // the control block is all zero (left so by a self-resetting launch): no memset needed
    bool result_in_tail = false;  // the last mesh call's {status, total} are in the result words, not in the control block
    uint64_t arena_limit = 0;  // when non-zero, a mesh call may only write below this arena offset (scene allocator)
    std::string err;

    // ---- scene state (dsp_scene_update): World.loaded_chunks + GameScene.chunk_meshes + last_chunk_pos
    struct SceneChunk { uint32_t loaded; uint32_t batch; dsp_mesh_entry entry; };
    struct SceneBatch { uint64_t offset, bytes; uint32_t live; };
    std::vector<SceneChunk> scene_chunks;  // by voxel slot (position: slot_pos): the loaded set is the previous update's visible set, so no map by position is needed
    uint64_t scene_n_loaded = 0;
    uint32_t scene_last_h = 0, scene_last_v = 0;  // render distances of the update that produced the loaded set
    std::vector<SceneBatch> scene_batches;
    std::vector<uint32_t> scene_free_batch_ids;
    std::map<uint64_t, uint64_t> scene_free;  // free arena intervals: offset -> bytes
    bool scene_init = false, scene_has_last = false;
    int32_t scene_last[3] = {0, 0, 0};
    uint64_t scene_meshes = 0, scene_vertices = 0, scene_indices = 0;
};

namespace {

int fail(dsp_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define DSP_CUDA(ctx, call)                                                                                         \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) return fail(ctx, DSP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int ensure_scratch(dsp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return DSP_OK;
    if (ctx->d_scratch) { DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); DSP_CUDA(ctx, cudaFree(ctx->d_scratch)); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0; }
    size_t cap = std::max<size_t>(bytes, 1u << 20);
    DSP_CUDA(ctx, cudaMalloc(&ctx->d_scratch, cap));
    ctx->scratch_bytes = cap;
    return DSP_OK;
}

// Pinned staging: returns `bytes` of pinned host memory whose previous contents have been consumed.
int stage_host(dsp_ctx* ctx, size_t bytes, uint8_t** out) {
    if (ctx->stage_pending) { DSP_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage)); ctx->stage_pending = false; }
    if (bytes > ctx->h_stage_bytes) {
        if (ctx->h_stage) DSP_CUDA(ctx, cudaFreeHost(ctx->h_stage));
        ctx->h_stage = nullptr; ctx->h_stage_bytes = 0;
        const size_t cap = std::max<size_t>(bytes + bytes / 2, 1u << 20);
        DSP_CUDA(ctx, cudaMallocHost(&ctx->h_stage, cap));
        ctx->h_stage_bytes = cap;
    }
    *out = ctx->h_stage;
    return DSP_OK;
}
int stage_release(dsp_ctx* ctx) {
    DSP_CUDA(ctx, cudaEventRecord(ctx->ev_stage, ctx->stream));
    ctx->stage_pending = true;
    return DSP_OK;
}

bool find_slot(const dsp_ctx* ctx, const int32_t p[3], uint32_t* out, int* region_index = nullptr) {
    if (region_index) *region_index = -1;
    auto it = ctx->map.find(PosKey{p[0], p[1], p[2]});
    if (it != ctx->map.end()) { *out = it->second; return true; }
    for (size_t k = 0; k < ctx->regions.size(); ++k) {
        const Region& r = ctx->regions[k];
        const int64_t a = static_cast<int64_t>(p[0]) - r.origin[0], b = static_cast<int64_t>(p[1]) - r.origin[1],
                      c = static_cast<int64_t>(p[2]) - r.origin[2];
        if (a < 0 || b < 0 || c < 0 || a >= r.dims[0] || b >= r.dims[1] || c >= r.dims[2]) continue;
        const uint64_t s = r.first_slot + static_cast<uint64_t>(a) + r.dims[0] * (static_cast<uint64_t>(b) + static_cast<uint64_t>(r.dims[1]) * c);
        if (ctx->resident[s]) {
            *out = static_cast<uint32_t>(s);
            if (region_index) *region_index = static_cast<int>(k);
            return true;
        }
    }
    return false;
}

// Does the box hold a position that is resident already (in another box region or in the position map)?
bool region_overlaps(const dsp_ctx* ctx, const int32_t origin[3], const uint32_t dims[3]) {
    for (const Region& r : ctx->regions) {
        bool apart = false;
        for (int a = 0; a < 3; ++a) {
            const int64_t lo0 = origin[a], hi0 = lo0 + dims[a], lo1 = r.origin[a], hi1 = lo1 + r.dims[a];
            if (hi0 <= lo1 || hi1 <= lo0) apart = true;
        }
        if (!apart) return true;  // (slots unloaded from inside a box keep their address, so the whole box stays reserved)
    }
    const uint64_t n = static_cast<uint64_t>(dims[0]) * dims[1] * dims[2];
    auto inside = [&](const PosKey& k) {
        const int64_t a = static_cast<int64_t>(k.x) - origin[0], b = static_cast<int64_t>(k.y) - origin[1], c = static_cast<int64_t>(k.z) - origin[2];
        return a >= 0 && b >= 0 && c >= 0 && a < dims[0] && b < dims[1] && c < dims[2];
    };
    if (ctx->map.size() <= n) {
        for (const auto& kv : ctx->map) if (inside(kv.first)) return true;
    } else {
        for (uint32_t c = 0; c < dims[2]; ++c)
            for (uint32_t b = 0; b < dims[1]; ++b)
                for (uint32_t a = 0; a < dims[0]; ++a)
                    if (ctx->map.count(PosKey{origin[0] + static_cast<int32_t>(a), origin[1] + static_cast<int32_t>(b), origin[2] + static_cast<int32_t>(c)})) return true;
    }
    return false;
}

bool slot_in_box_region(const dsp_ctx* ctx, uint64_t s) {
    for (const Region& r : ctx->regions) {
        const uint64_t cnt = static_cast<uint64_t>(r.dims[0]) * r.dims[1] * r.dims[2];
        if (s >= r.first_slot && s < r.first_slot + cnt) return true;
    }
    return false;
}

// World::get_chunk's "insert if absent" (world.rs:29-31, 44-46) on slots.  *fresh (may be NULL) tells whether the slot
// was created by this call (the chunk still has to be generated / uploaded) or was resident already.
int acquire_slot(dsp_ctx* ctx, const int32_t p[3], uint32_t* out, bool* fresh = nullptr) {
    if (fresh) *fresh = false;
    // A caller that hands over the chunks it first loaded together, in the same order, again (a region remeshed every frame) walks
    // the slots in ascending order, so "the slot after the previous answer" is tried before the hash map and the box list are asked:
    //   * previous answer in a box region (dsp_generate_region: slots are arithmetic, x fastest): the position of the next slot follows
    //     from the box coordinates of the previous one.  A resident box slot has never been unloaded (box slots are not recycled), so its position cannot
    //     have entered the map since -- the answer is the one find_slot would give;
    //   * previous answer from the map: the position recorded for the next slot (slot_pos is current for every resident slot handed
    //     out through the map; box slots are never recorded there) is compared -- a vector read in place of a node walk.
    // The position resolve sits on the calling thread between the pieces of the host-chunk calls: 14-20 ns per chunk through
    // find_slot, 60-80 of the packed upload's 210 us per 4,096 chunks.
    {
        const uint64_t g = static_cast<uint64_t>(ctx->last_acquired) + 1;
        bool hit = false;
        if (ctx->last_region >= 0) {
            const Region& r = ctx->regions[static_cast<size_t>(ctx->last_region)];
            uint32_t a = ctx->last_local[0] + 1, b = ctx->last_local[1], c = ctx->last_local[2];  // box coordinates of slot g
            if (a == r.dims[0]) { a = 0; if (++b == r.dims[1]) { b = 0; ++c; } }
            hit = c < r.dims[2] && static_cast<int64_t>(p[0]) == static_cast<int64_t>(r.origin[0]) + a && static_cast<int64_t>(p[1]) == static_cast<int64_t>(r.origin[1]) + b &&
                  static_cast<int64_t>(p[2]) == static_cast<int64_t>(r.origin[2]) + c && ctx->resident[g];
            if (hit) { ctx->last_local[0] = a; ctx->last_local[1] = b; ctx->last_local[2] = c; }
        } else {
            hit = g < ctx->slot_pos.size() && ctx->resident[g] && ctx->slot_pos[g] == PosKey{p[0], p[1], p[2]} && !slot_in_box_region(ctx, g);
        }
        if (hit) {
            ctx->last_acquired = static_cast<uint32_t>(g);
            *out = static_cast<uint32_t>(g);
            return DSP_OK;
        }
    }
    if (find_slot(ctx, p, out, &ctx->last_region)) {
        ctx->last_acquired = *out;
        if (ctx->last_region >= 0) {
            const Region& r = ctx->regions[static_cast<size_t>(ctx->last_region)];
            for (int k = 0; k < 3; ++k) ctx->last_local[k] = static_cast<uint32_t>(static_cast<int64_t>(p[k]) - r.origin[k]);
        }
        return DSP_OK;
    }
    if (fresh) *fresh = true;
    uint32_t s;
    if (!ctx->free_slots.empty()) { s = ctx->free_slots.back(); ctx->free_slots.pop_back(); }
    else if (ctx->next_unused < ctx->max_chunks) { s = static_cast<uint32_t>(ctx->next_unused++); }
    else return fail(ctx, DSP_ERR_STORE_FULL, "voxel store full: %llu chunks resident of %llu",
                     (unsigned long long)ctx->n_resident, (unsigned long long)ctx->max_chunks);
    ctx->map.emplace(PosKey{p[0], p[1], p[2]}, s);
    if (ctx->slot_pos.size() <= s) ctx->slot_pos.resize(static_cast<size_t>(s) + 1);
    ctx->slot_pos[s] = PosKey{p[0], p[1], p[2]};
    ctx->resident[s] = 1;
    ctx->n_resident++;
    ctx->nbr_dirty = true;
    ctx->last_acquired = s;
    *out = s;
    return DSP_OK;
}

// Undo of acquire_slot for the slots a failing call created (nothing of theirs is in flight: the caller synchronises).
void release_fresh_slots(dsp_ctx* ctx, const std::vector<uint32_t>& fresh) {
    for (uint32_t s : fresh) {
        if (s >= ctx->resident.size() || !ctx->resident[s]) continue;
        ctx->map.erase(ctx->slot_pos[s]);
        ctx->free_slots.push_back(s);
        ctx->resident[s] = 0;
        ctx->n_resident--;
    }
    if (!fresh.empty()) ctx->nbr_dirty = true;
}

// Per-call duplicate detection: a slot may appear once in the batch of an upload call (two copies into one slot race,
// and the palette remap would run twice over it).  Returns false when `slot` was already stamped by this call.
bool stamp_slot(dsp_ctx* ctx, uint32_t slot) {
    if (ctx->slot_stamp.size() <= slot) ctx->slot_stamp.resize(std::max<size_t>(static_cast<size_t>(slot) + 1, ctx->slot_stamp.size() * 2), 0u);
    if (ctx->slot_stamp[slot] == ctx->stamp) return false;
    ctx->slot_stamp[slot] = ctx->stamp;
    return true;
}
void new_stamp(dsp_ctx* ctx) {
    if (++ctx->stamp == 0u) { std::fill(ctx->slot_stamp.begin(), ctx->slot_stamp.end(), 0u); ctx->stamp = 1u; }
}

// Entries [i0, i1) of an upload batch: position -> slot (acquire_slot), one chunk per position (stamp_slot), the int4 position for the
// device.  Behind every answer that was resident already, the entries that simply CONTINUE it -- the next positions in box order
// inside a box region, or the positions recorded for the following map slots -- are taken in a tight loop over local variables:
// a batch handed over again in the order it was first loaded costs ~2 ns per chunk instead of 8-16 (this loop runs on the calling
// thread between the pieces of the host-chunk calls, with the GPU waiting behind it).  Anything else -- a fresh position, a jump,
// a hole, a duplicate -- leaves the loop and is decided by acquire_slot / stamp_slot as before.
int resolve_positions(dsp_ctx* ctx, const int32_t* pos, uint64_t i0, uint64_t i1, uint32_t* h_slots, int4* h_pos, std::vector<uint32_t>& fresh_slots) {
    uint64_t i = i0;
    while (i < i1) {
        bool fresh = false;
        const int32_t* p0 = pos + 3 * i;
        int rc = acquire_slot(ctx, p0, &h_slots[i], &fresh);
        if (rc != DSP_OK) return rc;
        if (fresh) fresh_slots.push_back(h_slots[i]);
        if (!stamp_slot(ctx, h_slots[i]))
            return fail(ctx, DSP_ERR_BAD_ARG, "position (%d, %d, %d) appears twice in the batch (entry %llu): one chunk per position", p0[0], p0[1], p0[2], (unsigned long long)i);
        h_pos[i] = make_int4(p0[0], p0[1], p0[2], 0);
        ++i;
        if (fresh || i == i1) continue;
        if (ctx->slot_stamp.size() < ctx->next_unused) ctx->slot_stamp.resize(ctx->next_unused, 0u);
        const uint8_t* const res = ctx->resident.data();
        uint32_t* const stamps = ctx->slot_stamp.data();
        const uint32_t cur = ctx->stamp;
        uint32_t s = h_slots[i - 1];
        if (ctx->last_region >= 0) {
            const Region r = ctx->regions[static_cast<size_t>(ctx->last_region)];
            uint32_t a = ctx->last_local[0], b = ctx->last_local[1], c = ctx->last_local[2];
            for (; i < i1; ++i) {
                uint32_t na = a + 1, nb = b, nc = c;
                if (na == r.dims[0]) { na = 0; if (++nb == r.dims[1]) { nb = 0; ++nc; } }
                if (nc >= r.dims[2]) break;
                const int32_t* p = pos + 3 * i;
                if (static_cast<int64_t>(p[0]) != static_cast<int64_t>(r.origin[0]) + na || static_cast<int64_t>(p[1]) != static_cast<int64_t>(r.origin[1]) + nb ||
                    static_cast<int64_t>(p[2]) != static_cast<int64_t>(r.origin[2]) + nc || !res[s + 1] || stamps[s + 1] == cur) break;
                ++s; stamps[s] = cur; a = na; b = nb; c = nc;
                h_slots[i] = s;
                h_pos[i] = make_int4(p[0], p[1], p[2], 0);
            }
            ctx->last_local[0] = a; ctx->last_local[1] = b; ctx->last_local[2] = c;
            ctx->last_acquired = s;
        } else if (ctx->regions.empty()) {
            const uint64_t lim = std::min<uint64_t>(ctx->slot_pos.size(), ctx->next_unused);
            const PosKey* const sp = ctx->slot_pos.data();
            for (; i < i1 && static_cast<uint64_t>(s) + 1 < lim; ++i) {
                const int32_t* p = pos + 3 * i;
                const PosKey q = sp[s + 1];
                if (q.x != p[0] || q.y != p[1] || q.z != p[2] || !res[s + 1] || stamps[s + 1] == cur) break;
                ++s; stamps[s] = cur;
                h_slots[i] = s;
                h_pos[i] = make_int4(p[0], p[1], p[2], 0);
            }
            ctx->last_acquired = s;
        }
    }
    return DSP_OK;
}

// Copies n elements of elem_bytes each from host `src` to dst_base[slot[i]], one cudaMemcpyAsync per
// run of consecutive slots (a fresh context hands out consecutive slots, so usually one copy).
int upload_by_runs(dsp_ctx* ctx, uint8_t* dst_base, size_t elem_bytes, const uint32_t* slots, uint64_t n, const uint8_t* src) {
    uint64_t i = 0;
    while (i < n) {
        uint64_t j = i + 1;
        while (j < n && slots[j] == slots[j - 1] + 1) ++j;
        DSP_CUDA(ctx, cudaMemcpyAsync(dst_base + static_cast<size_t>(slots[i]) * elem_bytes, src + i * elem_bytes, (j - i) * elem_bytes,
                                      cudaMemcpyHostToDevice, ctx->stream));
        i = j;
    }
    return DSP_OK;
}

// Positions (as int4) and, when the slots are not one consecutive run, the slot list go through one
// pinned staging block: [int4 x n][u32 x n].  *d_list is nullptr for a consecutive run.
int upload_positions_and_slots(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint32_t* slots, const uint32_t** d_list, uint32_t* first) {
    bool consecutive = true;
    for (uint64_t i = 1; i < n && consecutive; ++i) consecutive = slots[i] == slots[i - 1] + 1;
    *first = n ? slots[0] : 0;
    *d_list = nullptr;
    uint8_t* h;
    int rc = stage_host(ctx, n * (sizeof(int4) + sizeof(uint32_t)), &h);
    if (rc != DSP_OK) return rc;
    if (!consecutive) {
        uint32_t* hl = reinterpret_cast<uint32_t*>(h + n * sizeof(int4));
        std::memcpy(hl, slots, n * sizeof(uint32_t));
        DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_slots, hl, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        *d_list = ctx->d_slots;
    }
    if (pos) {
        int4* p4 = reinterpret_cast<int4*>(h);
        for (uint64_t i = 0; i < n; ++i) p4[i] = make_int4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0);
        if (consecutive) {
            DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_pos + *first, h, n * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
            DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_summary + *first, dsp::kSummaryUnknown, n, ctx->stream));  // the slots are being (re)defined
            DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_bvalid + *first, 0, n, ctx->stream));
            ctx->bmask_stale = true;
        } else {  // recycled slots are scattered: one copy + one scatter launch, not one copy per run of slots
            if ((rc = ensure_scratch(ctx, n * sizeof(int4))) != DSP_OK) return rc;
            DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, h, n * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
            dsp::scatter_pos_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_pos, ctx->d_summary, ctx->d_bvalid, ctx->d_slots,
                                                                                                  reinterpret_cast<const int4*>(ctx->d_scratch), static_cast<uint32_t>(n));
            ctx->bmask_stale = true;
            DSP_CUDA(ctx, cudaGetLastError());
        }
    }
    return stage_release(ctx);
}

#ifdef DSP_PHASE_TIMERS
constexpr size_t kPhaseWords = 32 + 1024 * 16;  // + a 16-word timeline per CTA (debug builds only)
#else
constexpr size_t kPhaseWords = 32;
#endif

int drain_kernel_event(dsp_ctx* ctx, int e) {
    if (!ctx->ev_pending[e]) return DSP_OK;
    float ms = 0.0f;
    DSP_CUDA(ctx, cudaEventSynchronize(ctx->ev_k1[e]));
    DSP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_k0[e], ctx->ev_k1[e]));
    ctx->kernel_ms_total += ms;
    ctx->kernel_ms_last = ms;
    ctx->kernel_count++;
    ctx->ev_pending[e] = false;
    return DSP_OK;
}

// dsp_mesh_kernel_stats times a mesh CALL's kernels: the event pair opens in front of the pre-pass (classify) launch when there
// is one, otherwise in front of the mesh kernel, and closes behind the mesh kernel.
int begin_kernel_timing(dsp_ctx* ctx) {
    if (!ctx->kernel_timing || ctx->timing_open) return DSP_OK;
    const int e = ctx->ev_next;
    int rc = drain_kernel_event(ctx, e);  // the pair from kEvRing launches ago: long finished
    if (rc != DSP_OK) return rc;
    DSP_CUDA(ctx, cudaEventRecord(ctx->ev_k0[e], ctx->stream));
    ctx->timing_open = true;
    return DSP_OK;
}
int end_kernel_timing(dsp_ctx* ctx) {
    if (!ctx->timing_open) return DSP_OK;
    const int e = ctx->ev_next;
    DSP_CUDA(ctx, cudaEventRecord(ctx->ev_k1[e], ctx->stream));
    ctx->ev_pending[e] = true;
    ctx->ev_next = (e + 1) % dsp_ctx::kEvRing;
    ctx->timing_open = false;
    return DSP_OK;
}

// Launch with the programmatic-stream-serialization attribute (the kernels start with griddepcontrol.wait, so the only thing
// that overlaps the previous kernel's tail is CTA scheduling and launch latency -- safe behind any predecessor).
template <class Kern>
cudaError_t launch_pdl(dsp_ctx* ctx, Kern kern, unsigned grid, int smem, const dsp::MeshParams& p, unsigned block = 256) {
    if (!ctx->use_pdl) { kern<<<grid, block, smem, ctx->stream>>>(p); return cudaGetLastError(); }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = static_cast<size_t>(smem); cfg.stream = ctx->stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <class Cfg, bool HALO>
int launch_mesh_t(dsp_ctx* ctx, const dsp::MeshParams& p, int variant) {
    void (*kern)(const dsp::MeshParams) = dsp::mesh_chunks_bump_kernel<Cfg, HALO>;
    constexpr int smem = Cfg::kBytes;
    int& occ = ctx->mesh_ctas_per_sm[variant][HALO ? 1 : 0];
    if (occ == 0) {
        DSP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        DSP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        if (occ < 1) return fail(ctx, DSP_ERR_INTERNAL, "mesh kernel does not fit on an SM (smem %d)", smem);
    }
    if (p.released != nullptr) {  // the launch runs beside the upload of its batch: the instantiation that waits for released pieces
        if (HALO) return fail(ctx, DSP_ERR_INTERNAL, "a launch beside its upload is chunk-local");
        kern = dsp::mesh_chunks_bump_kernel<Cfg, false, true>;
        if (!ctx->mesh_live_ready[variant]) {
            DSP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            ctx->mesh_live_ready[variant] = true;
        }
    }
    // (a launch that runs beside the upload of its own batch leaves one CTA slot per SM free: the unpack kernels it waits for must
    // find room next to its resident, possibly spinning CTAs)
    const uint64_t resident_ctas = static_cast<uint64_t>(ctx->num_sms) * (p.released != nullptr ? std::max(1, occ - 1) : occ);
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p.n, resident_ctas));
    dsp::MeshParams pp = p;
    pp.phase_cycles = ctx->d_phase;
    pp.prefetch_ahead = ctx->mesh_prefetch_mult * grid;
    int rc = begin_kernel_timing(ctx);
    if (rc != DSP_OK) return rc;
    if (pp.preset) {
        // DSP_MESH_ORDERED, pass 1: sizes and offsets of all chunks in call order (one warp per chunk; the second ticket line counts its CTAs out)
        dsp::MeshParams cp = pp;
        cp.ticket = pp.ticket + 32;
        void (*count_kern)(const dsp::MeshParams) = dsp::count_chunks_scan_kernel<HALO>;
        const unsigned cgrid = static_cast<unsigned>((p.n + dsp::kCountChunksPerCta - 1) / dsp::kCountChunksPerCta);
        DSP_CUDA(ctx, launch_pdl(ctx, count_kern, cgrid, 0, cp, dsp::kCountThreads));
        ctx->launches++;
    }
    DSP_CUDA(ctx, launch_pdl(ctx, kern, grid, smem, pp));
    ctx->launches++;
    return end_kernel_timing(ctx);
}

// Kernel shapes.  0 is what ships (and what bench.py reports); DSP_MESH_VARIANT=1..3 selects the others for A/B runs:
//   0: 3 CTAs/SM, one bulk store per 256-face tile     1: 3 CTAs/SM, one bulk store per warp (32 faces)
//   2: 4 CTAs/SM (8 KiB record window), tile           3: 4 CTAs/SM, warp
constexpr int kMeshVariants = 4;
template <bool HALO>
int launch_mesh_variant(dsp_ctx* ctx, const dsp::MeshParams& p) {
    // Cross-chunk culling leaves little to emit per chunk and adds dependent neighbour reads: what it needs is more CTAs in
    // flight, so its default is the 4-CTAs/SM shape (config-5 world: 28.4 -> 24.1 ms); DSP_MESH_VARIANT still overrides.
    const int variant = (HALO && !ctx->mesh_variant_forced) ? 2 : ctx->mesh_variant;
    switch (variant) {
        case 1: return launch_mesh_t<dsp::MeshCfg<3, 1, true>, HALO>(ctx, p, 1);
        case 2: return launch_mesh_t<dsp::MeshCfg<4, 1, false>, HALO>(ctx, p, 2);
        case 3: return launch_mesh_t<dsp::MeshCfg<4, 1, true>, HALO>(ctx, p, 3);
        default: return launch_mesh_t<dsp::MeshCfg<3, 1, false>, HALO>(ctx, p, 0);
    }
}

// Quad modes (DSP_MESH_GREEDY / DSP_MESH_INDEXED): mesh_chunks_quads_kernel, always the atomic slot claim.
template <bool HALO, bool GREEDY, bool INDEXED>
int launch_quads_t(dsp_ctx* ctx, const dsp::MeshParams& p_in) {
    void (*kern)(const dsp::MeshParams) = dsp::mesh_chunks_quads_kernel<HALO, GREEDY, INDEXED>;
    constexpr int smem = dsp::QuadSmem<GREEDY, GREEDY && INDEXED>::kBytes;
    int& occ = ctx->quad_ctas_per_sm[(HALO ? 1 : 0) + (GREEDY ? 2 : 0) + (INDEXED ? 4 : 0)];
    if (occ == 0) {
        DSP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        DSP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        if (occ < 1) return fail(ctx, DSP_ERR_INTERNAL, "quad mesh kernel does not fit on an SM (smem %d)", smem);
    }
    const int ctas = ctx->quad_ctas_cap > 0 ? std::min(occ, ctx->quad_ctas_cap) : occ;  // (DSP_QUAD_CTAS: A/B runs)
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p_in.n, static_cast<uint64_t>(ctx->num_sms) * ctas));
    dsp::MeshParams p = p_in;
    p.phase_cycles = ctx->d_phase;
    int rc = begin_kernel_timing(ctx);
    if (rc != DSP_OK) return rc;
    DSP_CUDA(ctx, launch_pdl(ctx, kern, grid, smem, p));
    ctx->launches++;
    return end_kernel_timing(ctx);
}
template <bool HALO>
int launch_quads(dsp_ctx* ctx, const dsp::MeshParams& p, uint32_t mode) {
    const bool g = (mode & DSP_MESH_GREEDY) != 0, i = (mode & DSP_MESH_INDEXED) != 0;
    if (g) return i ? launch_quads_t<HALO, true, true>(ctx, p) : launch_quads_t<HALO, true, false>(ctx, p);
    return launch_quads_t<HALO, false, true>(ctx, p);
}

int alloc_exportable_arena(dsp_ctx* ctx, uint64_t bytes, std::string* err);
void free_exportable_arena(dsp_ctx* ctx);
int build_neighbour_table(dsp_ctx* ctx);
int apply_halo_batches(dsp_ctx* ctx);
int apply_seam_batches(dsp_ctx* ctx);
void fill_seam_params(const dsp_ctx* ctx, dsp::MeshParams& p);
void destroy_seams(dsp_ctx* ctx);

int check_mesh_args(dsp_ctx* ctx, uint64_t n, uint32_t mode, uint64_t arena_offset) {
    if (mode & ~(DSP_MESH_HALO | DSP_MESH_ORDERED | DSP_MESH_GREEDY | DSP_MESH_INDEXED)) return fail(ctx, DSP_ERR_BAD_ARG, "unknown mesh mode bits 0x%x", mode);
    if ((mode & DSP_MESH_ORDERED) && (mode & (DSP_MESH_GREEDY | DSP_MESH_INDEXED)))
        return fail(ctx, DSP_ERR_BAD_ARG, "DSP_MESH_ORDERED is not available with DSP_MESH_GREEDY / DSP_MESH_INDEXED");
    if (arena_offset % DSP_SLOT_ALIGN != 0 || arena_offset > ctx->arena_bytes)
        return fail(ctx, DSP_ERR_BAD_ARG, "arena_offset %llu must be a multiple of 128 and <= arena size", (unsigned long long)arena_offset);
    if (n > ctx->max_chunks) return fail(ctx, DSP_ERR_BAD_ARG, "n = %llu exceeds max_chunks", (unsigned long long)n);
    if (ctx->n_blocks == 0) return fail(ctx, DSP_ERR_BAD_ARG, "dsp_set_uv_table has not been called");
    return DSP_OK;
}

// One mesh launch over `n` chunks whose entries start at entry index `entry0`; `sub` picks the ticket counter.
int launch_mesh(dsp_ctx* ctx, uint64_t n, const uint32_t* d_list, uint32_t first_slot, uint32_t mode, uint64_t arena_offset, uint64_t entry0, int sub, bool self_reset = false,
                bool work_list = false, const uint16_t* host_tiles = nullptr, const uint32_t* released = nullptr, uint32_t piece_chunks = 0) {
    const bool halo = (mode & DSP_MESH_HALO) != 0;
    dsp::MeshParams p{};
    p.voxels = ctx->d_voxels;
    p.host_tiles = host_tiles;  // the chunks of this launch are pulled out of (pinned) host memory by the kernel itself
    p.released = released;      // the launch runs beside the upload of its batch (packed upload, chunk-local parity mode only)
    p.piece_chunks = piece_chunks;
    p.voxels_out = ctx->d_voxels;
    p.chunk_pos = ctx->d_pos;
    p.slot_list = d_list;
    p.nbr_slots = ctx->d_nbr;
    p.bmask = ctx->d_bmask;
    p.halo_slabs = ctx->d_halo;
    fill_seam_params(ctx, p);
    p.uvmap = ctx->d_uvmap;
    p.n_blocks = ctx->n_blocks;
    p.first_slot = first_slot;
    p.n = static_cast<uint32_t>(n);
    p.arena = ctx->d_arena + arena_offset;
    p.arena_offset = arena_offset;
    p.arena_room = (ctx->arena_limit ? ctx->arena_limit : ctx->arena_bytes) - arena_offset;
    p.entries = ctx->d_entries + entry0;
    p.ticket = ctx->d_tickets + 32 * sub;  // its own cache line
    p.ctl = reinterpret_cast<uint32_t*>(ctx->d_ctl);  // control block: [-][status][done][n_front][total 8 B][n_back][n_uniform]
    p.status = reinterpret_cast<uint32_t*>(ctx->d_ctl) + 1;
    if (work_list) {  // classify_chunks_kernel left the chunks that need meshing in d_work_slots, their count in the control block
        p.slot_list = ctx->d_work_slots;
        p.entry_map = ctx->d_work_entry;
        p.n_dev = reinterpret_cast<const uint32_t*>(ctx->d_ctl) + 3;
        p.uid = ctx->d_uid;
    }
    if (self_reset) {
        p.done = reinterpret_cast<uint32_t*>(ctx->d_ctl) + 2;
        p.result = reinterpret_cast<uint64_t*>(ctx->d_ctl + 32 + 8 * ctx->max_chunks);
    }
    p.total_bytes = reinterpret_cast<uint64_t*>(ctx->d_ctl + 16);
    p.face_counts = reinterpret_cast<uint32_t*>(ctx->d_ctl + 32);
    p.preset = (mode & DSP_MESH_ORDERED) != 0;
    if (!halo && !work_list && ctx->summary_skip) {
        p.summary = ctx->d_summary;
        p.uid = ((ctx->summary_skip & 2) && !(mode & (DSP_MESH_GREEDY | DSP_MESH_INDEXED))) ? ctx->d_uid : nullptr;  // (the quad kernels only use the all-air bit)
        p.summary_air_skip = (ctx->summary_skip & 1) != 0;
    }
    if (mode & (DSP_MESH_GREEDY | DSP_MESH_INDEXED)) return halo ? launch_quads<true>(ctx, p, mode) : launch_quads<false>(ctx, p, mode);
    return halo ? launch_mesh_variant<true>(ctx, p) : launch_mesh_variant<false>(ctx, p);
}

// Cross-chunk culling reads a neighbour through its border occupancy masks.  Where a writer could not keep them current
// (anything but the analytic noise-region generator) the chunk's flag is clear: queue and recompute the masks of the chunks this
// call touches -- its own and their resident neighbours -- before the mesh kernel runs.  Two small launches, and only while the
// context knows that stale masks may exist.
int ensure_border_masks(dsp_ctx* ctx, uint64_t n, const uint32_t* d_list, uint32_t first_slot) {
    if (!ctx->bmask_stale || n == 0) return DSP_OK;
    if (!ctx->d_stale_list) DSP_CUDA(ctx, cudaMalloc(&ctx->d_stale_list, (ctx->max_chunks + 1) * sizeof(uint32_t)));
    uint32_t* count = ctx->d_stale_list + ctx->max_chunks;
    DSP_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(uint32_t), ctx->stream));
    dsp::find_stale_masks_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_bvalid, ctx->d_nbr, d_list, first_slot, static_cast<uint32_t>(n),
                                                                                                 ctx->d_stale_list, count);
    DSP_CUDA(ctx, cudaGetLastError());
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(std::min<uint64_t>(7 * n, ctx->max_chunks), static_cast<uint64_t>(ctx->num_sms) * 8));
    dsp::refresh_border_masks_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_voxels, ctx->d_bmask, ctx->d_bvalid, ctx->d_summary, ctx->d_uid, ctx->d_stale_list, count);
    DSP_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    if (n >= ctx->n_resident) ctx->bmask_stale = false;  // the call covers every resident chunk: nothing stale is left behind
    return DSP_OK;
}

int enqueue_mesh(dsp_ctx* ctx, uint64_t n, const uint32_t* d_list, uint32_t first_slot, uint32_t mode, uint64_t arena_offset) {
    int rc = check_mesh_args(ctx, n, mode, arena_offset);
    if (rc != DSP_OK) return rc;
    ctx->last_mesh_n = n;
    // Control block.  The mesh kernels reset it themselves at their end (control_block_epilogue), so in steady state a mesh call
    // is one kernel launch (two with DSP_MESH_ORDERED: the count + scan pass) and nothing else; a memset is only needed for the
    // first call and after a multi-launch call.
    const bool ordered = (mode & DSP_MESH_ORDERED) != 0;
    if (!ctx->ctl_clean || n == 0) {
        DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_ctl, 0, 32, ctx->stream));
        DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_tickets, 0, 256, ctx->stream));  // (the count pass of DSP_MESH_ORDERED has the second ticket line)
        ctx->ctl_clean = true;
    }

    ctx->result_in_tail = false;
    if (n == 0) return DSP_OK;
    if (mode & DSP_MESH_HALO) {
        if ((rc = build_neighbour_table(ctx)) != DSP_OK) return rc;
        if ((rc = ensure_border_masks(ctx, n, d_list, first_slot)) != DSP_OK) return rc;
    }
    ctx->ctl_clean = false;  // until the launch below is known to have been enqueued
    bool work_list = false;
    const bool halo_prepass = (mode & DSP_MESH_HALO) != 0;
    // merged-quad calls with chunk-local culling: all-air chunks and chunks of one solid block id are finished by the pre-pass
    // from their summaries (six quads, no tile read); worth its launch from about a thousand chunks on
    const bool greedy_prepass = !halo_prepass && (mode & DSP_MESH_GREEDY) && n >= 1024;
    if ((halo_prepass || greedy_prepass) && !ordered) {
        // Cross-chunk culling leaves most chunks of a large world without a single face (all air, or solid among solid
        // neighbours): a pre-pass over the chunk summaries writes their empty entries and hands the mesh kernel only the rest.
        if (!ctx->d_work_slots) {
            DSP_CUDA(ctx, cudaMalloc(&ctx->d_work_slots, ctx->max_chunks * sizeof(uint32_t)));
            DSP_CUDA(ctx, cudaMalloc(&ctx->d_work_entry, ctx->max_chunks * sizeof(uint32_t)));
        }
        const unsigned cgrid = static_cast<unsigned>((n + 255) / 256);
        if ((rc = begin_kernel_timing(ctx)) != DSP_OK) return rc;  // the pre-pass is part of the call's kernel time
        if (greedy_prepass) {
            dsp::ClassifyGreedyParams cp{};
            cp.summary = ctx->d_summary; cp.uid = ctx->d_uid; cp.chunk_pos = ctx->d_pos; cp.slot_list = d_list; cp.first_slot = first_slot;
            cp.n = static_cast<uint32_t>(n); cp.arena_offset = arena_offset;
            cp.arena_room = (ctx->arena_limit ? ctx->arena_limit : ctx->arena_bytes) - arena_offset;
            cp.arena = ctx->d_arena + arena_offset; cp.entries = ctx->d_entries; cp.work_slots = ctx->d_work_slots; cp.work_entry = ctx->d_work_entry;
            cp.n_work = reinterpret_cast<uint32_t*>(ctx->d_ctl) + 3;
            cp.total_bytes = reinterpret_cast<unsigned long long*>(ctx->d_ctl + 16);
            cp.status = reinterpret_cast<uint32_t*>(ctx->d_ctl) + 1;
            cp.uvmap = ctx->d_uvmap; cp.n_blocks = ctx->n_blocks;
            if (mode & DSP_MESH_INDEXED) dsp::classify_greedy_kernel<true><<<cgrid, 256, 0, ctx->stream>>>(cp);
            else dsp::classify_greedy_kernel<false><<<cgrid, 256, 0, ctx->stream>>>(cp);
        } else {
            dsp::classify_chunks_kernel<<<cgrid, 256, 0, ctx->stream>>>(
                ctx->d_summary, ctx->d_nbr, d_list, first_slot, static_cast<uint32_t>(n), arena_offset, ctx->d_entries, ctx->d_work_slots, ctx->d_work_entry,
                reinterpret_cast<uint32_t*>(ctx->d_ctl) + 3);
        }
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        work_list = true;
    }
    rc = launch_mesh(ctx, n, d_list, first_slot, mode, arena_offset, 0, 0, true, work_list);
    if (rc == DSP_OK) { ctx->ctl_clean = true; ctx->result_in_tail = true; }
    return rc;
}

int validate_slots(dsp_ctx* ctx, uint64_t n, const uint32_t* slots) {
    for (uint64_t i = 0; i < n; ++i)
        if (slots[i] >= ctx->next_unused || !ctx->resident[slots[i]])
            return fail(ctx, DSP_ERR_NOT_RESIDENT, "slot %u (entry %llu) is not resident", slots[i], (unsigned long long)i);
    return DSP_OK;
}
int validate_range(dsp_ctx* ctx, uint32_t first, uint64_t n) {
    if (first + n > ctx->next_unused) return fail(ctx, DSP_ERR_NOT_RESIDENT, "slot range [%u, %llu) is not resident", first, (unsigned long long)(first + n));
    for (uint64_t i = 0; i < n; ++i)
        if (!ctx->resident[first + i]) return fail(ctx, DSP_ERR_NOT_RESIDENT, "slot %llu is not resident", (unsigned long long)(first + i));
    return DSP_OK;
}

// Received border slabs -> neighbour table, on the device: entry (slot, face) that is still "none" after the resident
// neighbours were resolved points at the slab.  Newest batch first, so a (slot, face) registered again uses the newer slab.
int apply_halo_batches(dsp_ctx* ctx) {
    { int rc = apply_seam_batches(ctx); if (rc != DSP_OK) return rc; }  // a connected seam takes precedence over hand-registered slabs
    for (auto it = ctx->halo_batches.rbegin(); it != ctx->halo_batches.rend(); ++it) {
        if (it->n == 0) continue;
        dsp::apply_halo_batch_kernel<<<(it->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_nbr, ctx->d_halo_slots + it->first_index, it->face, 0u, it->first_index, it->n);
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return DSP_OK;
}

// Neighbour slots for halo culling: order top(+Y), bottom(-Y), rear(-Z), front(+Z), right(-X), left(+X)
// (the reference's face order, chunk_mesh.rs:46-51).  Resolved on the host from the chunk map.
int build_neighbour_table(dsp_ctx* ctx) {
    if (!ctx->d_nbr) DSP_CUDA(ctx, cudaMalloc(&ctx->d_nbr, ctx->max_chunks * 6 * sizeof(uint32_t)));
    if (!ctx->nbr_dirty) return DSP_OK;
    // Fast path: the world is one fully resident box (dsp_generate_region, nothing unloaded, no individually loaded
    // chunks) -- neighbour slots are arithmetic, computed on the device; received halo slabs are scattered over the
    // box-boundary entries.  Millions of chunks never touch the host.
    if (ctx->map.empty() && ctx->regions.size() == 1) {
        const Region& r = ctx->regions[0];
        const uint64_t n = static_cast<uint64_t>(r.dims[0]) * r.dims[1] * r.dims[2];
        if (ctx->n_resident == n) {
            dsp::region_neighbours_kernel<<<static_cast<unsigned>(std::min<uint64_t>((n + 255) / 256, 65535)), 256, 0, ctx->stream>>>(
                ctx->d_nbr, r.first_slot, make_uint3(r.dims[0], r.dims[1], r.dims[2]), n);
            DSP_CUDA(ctx, cudaGetLastError());
            ctx->launches++;
            { int rc = apply_halo_batches(ctx); if (rc != DSP_OK) return rc; }
            ctx->nbr_dirty = false;
            return DSP_OK;
        }
    }
    static const int dirs[6][3] = {{0, 1, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}, {-1, 0, 0}, {1, 0, 0}};
    std::vector<uint32_t> nbr(ctx->next_unused * 6, DSP_NO_SLOT);
    auto fill = [&](uint32_t slot, const int32_t p[3]) {
        for (int f = 0; f < 6; ++f) {
            const int32_t q[3] = {p[0] + dirs[f][0], p[1] + dirs[f][1], p[2] + dirs[f][2]};
            uint32_t s;
            if (find_slot(ctx, q, &s)) nbr[static_cast<size_t>(slot) * 6 + f] = s;
        }
    };
    for (const auto& kv : ctx->map) { const int32_t p[3] = {kv.first.x, kv.first.y, kv.first.z}; fill(kv.second, p); }
    for (const Region& r : ctx->regions)
        for (uint32_t c = 0; c < r.dims[2]; ++c)
            for (uint32_t b = 0; b < r.dims[1]; ++b)
                for (uint32_t a = 0; a < r.dims[0]; ++a) {
                    const uint64_t s = r.first_slot + a + static_cast<uint64_t>(r.dims[0]) * (b + static_cast<uint64_t>(r.dims[1]) * c);
                    if (!ctx->resident[s]) continue;
                    const int32_t p[3] = {r.origin[0] + static_cast<int32_t>(a), r.origin[1] + static_cast<int32_t>(b), r.origin[2] + static_cast<int32_t>(c)};
                    fill(static_cast<uint32_t>(s), p);
                }
    if (!nbr.empty()) {
        DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_nbr, nbr.data(), nbr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    { int rc = apply_halo_batches(ctx); if (rc != DSP_OK) return rc; }  // where no resident neighbour exists, a received slab may stand in
    ctx->nbr_dirty = false;
    return DSP_OK;
}

// Device-side edit pipeline (terrain_kernels.cuh "device-side edit pipeline"): usable when every resident chunk lives in
// a full box region, so that world -> slot is arithmetic.  On success the ascending dirty-slot list is in ctx->d_slots and
// its length in *n_dirty.  *handled == false: the caller must take the host path.
int apply_edits_device(dsp_ctx* ctx, uint64_t n, const dsp_edit* edits, bool halo, bool* handled, uint32_t* n_dirty) {
    *handled = false;
    *n_dirty = 0;
    if (!ctx->map.empty() || ctx->regions.empty() || ctx->regions.size() > dsp::kMaxEditRegions || ctx->region_holes || n >= (1u << 20)) return DSP_OK;
    uint32_t table_size = 1024;
    while (table_size < 4 * n) table_size <<= 1;
    const uint32_t n_words = static_cast<uint32_t>((ctx->next_unused + 31) / 32);
    const size_t off_table = (n * sizeof(int4) + 255) & ~size_t(255);
    const size_t off_bitmap = off_table + static_cast<size_t>(table_size) * 8;
    const size_t off_count = off_bitmap + static_cast<size_t>(n_words) * 4;
    int rc = ensure_scratch(ctx, off_count + 256);
    if (rc != DSP_OK) return rc;
    uint8_t* h;
    if ((rc = stage_host(ctx, n * sizeof(dsp_edit), &h)) != DSP_OK) return rc;
    std::memcpy(h, edits, n * sizeof(dsp_edit));
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, h, n * sizeof(dsp_edit), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = stage_release(ctx)) != DSP_OK) return rc;
    DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch + off_table, 0, off_count + 4 - off_table, ctx->stream));
    dsp::EditParams p{};
    p.edits = reinterpret_cast<const int4*>(ctx->d_scratch);
    p.n = static_cast<uint32_t>(n);
    p.n_regions = static_cast<int>(ctx->regions.size());
    for (int r = 0; r < p.n_regions; ++r) {
        const Region& g = ctx->regions[r];
        p.regions[r] = dsp::EditRegion{g.origin[0], g.origin[1], g.origin[2], g.dims[0], g.dims[1], g.dims[2], g.first_slot};
    }
    p.table = reinterpret_cast<unsigned long long*>(ctx->d_scratch + off_table);
    p.table_mask = table_size - 1;
    p.voxels = ctx->d_voxels;
    p.dirty_bitmap = reinterpret_cast<uint32_t*>(ctx->d_scratch + off_bitmap);
    p.nbr = nullptr;
    p.summary = ctx->d_summary;
    p.bvalid = ctx->d_bvalid;
    ctx->bmask_stale = true;
    if (halo) {
        if ((rc = build_neighbour_table(ctx)) != DSP_OK) return rc;
        p.nbr = ctx->d_nbr;
    }
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    dsp::edit_file_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    DSP_CUDA(ctx, cudaGetLastError());
    dsp::edit_apply_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    DSP_CUDA(ctx, cudaGetLastError());
    uint32_t* d_count = reinterpret_cast<uint32_t*>(ctx->d_scratch + off_count);
    dsp::bitmap_to_list_kernel<<<1, 1024, 0, ctx->stream>>>(p.dirty_bitmap, n_words, ctx->d_slots, d_count);
    DSP_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 64, d_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_dirty = *reinterpret_cast<uint32_t*>(ctx->h_pinned + 64);
    *handled = true;
    return DSP_OK;
}

// `first_entry`: entries below it have already been copied out (sub-batch pipeline of the host-chunk calls)
int finish_mesh(dsp_ctx* ctx, dsp_mesh_entry* out_entries, uint64_t* out_total, uint64_t first_entry = 0) {
    const uint64_t n = ctx->last_mesh_n;
    if (out_total) *out_total = 0;
    if (n > first_entry && out_entries)
        DSP_CUDA(ctx, cudaMemcpyAsync(out_entries + first_entry, ctx->d_entries + first_entry, (n - first_entry) * sizeof(dsp_mesh_entry), cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t status;
    uint64_t total;
    if (ctx->result_in_tail) {  // written by the last CTA of a self-resetting launch
        DSP_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_ctl + 32 + 8 * ctx->max_chunks, 16, cudaMemcpyDeviceToHost, ctx->stream));
        DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        status = static_cast<uint32_t>(reinterpret_cast<uint64_t*>(ctx->h_pinned)[0]);
        total = reinterpret_cast<uint64_t*>(ctx->h_pinned)[1];
    } else {
        DSP_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_ctl, 32, cudaMemcpyDeviceToHost, ctx->stream));
        DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        status = reinterpret_cast<uint32_t*>(ctx->h_pinned)[1];
        total = *reinterpret_cast<uint64_t*>(ctx->h_pinned + 16);
    }
    if (out_total) *out_total = total;
    if (status & dsp::kStatusSpin) return fail(ctx, DSP_ERR_INTERNAL, "the mesh launch that runs beside the upload of its batch timed out waiting for chunks to be released");
    if (status & dsp::kStatusSeam)
        return fail(ctx, DSP_ERR_INTERNAL, "a region seam timed out: the neighbour's border layers (or its ack) for epoch %u never arrived -- every side of a seam "
                    "must call dsp_seam_exchange_async once per epoch, before its halo-mode mesh call", ctx->seam_epoch);
    if (status & dsp::kStatusOverflow)
        return fail(ctx, DSP_ERR_ARENA_FULL, "vertex arena too small: need %llu bytes from the given offset", (unsigned long long)total);
    return DSP_OK;
}

}  // namespace

extern "C" {

int dsp_abi_version(void) { return DSP_ABI_VERSION; }

const char* dsp_status_string(int s) {
    switch (s) {
        case DSP_OK: return "ok";
        case DSP_ERR_BAD_ARG: return "bad argument";
        case DSP_ERR_CUDA: return "CUDA error";
        case DSP_ERR_NO_DEVICE: return "no CUDA device";
        case DSP_ERR_ARENA_FULL: return "vertex arena full";
        case DSP_ERR_STORE_FULL: return "voxel store full";
        case DSP_ERR_NOT_RESIDENT: return "chunk not resident";
        case DSP_ERR_INTERNAL: return "internal error";
        default: return "unknown status";
    }
}

const char* dsp_last_error(const dsp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dsp_create(int device, uint64_t max_chunks, uint64_t arena_bytes, dsp_ctx** out_ctx) { return dsp_create_ex(device, max_chunks, arena_bytes, 0u, out_ctx); }

int dsp_create_ex(int device, uint64_t max_chunks, uint64_t arena_bytes, uint32_t flags, dsp_ctx** out_ctx) {
    if (!out_ctx) return fail(nullptr, DSP_ERR_BAD_ARG, "out_ctx is NULL");
    if (flags & ~DSP_CREATE_EXPORTABLE_ARENA) return fail(nullptr, DSP_ERR_BAD_ARG, "unknown dsp_create_ex flags 0x%x", flags);
    *out_ctx = nullptr;
    if (max_chunks == 0 || max_chunks >= 0x80000000ull) return fail(nullptr, DSP_ERR_BAD_ARG, "max_chunks must be in [1, 2^31-1]");
    if (arena_bytes % DSP_SLOT_ALIGN != 0) return fail(nullptr, DSP_ERR_BAD_ARG, "arena_bytes must be a multiple of 128");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, DSP_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, DSP_ERR_BAD_ARG, "device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, DSP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, DSP_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    dsp_ctx* ctx = new dsp_ctx();
    ctx->device = device;
    ctx->max_chunks = max_chunks;
    ctx->arena_bytes = arena_bytes;
    ctx->num_sms = prop.multiProcessorCount;
    if (const char* s = getenv("DSP_PDL")) ctx->use_pdl = atoi(s) != 0;
    if (const char* s = getenv("DSP_SUMMARY_SKIP")) ctx->summary_skip = atoi(s) & 3;
    if (const char* s = getenv("DSP_E2E_PULL")) ctx->e2e_pull = std::min(2, std::max(0, atoi(s)));
    if (const char* s = getenv("DSP_E2E_PULL_PIECES")) ctx->e2e_pull_pieces = atoi(s);
    if (const char* s = getenv("DSP_E2E_PULL_DIV")) ctx->e2e_pull_first_div = static_cast<uint64_t>(std::min(64, std::max(2, atoi(s))));
    if (const char* s = getenv("DSP_E2E_HEAD")) ctx->e2e_head_chunks = static_cast<uint64_t>(std::min(512, std::max(0, atoi(s))));
    if (const char* s = getenv("DSP_E2E_PACK")) ctx->e2e_pack = atoi(s);
    if (const char* s = getenv("DSP_HOST_THREADS")) ctx->host_threads = std::min(63, std::max(-1, atoi(s)));
    if (const char* s = getenv("DSP_E2E_PACK_KEEP")) ctx->e2e_pack_keep = std::min(16, std::max(0, atoi(s)));
    if (const char* s = getenv("DSP_E2E_PACK_PIECES")) ctx->e2e_pack_pieces = std::min(16, std::max(1, atoi(s)));
    if (const char* s = getenv("DSP_E2E_PACK_PERSIST")) ctx->e2e_pack_persist = atoi(s);
    if (const char* s = getenv("DSP_E2E_PACK_PERSIST_PIECES")) ctx->e2e_pack_persist_pieces = std::min(32, std::max(1, atoi(s)));
    if (const char* s = getenv("DSP_QUAD_CTAS")) ctx->quad_ctas_cap = atoi(s);
    if (const char* s = getenv("DSP_MESH_PREFETCH")) ctx->mesh_prefetch_mult = static_cast<unsigned>(std::max(0, atoi(s)));
    if (const char* s = getenv("DSP_MESH_VARIANT")) { const int v = atoi(s); if (v >= 0 && v < kMeshVariants) { ctx->mesh_variant = v; ctx->mesh_variant_forced = true; } }
    auto bail = [&](const char* what, cudaError_t err) {
        std::string msg = std::string(what) + ": " + cudaGetErrorString(err);
        dsp_destroy(ctx);
        return fail(nullptr, DSP_ERR_CUDA, "%s", msg.c_str());
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaMalloc(&ctx->d_voxels, max_chunks * dsp::kChunkBytes)) != cudaSuccess) return bail("cudaMalloc(voxel store)", e);
    if ((e = cudaMalloc(&ctx->d_pos, max_chunks * sizeof(int4))) != cudaSuccess) return bail("cudaMalloc(chunk positions)", e);
    if (flags & DSP_CREATE_EXPORTABLE_ARENA) {
        std::string why;
        if (alloc_exportable_arena(ctx, arena_bytes, &why) != DSP_OK) { dsp_destroy(ctx); return fail(nullptr, DSP_ERR_CUDA, "exportable arena: %s", why.c_str()); }
    } else if ((e = cudaMalloc(&ctx->d_arena, std::max<uint64_t>(arena_bytes, 128))) != cudaSuccess) return bail("cudaMalloc(vertex arena)", e);
    if ((e = cudaMalloc(&ctx->d_entries, max_chunks * sizeof(dsp::MeshEntry))) != cudaSuccess) return bail("cudaMalloc(entries)", e);
    if ((e = cudaMalloc(&ctx->d_ctl, 32 + 8 * max_chunks + 64)) != cudaSuccess) return bail("cudaMalloc(control)", e);
    if ((e = cudaMemset(ctx->d_ctl, 0, 32 + 8 * max_chunks + 64)) != cudaSuccess) return bail("cudaMemset(control)", e);  // (scan words: epoch 0 is never used)
    if ((e = cudaMalloc(&ctx->d_slots, max_chunks * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc(slot list)", e);
    if ((e = cudaMalloc(&ctx->d_summary, max_chunks)) != cudaSuccess) return bail("cudaMalloc(chunk summaries)", e);
    if ((e = cudaMemset(ctx->d_summary, 0, max_chunks)) != cudaSuccess) return bail("cudaMemset(chunk summaries)", e);
    if ((e = cudaMalloc(&ctx->d_uid, max_chunks * sizeof(uint16_t))) != cudaSuccess) return bail("cudaMalloc(uniform ids)", e);
    if ((e = cudaMalloc(&ctx->d_bmask, max_chunks * dsp::kBorderMaskStride * sizeof(uint16_t))) != cudaSuccess) return bail("cudaMalloc(border masks)", e);
    if ((e = cudaMalloc(&ctx->d_bvalid, (max_chunks + 3) & ~uint64_t(3))) != cudaSuccess) return bail("cudaMalloc(border mask flags)", e);
    if ((e = cudaMemset(ctx->d_bvalid, 0, (max_chunks + 3) & ~uint64_t(3))) != cudaSuccess) return bail("cudaMemset(border mask flags)", e);
    if ((e = cudaMalloc(&ctx->d_phase, kPhaseWords * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc(phase)", e);
    cudaMemset(ctx->d_phase, 0, kPhaseWords * sizeof(unsigned long long));
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (int i = 0; i < dsp_ctx::kMaxSub; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_sub[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    // one 128-byte line per ticket counter, away from the control block: the ticket and the arena cursor are the two hot atomics
    // of a launch, and on one line they queue up behind each other in the same L2 slice
    if ((e = cudaMalloc(&ctx->d_tickets, dsp_ctx::kMaxSub * 128)) != cudaSuccess) return bail("cudaMalloc(tickets)", e);
    if ((e = cudaMemset(ctx->d_tickets, 0, dsp_ctx::kMaxSub * 128)) != cudaSuccess) return bail("cudaMemset(tickets)", e);
    if ((e = cudaMallocHost(&ctx->h_pinned, 4096)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i < dsp_ctx::kEvRing; ++i)
        if ((e = cudaEventCreate(&ctx->ev_k0[i])) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_k1[i])) != cudaSuccess) return bail("cudaEventCreate", e);
    ctx->resident.assign(max_chunks, 0);
    *out_ctx = ctx;
    return DSP_OK;
}

int dsp_destroy(dsp_ctx* ctx) {
    if (!ctx) return DSP_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    destroy_seams(ctx);
    cudaFree(ctx->d_halo_slots);
    if (ctx->arena_vmm_size) free_exportable_arena(ctx);
    cudaFree(ctx->d_voxels); cudaFree(ctx->d_pos); cudaFree(ctx->d_nbr); cudaFree(ctx->d_halo); cudaFree(ctx->d_arena); cudaFree(ctx->d_entries);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->copy_stream2) { cudaStreamSynchronize(ctx->copy_stream2); cudaStreamDestroy(ctx->copy_stream2); }
    if (ctx->d2h_stream) { cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); }
    for (int i = 0; i < dsp_ctx::kMaxSub; ++i) if (ctx->ev_meshed[i]) cudaEventDestroy(ctx->ev_meshed[i]);
    for (int i = 0; i < dsp_ctx::kMaxSub; ++i) if (ctx->ev_sub[i]) cudaEventDestroy(ctx->ev_sub[i]);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    cudaFree(ctx->d_tickets);
    cudaFree(ctx->d_uid); cudaFree(ctx->d_bmask); cudaFree(ctx->d_bvalid);
    cudaFree(ctx->d_stale_list);
    cudaFree(ctx->d_summary); cudaFree(ctx->d_work_slots); cudaFree(ctx->d_work_entry);
    cudaFree(ctx->d_phase); cudaFree(ctx->d_ctl); cudaFree(ctx->d_slots); cudaFree(ctx->d_uv_rows); cudaFree(ctx->d_uvmap); cudaFree(ctx->d_scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    delete ctx->pack_pool;  // (joins the workers)
    if (ctx->h_pack) cudaFreeHost(ctx->h_pack);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
    for (int i = 0; i < dsp_ctx::kEvRing; ++i) {
        if (ctx->ev_k0[i]) cudaEventDestroy(ctx->ev_k0[i]);
        if (ctx->ev_k1[i]) cudaEventDestroy(ctx->ev_k1[i]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DSP_OK;
}

int dsp_set_uv_table(dsp_ctx* ctx, uint32_t n_blocks, const float* uv_rows) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (n_blocks == 0 || n_blocks > 65536 || !uv_rows) return fail(ctx, DSP_ERR_BAD_ARG, "n_blocks must be in [1,65536] and uv_rows non-NULL");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_blocks > ctx->uv_cap) {
        DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_uv_rows); cudaFree(ctx->d_uvmap);
        ctx->d_uv_rows = ctx->d_uvmap = nullptr;
        DSP_CUDA(ctx, cudaMalloc(&ctx->d_uv_rows, n_blocks * sizeof(float4)));
        DSP_CUDA(ctx, cudaMalloc(&ctx->d_uvmap, (n_blocks + 1) * sizeof(float4)));
        ctx->uv_cap = n_blocks;
    }
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_uv_rows, uv_rows, n_blocks * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    dsp::uv_map_kernel<<<(n_blocks + 1 + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_uv_rows, ctx->d_uvmap, n_blocks);
    DSP_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->n_blocks = n_blocks;
    return DSP_OK;
}

// World::get_chunk (world.rs:28-47): a position that is already in `chunks` is returned as it is -- edits made to it since
// it was generated survive a second load_chunk; only absent positions are generated.  `regenerate` forces the generator
// over resident positions as well (dsp_regenerate_chunks: no counterpart in the reference).
static int generate_chunks_impl(dsp_ctx* ctx, uint64_t n, const int32_t* pos, int gen_kind, uint32_t seed, uint32_t* out_slots, bool regenerate) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (gen_kind < 0 || gen_kind > 3) return fail(ctx, DSP_ERR_BAD_ARG, "unknown gen_kind %d", gen_kind);
    if (n == 0) return DSP_OK;
    if (!pos) return fail(ctx, DSP_ERR_BAD_ARG, "pos is NULL");
    if (n > ctx->max_chunks) return fail(ctx, DSP_ERR_STORE_FULL, "n exceeds max_chunks");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> slots(n), fresh_slots;
    std::vector<int32_t> todo_pos;
    std::vector<uint32_t> todo_slots;
    new_stamp(ctx);
    for (uint64_t i = 0; i < n; ++i) {
        bool fresh = false;
        int rc = acquire_slot(ctx, pos + 3 * i, &slots[i], &fresh);
        if (rc != DSP_OK) { release_fresh_slots(ctx, fresh_slots); return rc; }  // the call did not happen
        if (fresh) fresh_slots.push_back(slots[i]);
        if ((fresh || regenerate) && stamp_slot(ctx, slots[i])) {  // (a position listed twice is generated once)
            todo_pos.insert(todo_pos.end(), pos + 3 * i, pos + 3 * i + 3);
            todo_slots.push_back(slots[i]);
        }
    }
    if (out_slots) std::memcpy(out_slots, slots.data(), n * sizeof(uint32_t));
    const uint64_t m = todo_slots.size();
    if (m == 0) return DSP_OK;
    const uint32_t* d_list; uint32_t first;
    int rc = upload_positions_and_slots(ctx, m, todo_pos.data(), todo_slots.data(), &d_list, &first);
    if (rc != DSP_OK) { cudaStreamSynchronize(ctx->stream); release_fresh_slots(ctx, fresh_slots); return rc; }
    dsp::TerrainParams p{ctx->d_voxels, ctx->d_pos, d_list, first, static_cast<uint32_t>(m), gen_kind, seed, make_int3(0, 0, 0), make_uint3(0, 0, 0), nullptr, ctx->d_summary, ctx->d_uid, ctx->d_bvalid};
    ctx->bmask_stale = true;
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(m, static_cast<uint64_t>(ctx->num_sms) * 64));
    dsp::terrain_fill_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    DSP_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return DSP_OK;
}

int dsp_generate_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, int gen_kind, uint32_t seed, uint32_t* out_slots) {
    return generate_chunks_impl(ctx, n, pos, gen_kind, seed, out_slots, false);
}
int dsp_regenerate_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, int gen_kind, uint32_t seed, uint32_t* out_slots) {
    return generate_chunks_impl(ctx, n, pos, gen_kind, seed, out_slots, true);
}

int dsp_generate_region(dsp_ctx* ctx, const int32_t origin[3], const uint32_t dims[3], int gen_kind, uint32_t seed, uint32_t* out_first_slot) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (!origin || !dims || !out_first_slot) return fail(ctx, DSP_ERR_BAD_ARG, "NULL argument");
    if (gen_kind < 0 || gen_kind > 3) return fail(ctx, DSP_ERR_BAD_ARG, "unknown gen_kind %d", gen_kind);
    const uint64_t n = static_cast<uint64_t>(dims[0]) * dims[1] * dims[2];
    if (n == 0) return fail(ctx, DSP_ERR_BAD_ARG, "empty region");
    if (ctx->next_unused + n > ctx->max_chunks) return fail(ctx, DSP_ERR_STORE_FULL, "region of %llu chunks does not fit (%llu unused slots)",
                                                            (unsigned long long)n, (unsigned long long)(ctx->max_chunks - ctx->next_unused));
    if (region_overlaps(ctx, origin, dims))
        return fail(ctx, DSP_ERR_BAD_ARG, "region (%d, %d, %d) + %u x %u x %u overlaps chunks that are already resident (a position has one chunk: "
                    "use dsp_clone_region for identical copies elsewhere)", origin[0], origin[1], origin[2], dims[0], dims[1], dims[2]);
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t first = static_cast<uint32_t>(ctx->next_unused);
    ctx->next_unused += n;
    Region r{{origin[0], origin[1], origin[2]}, {dims[0], dims[1], dims[2]}, first};
    ctx->regions.push_back(r);
    std::fill(ctx->resident.begin() + first, ctx->resident.begin() + first + n, 1);
    ctx->n_resident += n;
    ctx->nbr_dirty = true;
    if (gen_kind == DSP_GEN_NOISE) {
        // heightmap computed once per chunk column and reused up the stack (terrain_fill_region_noise_kernel)
        dsp::RegionFillParams rp{ctx->d_voxels, ctx->d_pos, first, make_int3(origin[0], origin[1], origin[2]), make_uint3(dims[0], dims[1], dims[2]), 0u, seed, ctx->d_summary, ctx->d_uid, ctx->d_bmask, ctx->d_bvalid};
        // enough CTAs to fill the machine a few times over, but as tall a run per CTA as that allows
        const uint64_t columns = static_cast<uint64_t>(dims[0]) * dims[2], want = static_cast<uint64_t>(ctx->num_sms) * 16;
        uint32_t pieces = static_cast<uint32_t>(std::min<uint64_t>(dims[1], std::max<uint64_t>(1, (want + columns - 1) / columns)));
        rp.ny = (dims[1] + pieces - 1) / pieces;
        pieces = (dims[1] + rp.ny - 1) / rp.ny;
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(columns * pieces, static_cast<uint64_t>(ctx->num_sms) * 32));
        dsp::terrain_fill_region_noise_kernel<<<grid, 256, 0, ctx->stream>>>(rp);
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    } else {
        if (n > 0xFFFFFFFFull) return fail(ctx, DSP_ERR_BAD_ARG, "region too large");
        // positions are arithmetic in a box: the fill kernel derives them from the chunk's index and records them itself
        dsp::TerrainParams p{ctx->d_voxels, ctx->d_pos, nullptr, first, static_cast<uint32_t>(n), gen_kind, seed,
                             make_int3(origin[0], origin[1], origin[2]), make_uint3(dims[0], dims[1], dims[2]), ctx->d_pos, ctx->d_summary, ctx->d_uid, ctx->d_bvalid};
        ctx->bmask_stale = true;
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(n, static_cast<uint64_t>(ctx->num_sms) * 64));
        dsp::terrain_fill_kernel<<<grid, 256, 0, ctx->stream>>>(p);
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    *out_first_slot = first;
    return DSP_OK;
}

int dsp_clone_region(dsp_ctx* ctx, uint32_t src_first_slot, const int32_t new_origin[3], uint32_t* out_first_slot) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (!new_origin || !out_first_slot) return fail(ctx, DSP_ERR_BAD_ARG, "NULL argument");
    const Region* src = nullptr;
    for (const Region& r : ctx->regions) if (r.first_slot == src_first_slot) src = &r;
    if (!src) return fail(ctx, DSP_ERR_BAD_ARG, "slot %u is not the first slot of a dsp_generate_region box", src_first_slot);
    const Region from = *src;
    const uint64_t n = static_cast<uint64_t>(from.dims[0]) * from.dims[1] * from.dims[2];
    for (uint64_t i = 0; i < n; ++i)
        if (!ctx->resident[from.first_slot + i]) return fail(ctx, DSP_ERR_NOT_RESIDENT, "source region has unloaded chunks");
    if (ctx->next_unused + n > ctx->max_chunks) return fail(ctx, DSP_ERR_STORE_FULL, "clone of %llu chunks does not fit", (unsigned long long)n);
    if (region_overlaps(ctx, new_origin, from.dims)) return fail(ctx, DSP_ERR_BAD_ARG, "the clone would overlap resident chunks");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t first = static_cast<uint32_t>(ctx->next_unused);
    ctx->next_unused += n;
    ctx->regions.push_back(Region{{new_origin[0], new_origin[1], new_origin[2]}, {from.dims[0], from.dims[1], from.dims[2]}, first});
    std::fill(ctx->resident.begin() + first, ctx->resident.begin() + first + n, 1);
    ctx->n_resident += n;
    ctx->nbr_dirty = true;
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_voxels + static_cast<size_t>(first) * dsp::kChunkVoxels, ctx->d_voxels + static_cast<size_t>(from.first_slot) * dsp::kChunkVoxels,
                                  n * dsp::kChunkBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_summary + first, ctx->d_summary + from.first_slot, n, cudaMemcpyDeviceToDevice, ctx->stream));
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_uid + first, ctx->d_uid + from.first_slot, n * sizeof(uint16_t), cudaMemcpyDeviceToDevice, ctx->stream));
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_bmask + static_cast<size_t>(first) * dsp::kBorderMaskStride, ctx->d_bmask + static_cast<size_t>(from.first_slot) * dsp::kBorderMaskStride,
                                  n * dsp::kBorderMaskStride * sizeof(uint16_t), cudaMemcpyDeviceToDevice, ctx->stream));
    DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_bvalid + first, 0, n, ctx->stream));  // (re-derived on demand; the copied masks are only overwritten with the same values)
    ctx->bmask_stale = true;
    dsp::region_positions_kernel<<<static_cast<unsigned>(std::min<uint64_t>((n + 255) / 256, 65535)), 256, 0, ctx->stream>>>(
        ctx->d_pos, first, make_int3(new_origin[0], new_origin[1], new_origin[2]), make_uint3(from.dims[0], from.dims[1], from.dims[2]), n);
    DSP_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    *out_first_slot = first;
    return DSP_OK;
}

int dsp_invalidate_chunks(dsp_ctx* ctx, uint64_t n, const uint32_t* slots) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (n == 0) return DSP_OK;
    if (!slots) return fail(ctx, DSP_ERR_BAD_ARG, "slots is NULL");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = validate_slots(ctx, n, slots);
    if (rc != DSP_OK) return rc;
    uint8_t* h;
    if ((rc = stage_host(ctx, n * sizeof(uint32_t), &h)) != DSP_OK) return rc;
    std::memcpy(h, slots, n * sizeof(uint32_t));
    if ((rc = ensure_scratch(ctx, n * sizeof(uint32_t))) != DSP_OK) return rc;
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, h, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = stage_release(ctx)) != DSP_OK) return rc;
    dsp::invalidate_summaries_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_summary, ctx->d_bvalid, reinterpret_cast<const uint32_t*>(ctx->d_scratch), static_cast<uint32_t>(n));
    ctx->bmask_stale = true;
    DSP_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return DSP_OK;
}

int dsp_load_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, const uint16_t* palette_ids,
                    const uint32_t* palette_offsets, uint32_t* out_slots) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (n == 0) return DSP_OK;
    if (!pos || !blocks) return fail(ctx, DSP_ERR_BAD_ARG, "pos/blocks is NULL");
    if ((palette_ids == nullptr) != (palette_offsets == nullptr)) return fail(ctx, DSP_ERR_BAD_ARG, "palette_ids and palette_offsets go together");
    if (n > ctx->max_chunks) return fail(ctx, DSP_ERR_STORE_FULL, "n exceeds max_chunks");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> slots(n), fresh_slots;
    new_stamp(ctx);
    for (uint64_t i = 0; i < n; ++i) {
        bool fresh = false;
        int rc = acquire_slot(ctx, pos + 3 * i, &slots[i], &fresh);
        if (rc == DSP_OK && !stamp_slot(ctx, slots[i]))
            rc = fail(ctx, DSP_ERR_BAD_ARG, "position (%d, %d, %d) appears twice in the batch (entry %llu): one chunk per position", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], (unsigned long long)i);
        if (rc != DSP_OK) { release_fresh_slots(ctx, fresh_slots); return rc; }
        if (fresh) fresh_slots.push_back(slots[i]);
    }
    const uint32_t* d_unused; uint32_t first_unused;
    int rc = upload_positions_and_slots(ctx, n, pos, slots.data(), &d_unused, &first_unused);
    if (rc != DSP_OK) { cudaStreamSynchronize(ctx->stream); release_fresh_slots(ctx, fresh_slots); return rc; }
    if ((rc = upload_by_runs(ctx, reinterpret_cast<uint8_t*>(ctx->d_voxels), dsp::kChunkBytes, slots.data(), n,
                             reinterpret_cast<const uint8_t*>(blocks))) != DSP_OK) return rc;
    if (palette_ids) {
        const size_t pal_n = palette_offsets[n];
        const size_t off_bytes = (n + 1) * sizeof(uint32_t), list_bytes = n * sizeof(uint32_t);
        const size_t pal_off = (off_bytes + list_bytes + 15) & ~size_t(15);
        if ((rc = ensure_scratch(ctx, pal_off + pal_n * sizeof(uint16_t))) != DSP_OK) return rc;
        uint32_t* d_off = reinterpret_cast<uint32_t*>(ctx->d_scratch);
        uint32_t* d_list = d_off + (n + 1);
        uint16_t* d_pal = reinterpret_cast<uint16_t*>(ctx->d_scratch + pal_off);
        DSP_CUDA(ctx, cudaMemcpyAsync(d_off, palette_offsets, off_bytes, cudaMemcpyHostToDevice, ctx->stream));
        DSP_CUDA(ctx, cudaMemcpyAsync(d_list, slots.data(), list_bytes, cudaMemcpyHostToDevice, ctx->stream));
        DSP_CUDA(ctx, cudaMemcpyAsync(d_pal, palette_ids, pal_n * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(n, static_cast<uint64_t>(ctx->num_sms) * 32));
        dsp::palette_remap_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_voxels, d_list, d_pal, d_off, static_cast<uint32_t>(n));
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `slots` (pageable) must outlive the copy
    }
    if (out_slots) std::memcpy(out_slots, slots.data(), n * sizeof(uint32_t));
    return DSP_OK;
}

int dsp_unload_chunks(dsp_ctx* ctx, uint64_t n, const uint32_t* slots) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (n && !slots) return fail(ctx, DSP_ERR_BAD_ARG, "slots is NULL");
    int rc = validate_slots(ctx, n, slots);
    if (rc != DSP_OK) return rc;
    if (!ctx->seam_slots.empty())
        for (uint64_t i = 0; i < n; ++i)
            if (ctx->seam_slots.count(slots[i])) return fail(ctx, DSP_ERR_BAD_ARG, "slot %u belongs to a region seam: seam chunks stay resident for the life of the context", slots[i]);
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // nothing in flight may still use the slots that are recycled
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t s = slots[i];
        if (!ctx->resident[s]) continue;  // duplicate in the list
        bool in_region = false;
        for (const Region& r : ctx->regions) {
            const uint64_t cnt = static_cast<uint64_t>(r.dims[0]) * r.dims[1] * r.dims[2];
            if (s >= r.first_slot && s < r.first_slot + cnt) { in_region = true; ctx->region_holes = true; break; }
        }
        if (!in_region) {
            ctx->map.erase(ctx->slot_pos[s]);
            ctx->free_slots.push_back(s);  // region slots are not recycled (their addresses are arithmetic)
        }
        ctx->resident[s] = 0;
        ctx->n_resident--;
        for (auto& hb : ctx->halo_batches) {  // its received slabs go with it (rare path: linear scan, then the device copy is refreshed)
            bool hit = false;
            for (uint32_t& v : hb.slots) if (v == s) { v = DSP_NO_SLOT; hit = true; }
            if (hit) DSP_CUDA(ctx, cudaMemcpy(ctx->d_halo_slots + hb.first_index, hb.slots.data(), hb.n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    ctx->nbr_dirty = true;
    return DSP_OK;
}

int dsp_find_chunk(const dsp_ctx* ctx, const int32_t pos[3], uint32_t* out_slot) {
    if (!ctx || !pos || !out_slot) return DSP_ERR_BAD_ARG;
    return find_slot(ctx, pos, out_slot) ? DSP_OK : DSP_ERR_NOT_RESIDENT;
}

uint64_t dsp_resident_chunks(const dsp_ctx* ctx) { return ctx ? ctx->n_resident : 0; }
uint64_t dsp_max_chunks(const dsp_ctx* ctx) { return ctx ? ctx->max_chunks : 0; }

int dsp_mesh_chunks_async(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, uint32_t mode, uint64_t arena_offset) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (n && !slots) return fail(ctx, DSP_ERR_BAD_ARG, "slots is NULL");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = validate_slots(ctx, n, slots);
    if (rc != DSP_OK) return rc;
    if (n > ctx->max_chunks) return fail(ctx, DSP_ERR_BAD_ARG, "n exceeds max_chunks");
    const uint32_t* d_list; uint32_t first;
    if ((rc = upload_positions_and_slots(ctx, n, nullptr, slots, &d_list, &first)) != DSP_OK) return rc;
    return enqueue_mesh(ctx, n, d_list, first, mode, arena_offset);
}

int dsp_mesh_range_async(dsp_ctx* ctx, uint32_t first_slot, uint64_t n, uint32_t mode, uint64_t arena_offset) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = validate_range(ctx, first_slot, n);
    if (rc != DSP_OK) return rc;
    return enqueue_mesh(ctx, n, nullptr, first_slot, mode, arena_offset);
}

int dsp_sync(dsp_ctx* ctx) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DSP_OK;
}

int dsp_mesh_result(dsp_ctx* ctx, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    return finish_mesh(ctx, out_entries, out_total_bytes);
}

int dsp_mesh_chunks(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, uint32_t mode, uint64_t arena_offset, dsp_mesh_entry* out_entries,
                    uint64_t* out_total_bytes) {
    int rc = dsp_mesh_chunks_async(ctx, n, slots, mode, arena_offset);
    if (rc != DSP_OK) return rc;
    return finish_mesh(ctx, out_entries, out_total_bytes);
}

int dsp_mesh_range(dsp_ctx* ctx, uint32_t first_slot, uint64_t n, uint32_t mode, uint64_t arena_offset, dsp_mesh_entry* out_entries,
                   uint64_t* out_total_bytes) {
    int rc = dsp_mesh_range_async(ctx, first_slot, n, mode, arena_offset);
    if (rc != DSP_OK) return rc;
    return finish_mesh(ctx, out_entries, out_total_bytes);
}

// Packed upload (csrc/host_pack.hpp): worker threads reduce every chunk to the smallest of five lossless forms in pinned staging,
// unpack_host_chunks_kernel expands them into the voxel store (and leaves the summaries of air / uniform chunks), the ordinary mesh
// kernel follows -- in a few pieces, so that the GPU works on piece k while the host packs piece k + 1.  The caller's memory is only
// ever read by the host: pageable memory works as well as pinned.  Same contract as the other host-chunk paths (a failing call
// leaves no trace).  *handled = false: the call is not one for this path (nothing has been done).
static int mesh_host_chunks_packed(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, const uint32_t* palette_lens, uint32_t mode,
                                   uint64_t arena_offset, uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes, void* host_vertices,
                                   uint64_t host_capacity, bool* handled) {
    *handled = false;
    // vertices to the host as well (dsp_mesh_host_chunks_to_host): only into pinned memory, where the device-driven download applies
    uint8_t* zero_copy_dst = nullptr;
    if (host_vertices != nullptr) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, host_vertices) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr) zero_copy_dst = static_cast<uint8_t*>(attr.devicePointer);
        else { cudaGetLastError(); return DSP_OK; }
    }
    const bool readback = zero_copy_dst != nullptr;
    if (!ctx->e2e_pack || n < 256 || n > (1u << 17) || (mode & (DSP_MESH_ORDERED | DSP_MESH_HALO))) return DSP_OK;
    if (!ctx->pack_pool) {
        int workers = ctx->host_threads;
        if (workers < 0) {
            cpu_set_t set;
            CPU_ZERO(&set);
            const int cpus = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : static_cast<int>(std::thread::hardware_concurrency());
            workers = std::min(15, std::max(0, cpus - 1));
        }
        if (workers < 1) { ctx->e2e_pack = 0; return DSP_OK; }
        ctx->pack_pool = new dsp_host::PackPool(workers);
    }
    if (n * dsp_host::kPackSlotBytes > ctx->h_pack_bytes) {
        if (ctx->h_pack) { DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); DSP_CUDA(ctx, cudaFreeHost(ctx->h_pack)); ctx->h_pack = nullptr; ctx->h_pack_bytes = 0; }
        const size_t cap = std::max<size_t>(n * dsp_host::kPackSlotBytes, 4096 * dsp_host::kPackSlotBytes);
        if (cudaMallocHost(&ctx->h_pack, cap) != cudaSuccess) { cudaGetLastError(); ctx->e2e_pack = 0; return DSP_OK; }  // no pinned memory to spare: the other paths
        ctx->h_pack_bytes = cap;
    }
    // How much of the work the host cores take.  With many cores everything is packed: 2.4x / 2.1x faster than the kernel pull with
    // 15 / 11 workers (one / two ranks per box).  With few -- eight ranks on a 32-CPU box, 3 workers each, all of them reading host
    // memory at once -- packing everything is slower than the pull (0.73-0.98 against 0.64-0.83 ms per call) and sharing the chunks out
    // between the cores and the bus is faster than either (0.42-0.52 ms): of every 16 chunks `keep` are packed, the others cross the
    // bus as they are, fetched by the unpack kernel from the caller's array.  That needs an array the device can address; pageable
    // memory is packed whole (the alternative there is a pageable copy).  tools/e2e_pack_sweep.py has the sweep.
    uint32_t keep_of_16 = 16;
    const uint8_t* dev_blocks = nullptr;
    if (ctx->pack_pool->workers() <= 8 || ctx->e2e_pack_keep > 0) {
        cudaPointerAttributes a0{}, a1{};
        const uint8_t* last = reinterpret_cast<const uint8_t*>(blocks) + n * dsp::kChunkBytes - 1;
        if (cudaPointerGetAttributes(&a0, blocks) == cudaSuccess && a0.type == cudaMemoryTypeHost && a0.devicePointer != nullptr &&
            cudaPointerGetAttributes(&a1, last) == cudaSuccess && a1.type == cudaMemoryTypeHost && a1.devicePointer != nullptr &&
            static_cast<const uint8_t*>(a1.devicePointer) - static_cast<const uint8_t*>(a0.devicePointer) == last - reinterpret_cast<const uint8_t*>(blocks))
            dev_blocks = static_cast<const uint8_t*>(a0.devicePointer);
        else cudaGetLastError();
        if (dev_blocks != nullptr && ctx->e2e_pack < 2) keep_of_16 = ctx->pack_pool->workers() <= 4 ? 8u : 12u;
    }
    if (ctx->e2e_pack_keep > 0) keep_of_16 = dev_blocks != nullptr ? static_cast<uint32_t>(std::min(16, ctx->e2e_pack_keep)) : 16u;  // DSP_E2E_PACK_KEEP (A/B runs)
    *handled = true;
    ctx->last_mesh_n = n;
    const bool trace = getenv("DSP_TRACE") != nullptr;
    const auto tr0 = std::chrono::steady_clock::now();
    // (DSP_TRACE: time stamps are kept in memory and printed when the call is over -- a print in flight is longer than the steps it times)
    struct TraceLine { const char* what; double us; };
    TraceLine trace_lines[64];
    int n_trace_lines = 0;
    auto tr = [&](const char* what) {
        if (trace && n_trace_lines < 64) trace_lines[n_trace_lines++] = TraceLine{what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr0).count()};
    };
    int rc;
    uint8_t* h;
    if ((rc = stage_host(ctx, n * (sizeof(int4) + 2 * sizeof(uint32_t)), &h)) != DSP_OK) return rc;
    int4* h_pos = reinterpret_cast<int4*>(h);
    uint32_t* h_slots = reinterpret_cast<uint32_t*>(h + n * sizeof(int4));
    uint32_t* h_desc = h_slots + n;
    dsp_host::PackPool& pool = *ctx->pack_pool;
    pool.start(blocks, palette_lens, n, h_desc, ctx->h_pack, keep_of_16);  // the workers are off; this thread resolves positions to slots meanwhile
    tr("packing started");
    bool persist_launched_flag = false;
    std::vector<uint32_t> fresh_slots;
    unsigned long long* d_cursors = nullptr;  // (download) the arena cursor behind every piece
    int last_piece = -1;
    const uint64_t room = (ctx->arena_limit ? ctx->arena_limit : ctx->arena_bytes) - arena_offset;
    bool* const persist_launched_p = &persist_launched_flag;
    auto bail = [&](int code) {
        const std::string msg = ctx->err;
        pool.finish();  // nobody may still be reading the caller's voxels or writing the staging buffers
        if (*persist_launched_p)  // a mesh launch is waiting for the rest of the batch: let it run out (the call fails, its output is void)
            dsp::release_all_pieces_kernel<<<1, 32, 0, ctx->copy_stream>>>(ctx->d_tickets + 32 * (dsp_ctx::kMaxSub - 1));
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->copy_stream2);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
        release_fresh_slots(ctx, fresh_slots);
        stage_release(ctx);
        ctx->err = msg;
        return code;
    };
    if (cudaMemsetAsync(ctx->d_ctl, 0, 32, ctx->stream) != cudaSuccess || cudaMemsetAsync(ctx->d_tickets, 0, dsp_ctx::kMaxSub * 128, ctx->stream) != cudaSuccess)
        return bail(fail(ctx, DSP_ERR_CUDA, "control block reset failed: %s", cudaGetErrorString(cudaGetLastError())));
    ctx->ctl_clean = false;  // the pieces share the cursor and leave the block as it is
    ctx->result_in_tail = false;
    if (readback) {
        if ((rc = ensure_scratch(ctx, dsp_ctx::kMaxSub * sizeof(unsigned long long))) != DSP_OK) return bail(rc);
        d_cursors = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
    }
    // the side stream must not overtake earlier work of the main stream that still reads the voxel store
    if (cudaEventRecord(ctx->ev_done, ctx->stream) != cudaSuccess || cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done, 0) != cudaSuccess ||
        cudaStreamWaitEvent(ctx->copy_stream2, ctx->ev_done, 0) != cudaSuccess)
        return bail(fail(ctx, DSP_ERR_CUDA, "event record / wait failed"));
    new_stamp(ctx);
    tr("streams set up");
    // Chunk-local parity calls without a download: ONE mesh launch for the whole batch, started behind the first unpack kernel.  Its
    // CTAs take tickets in order and wait (thread 0, bounded) until the piece a ticket belongs to has been released: the last warp
    // of every unpack kernel sets the piece's word.  A mesh launch per piece costs ~15 us of ramp, tail and control-block epilogue
    // each time -- three of them took 118-124 us for work that is 72 us in one.  The launch leaves one CTA slot per SM free so that
    // the unpack kernels always find room beside its waiting CTAs; the unpack kernels (a handful of PCIe round trips each, ~13 us
    // whatever their size) alternate between two side streams so that their latencies overlap.  Pieces are of equal size here.
    // (batches beyond 16,384 chunks keep a launch per piece: its fixed cost no longer shows there, and that launch runs three CTAs per SM)
    const bool persist = ctx->e2e_pack_persist != 0 && !readback && mode == 0u && (n <= 16384 || ctx->e2e_pack_persist > 1);
    uint32_t* const d_released = ctx->d_tickets + 32 * (dsp_ctx::kMaxSub - 1);  // 32 words, one per piece (a ticket line no piece uses; zeroed with the others above)
    uint32_t* const d_done_warps = ctx->d_tickets + 32 * (dsp_ctx::kMaxSub - 2);  // 32 counters
    uint64_t persist_piece_chunks = 0;
    // pieces: a small first one (the GPU starts while most of the packing is still to do), the rest even
    std::vector<uint64_t> piece_begin;
    if (persist) {
        const uint64_t pieces = std::min<uint64_t>(static_cast<uint64_t>(std::min(32, ctx->e2e_pack_persist_pieces)), std::max<uint64_t>(1, n / 256));
        persist_piece_chunks = (((n + pieces - 1) / pieces) + 15) & ~uint64_t(15);
        for (uint64_t b = 0; b < n; b += persist_piece_chunks) piece_begin.push_back(b);
        piece_begin.push_back(n);
    } else {
        // (unit quads on their way to the host are 12x the voxels' bytes: the download bounds the call and wants finer pieces to hide behind)
        const uint64_t want = readback && !(mode & DSP_MESH_GREEDY) ? std::max(8, ctx->e2e_pack_pieces) : ctx->e2e_pack_pieces;
        const int pieces = static_cast<int>(std::min<uint64_t>(want, std::max<uint64_t>(1, n / 256)));
        piece_begin.push_back(0);
        if (pieces > 1) {
            const uint64_t head = ((n / (2 * static_cast<uint64_t>(pieces))) + 15) & ~uint64_t(15);
            piece_begin.push_back(head);
            for (int k = 1; k < pieces - 1; ++k) piece_begin.push_back((head + (n - head) * static_cast<uint64_t>(k) / static_cast<uint64_t>(pieces - 1) + 15) & ~uint64_t(15));
        }
        piece_begin.push_back(n);
    }
    uint64_t resolved = 0;
    std::vector<cudaEvent_t> tev;  // DSP_TRACE: device-side timeline (start, then after every kernel)
    static std::vector<cudaEvent_t> tev_pool;  // (created once: cudaEventCreate in flight would be most of what the trace shows)
    if (trace && tev_pool.empty()) { tev_pool.resize(40); for (cudaEvent_t& e : tev_pool) cudaEventCreate(&e); }
    auto tev_mark = [&] { if (trace && tev.size() < tev_pool.size()) { cudaEvent_t e = tev_pool[tev.size()]; cudaEventRecord(e, ctx->stream); tev.push_back(e); } };
    tev_mark();
    bool all_consecutive = true;
    if (persist) {  // the one launch needs to know up front whether the batch is a run of consecutive slots
        if ((rc = resolve_positions(ctx, pos, 0, n, h_slots, h_pos, fresh_slots)) != DSP_OK) return bail(rc);
        resolved = n;
        for (uint64_t i = 1; i < n && all_consecutive; ++i) all_consecutive = h_slots[i] == h_slots[i - 1] + 1;
    }
    for (size_t j = 0; j + 1 < piece_begin.size(); ++j) {
        const uint64_t i0 = piece_begin[j], i1 = piece_begin[j + 1], m = i1 - i0;
        if (m == 0) continue;
        bool consecutive = true;
        if (resolved < i1) {
            if ((rc = resolve_positions(ctx, pos, resolved, i1, h_slots, h_pos, fresh_slots)) != DSP_OK) return bail(rc);
            resolved = i1;
        }
        for (uint64_t i = i0 + 1; i < i1 && consecutive; ++i) consecutive = h_slots[i] == h_slots[i - 1] + 1;
        tr("piece resolved");
        if (j + 2 == piece_begin.size()) pool.help();  // last piece: nothing else for this thread to do but pack along
        while (pool.packed_upto() < i1) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        tr("piece packed");
        // the expansion of piece j runs on the side stream, next to the mesh kernel of piece j - 1 (it is a handful of PCIe round
        // trips, not bandwidth); the mesh kernel of piece j waits for it through an event
        dsp::unpack_host_chunks_kernel<<<static_cast<unsigned>((m + dsp::kUnpackChunksPerCta - 1) / dsp::kUnpackChunksPerCta), dsp::kUnpackChunksPerCta * 32, 0,
                                         (persist && (j & 1)) ? ctx->copy_stream2 : ctx->copy_stream>>>(
            ctx->d_pos, ctx->d_slots + i0, ctx->d_summary, ctx->d_uid, ctx->d_bvalid, ctx->d_voxels, h_pos + i0, h_slots + i0, h_desc + i0, ctx->h_pack + i0 * dsp_host::kPackSlotBytes,
            dev_blocks != nullptr ? dev_blocks + i0 * dsp::kChunkBytes : nullptr, static_cast<uint32_t>(m), persist ? d_done_warps + j : nullptr, persist ? d_released + j : nullptr);
        if (persist) {
            if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, DSP_ERR_CUDA, "unpack kernel launch failed"));
            ctx->bmask_stale = true;
            ctx->launches++;
            tr("unpack launched");
            if (!persist_launched_flag) {
                persist_launched_flag = true;  // (set first: a launch that fails half way must still be released by bail)
                if ((rc = launch_mesh(ctx, n, all_consecutive ? nullptr : ctx->d_slots, h_slots[0], mode, arena_offset, 0, 0, false, false, nullptr, d_released,
                                      static_cast<uint32_t>(persist_piece_chunks))) != DSP_OK) return bail(rc);
            }
            tr("piece enqueued");
            continue;
        }
        if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ctx->ev_sub[j], ctx->copy_stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->ev_sub[j], 0) != cudaSuccess)
            return bail(fail(ctx, DSP_ERR_CUDA, "unpack kernel launch failed"));
        ctx->bmask_stale = true;
        ctx->launches++;
        tev_mark();
        tr("unpack launched");
        if ((rc = launch_mesh(ctx, m, consecutive ? nullptr : ctx->d_slots + i0, h_slots[i0], mode, arena_offset, i0, static_cast<int>(j), false, false, nullptr)) != DSP_OK) return bail(rc);
        tev_mark();
        if (readback) {
            // this piece's vertices = the arena bytes between the cursor behind the previous piece and the cursor now: the device records
            // the cursor, a copy kernel on the third stream, ordered behind the piece by an event, stores the window into host memory
            dsp::publish_u64_kernel<<<1, 1, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(ctx->d_ctl + 16), d_cursors + j);
            if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ctx->ev_meshed[j], ctx->stream) != cudaSuccess ||
                cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_meshed[j], 0) != cudaSuccess)
                return bail(fail(ctx, DSP_ERR_CUDA, "event record / wait failed in the download pipeline"));
            dsp::copy_window_to_host_kernel<<<ctx->num_sms * 2, 256, 0, ctx->d2h_stream>>>(ctx->d_arena + arena_offset, zero_copy_dst, last_piece >= 0 ? d_cursors + last_piece : nullptr,
                                                                                           d_cursors + j, std::min<uint64_t>(room, host_capacity));
            if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, DSP_ERR_CUDA, "copy_window_to_host_kernel launch failed"));
            ctx->launches += 2;
            last_piece = static_cast<int>(j);
        }
        tr("piece enqueued");
    }
    pool.finish();
    tr("pool finished");
    if (out_slots) std::memcpy(out_slots, h_slots, n * sizeof(uint32_t));
    {
        static const uint64_t form_bytes[6] = {0, 0, 1024, 2048, 4096, 8192};
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t form = std::min<uint32_t>(h_desc[i] & 7u, 5u);  // (chunks left in place count as raw: they cross the bus as they are)
            ctx->upload_stats[form]++;
            ctx->upload_stats[7] += form_bytes[form] + sizeof(int4) + 2 * sizeof(uint32_t);
        }
        ctx->upload_stats[6] = static_cast<uint64_t>(pool.workers());
    }
    if ((rc = stage_release(ctx)) != DSP_OK) return rc;
    uint64_t total = 0;
    tr("all enqueued");
    rc = finish_mesh(ctx, out_entries, &total, 0);
    if (out_total_bytes) *out_total_bytes = total;
    if (readback) cudaStreamSynchronize(ctx->d2h_stream);
    if (rc != DSP_OK) {  // arena too small: the call did not happen -- the slots it created are given back (the bytes needed stay in *out_total_bytes)
        const std::string msg = ctx->err;
        release_fresh_slots(ctx, fresh_slots);
        ctx->err = msg;
        return rc;
    }
    if (readback && total > host_capacity)
        rc = fail(ctx, DSP_ERR_BAD_ARG, "host vertex buffer too small: %llu bytes needed, %llu given (what fits was copied)", (unsigned long long)total, (unsigned long long)host_capacity);
    tr("finished");
    if (trace) {
        for (int k = 0; k < n_trace_lines; ++k) fprintf(stderr, "[trace] %-28s %8.1f us\n", trace_lines[k].what, trace_lines[k].us);
        for (size_t k = 1; k < tev.size(); ++k) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tev[0], tev[k - 1]); cudaEventElapsedTime(&b, tev[0], tev[k]);
            fprintf(stderr, "[trace] device: %s %zu  %7.1f -> %7.1f us after the first event\n", (k & 1) ? "unpack" : "mesh  ", (k - 1) / 2, a * 1e3, b * 1e3);
        }
    }
    return rc;
}

// dsp_mesh_host_chunks, optionally followed (pipelined) by the device->host copy of the vertices: host_vertices != NULL
// mirrors the arena window [arena_offset, arena_offset + total) into host memory (entry offsets minus arena_offset index it).
static int mesh_host_chunks_impl(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, const uint32_t* palette_lens, uint32_t mode, uint64_t arena_offset,
                                 uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes, void* host_vertices, uint64_t host_capacity) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (out_total_bytes) *out_total_bytes = 0;
    if (n && (!pos || !blocks)) return fail(ctx, DSP_ERR_BAD_ARG, "pos/blocks is NULL");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = check_mesh_args(ctx, n, mode, arena_offset);
    if (rc != DSP_OK) return rc;
    const bool readback = host_vertices != nullptr;
    std::memset(ctx->upload_stats, 0, sizeof(ctx->upload_stats));
    // Entry table in pinned host memory: the entries of every sub-batch but the last leave on the download stream as soon as the
    // sub-batch is meshed (the D2H engine is idle while the voxels come up), so only the last piece's few KB remain for the tail.
    bool early_entries = false;
    if (!readback && out_entries && n > 1024) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, out_entries) == cudaSuccess && attr.type == cudaMemoryTypeHost) early_entries = true;
        else cudaGetLastError();
    }
    if ((readback || early_entries) && !ctx->d2h_stream) {
        DSP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < dsp_ctx::kMaxSub; ++i) DSP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_meshed[i], cudaEventDisableTiming));
    }
    if (n) {  // bulk calls: pack on the host, expand on the device (the bus carries a fraction of the voxels' bytes)
        bool handled = false;
        rc = mesh_host_chunks_packed(ctx, n, pos, blocks, palette_lens, mode, arena_offset, out_slots, out_entries, out_total_bytes, host_vertices, host_capacity, &handled);
        if (handled || rc != DSP_OK) return rc;
    }
    if (mode & (DSP_MESH_ORDERED | DSP_MESH_HALO)) {  // these need the whole batch resident first: plain load, then mesh
        std::vector<uint32_t> slots(n);
        if ((rc = dsp_load_chunks(ctx, n, pos, blocks, nullptr, nullptr, slots.data())) != DSP_OK) return rc;
        if (out_slots) std::memcpy(out_slots, slots.data(), n * sizeof(uint32_t));
        uint64_t total = 0;
        if ((rc = dsp_mesh_chunks(ctx, n, slots.data(), mode, arena_offset, out_entries, &total)) != DSP_OK) return rc;
        if (out_total_bytes) *out_total_bytes = total;
        if (readback) {
            if (total > host_capacity) return fail(ctx, DSP_ERR_BAD_ARG, "host vertex buffer too small: %llu bytes needed, %llu given", (unsigned long long)total, (unsigned long long)host_capacity);
            return total ? dsp_download_arena(ctx, arena_offset, total, host_vertices) : DSP_OK;
        }
        return DSP_OK;
    }
    ctx->last_mesh_n = n;
    const bool trace = getenv("DSP_TRACE") != nullptr;
    const auto tr0 = std::chrono::steady_clock::now();
    auto tr = [&](const char* what) { if (trace) fprintf(stderr, "[trace] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr0).count()); };
    // Are the voxels in pinned / registered host memory that the device can read?  Then no copy engine is involved at all: the mesh
    // kernel pulls every tile over PCIe with the bulk (TMA) copy that otherwise reads the voxel store, meshes it and files it in the
    // store on the way (MeshParams::host_tiles).  Chunks whose palette has one entry are never fetched (stage_host_chunks_kernel).
    // Pageable memory takes the copy-engine pipeline below.
    // Which is faster depends on how many tiles need not cross the bus: SM pulls reach 50.7 GB/s, the copy engine 55.1 GB/s
    // (tools/pcie_pull_probe.cu), so the pull is taken when the palette lengths spare at least a tenth of the chunks
    // (DSP_E2E_PULL=2 forces it, 0 switches it off).
    const uint16_t* dev_blocks = nullptr;
    if (ctx->e2e_pull && n && (!readback || ctx->e2e_pull >= 2)) {  // (with the download behind it the copy-engine pipeline measured faster: 1.09 against 1.23 ms)
        uint64_t n_air = 0;
        if (palette_lens != nullptr)
            for (uint64_t i = 0; i < n; ++i) n_air += palette_lens[i] == 1u;
        if (ctx->e2e_pull >= 2 || n_air * 10 >= n) {
            // first and last byte of the array must lie in one device-addressable host allocation
            cudaPointerAttributes a0{}, a1{};
            const uint8_t* last = reinterpret_cast<const uint8_t*>(blocks) + n * dsp::kChunkBytes - 1;
            if (cudaPointerGetAttributes(&a0, blocks) == cudaSuccess && a0.type == cudaMemoryTypeHost && a0.devicePointer != nullptr &&
                cudaPointerGetAttributes(&a1, last) == cudaSuccess && a1.type == cudaMemoryTypeHost && a1.devicePointer != nullptr &&
                static_cast<const uint8_t*>(a1.devicePointer) - static_cast<const uint8_t*>(a0.devicePointer) == last - reinterpret_cast<const uint8_t*>(blocks))
                dev_blocks = static_cast<const uint16_t*>(a0.devicePointer);
            else cudaGetLastError();
        }
    }
    const bool pull = dev_blocks != nullptr;
    // sub-batches: copy k+1 crosses PCIe while sub-batch k is meshed
    // Sub-batch plan.  Every copy costs a few microseconds of gap on the DMA engine, so the body is cut into pieces of about
    // 1,024 chunks (8 MiB); the LAST piece is small (256 chunks) because its kernel and the entry read-back are the only work
    // that cannot hide behind a copy.
    // The FIRST piece is small too (128 chunks): the host resolves a piece's positions to slots before it can enqueue the
    // piece's copy, and nothing crosses PCIe while it does that for the first one.
    std::vector<uint64_t> sub_begin;
    if (pull && !readback) {  // (with a download behind it the finer plan below keeps the three directions overlapped)
        // Kernel pull: a kernel boundary is a PCIe bubble (the next launch's CTAs wait for the previous one to drain), so there
        // are only two pieces -- a first fifth that starts pulling while the host still resolves the slots of the rest (the host
        // resolves a chunk in about a quarter of the time an average tile takes to cross the bus), then everything else in one launch.
        sub_begin.push_back(0);
        if (n > 1024 && ctx->e2e_pull_pieces >= 3) sub_begin.push_back((n / (4 * ctx->e2e_pull_first_div) + 63) & ~uint64_t(63));
        if (n > 1024) sub_begin.push_back((n / ctx->e2e_pull_first_div + 63) & ~uint64_t(63));
        sub_begin.push_back(n);
    } else {
        const uint64_t tail = n > 512 ? 256 : 0, head = n > 1024 ? ctx->e2e_head_chunks : 0, body = n - tail - head;
        uint64_t pieces = std::max<uint64_t>(1, (body + 1023) / 1024);
        pieces = std::min<uint64_t>(pieces, dsp_ctx::kMaxSub - 2);
        if (head) sub_begin.push_back(0);
        for (uint64_t k = 0; k < pieces; ++k) sub_begin.push_back(head + body * k / pieces);
        if (tail) sub_begin.push_back(head + body);
        sub_begin.push_back(n);
    }
    const int subs = static_cast<int>(sub_begin.size()) - 1;
    DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_ctl, 0, 32, ctx->stream));
    DSP_CUDA(ctx, cudaMemsetAsync(ctx->d_tickets, 0, dsp_ctx::kMaxSub * 128, ctx->stream));
    ctx->ctl_clean = false;  // the sub-batch launches share the cursor and leave the block as it is
    ctx->result_in_tail = false;
    if (n == 0) return finish_mesh(ctx, out_entries, out_total_bytes);
    uint8_t* h;
    if ((rc = stage_host(ctx, n * (sizeof(int4) + sizeof(uint32_t) + 1), &h)) != DSP_OK) return rc;
    int4* h_pos = reinterpret_cast<int4*>(h);
    uint32_t* h_slots = reinterpret_cast<uint32_t*>(h + n * sizeof(int4));
    uint8_t* h_air = h + n * (sizeof(int4) + sizeof(uint32_t));
    tr("staging ready");
    // the copy stream must not overtake earlier work of the main stream that still reads the voxel store
    DSP_CUDA(ctx, cudaEventRecord(ctx->ev_done, ctx->stream));
    DSP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done, 0));
    cudaStream_t main_stream = ctx->stream;
    std::vector<uint32_t> fresh_slots;
    new_stamp(ctx);
    // A failure part-way leaves no trace: earlier sub-batches are drained, the slots this call created are given back and the
    // staging block is released (the reference's load loop has no partial state either: it panics, chunk_mesh.rs:124).
    auto bail = [&](int code) {
        const std::string msg = ctx->err;
        ctx->stream = main_stream;
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
        release_fresh_slots(ctx, fresh_slots);
        stage_release(ctx);
        ctx->err = msg;
        return code;
    };
    uint64_t* const h_cursor = reinterpret_cast<uint64_t*>(ctx->h_pinned + 512);  // kMaxSub x u64 of the 4 KiB pinned block
    // Is the destination pinned (device-accessible) host memory?  Then the download needs no host in the loop (see below).
    uint8_t* zero_copy_dst = nullptr;
    unsigned long long* d_cursors = nullptr;
    if (readback) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, host_vertices) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr) {
            zero_copy_dst = static_cast<uint8_t*>(attr.devicePointer);
            if ((rc = ensure_scratch(ctx, dsp_ctx::kMaxSub * sizeof(unsigned long long))) != DSP_OK) return rc;
            d_cursors = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
        } else {
            cudaGetLastError();
        }
    }
    int last_sub = -1;
    uint64_t copied_to = 0, entries_out = 0;
    const uint64_t room = (ctx->arena_limit ? ctx->arena_limit : ctx->arena_bytes) - arena_offset;
    auto copy_out = [&](int j) -> int {  // vertices of sub-batch j: arena [copied_to, cursor(j)) -> host, on the third stream
        DSP_CUDA(ctx, cudaEventSynchronize(ctx->ev_meshed[j]));
        const uint64_t hi = std::min<uint64_t>(std::min<uint64_t>(h_cursor[j], room), host_capacity);
        if (hi > copied_to) {
            DSP_CUDA(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(host_vertices) + copied_to, ctx->d_arena + arena_offset + copied_to, hi - copied_to, cudaMemcpyDeviceToHost, ctx->d2h_stream));
            copied_to = hi;
        }
        return DSP_OK;
    };
    for (int j = 0; j < subs; ++j) {
        const uint64_t i0 = sub_begin[j], m = sub_begin[j + 1] - i0;
        if (m == 0) continue;
        bool consecutive = true;
        // (a position twice in the batch is refused: the copy of a later sub-batch would overwrite a slot an earlier one is still meshing)
        if ((rc = resolve_positions(ctx, pos, i0, i0 + m, h_slots, h_pos, fresh_slots)) != DSP_OK) return bail(rc);
        for (uint64_t i = 0; i < m; ++i) {
            h_air[i0 + i] = (pull && palette_lens != nullptr && palette_lens[i0 + i] == 1u) ? 1 : 0;
            if (i && h_slots[i0 + i] != h_slots[i0 + i - 1] + 1) consecutive = false;
        }
        // The copy stream carries nothing but the voxel tiles, back to back: small copies between them cost more in gaps than
        // in bytes (16 copies took 0.68 ms where 32 MiB alone take 0.61 ms), and small copies on ANOTHER stream wait in the
        // DMA engine until all tile copies are through (0.73 -> 0.93 ms).  So positions and the slot list do not use a copy
        // engine at all: a small kernel on the main stream reads them straight from the pinned staging block (UVA).
        // Kernel pull: the second launch goes to the side stream.  Behind the first one on the same stream its CTAs would wait
        // (griddepcontrol.wait) until the LAST tile of the first launch has crossed the bus and been meshed -- a PCIe bubble;
        // on its own stream each of its CTAs starts pulling the moment a CTA of the first launch retires.
        const bool side = pull && !readback && j >= 1 && ctx->e2e_pull_pieces == 2;
        if (side) ctx->stream = ctx->copy_stream;
        if (pull && palette_lens != nullptr)
            dsp::stage_host_chunks_kernel<<<static_cast<unsigned>((m + dsp::kStageChunksPerCta - 1) / dsp::kStageChunksPerCta), 256, 0, ctx->stream>>>(ctx->d_pos, ctx->d_slots + i0, ctx->d_summary, ctx->d_bvalid, ctx->d_voxels,
                                                                                             h_pos + i0, h_slots + i0, h_air + i0, static_cast<uint32_t>(m));
        else
            dsp::stage_positions_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, ctx->stream>>>(
                ctx->d_pos, ctx->d_slots + i0, ctx->d_summary, ctx->d_bvalid, h_pos + i0, h_slots + i0, static_cast<uint32_t>(m));
        ctx->bmask_stale = true;
        if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, DSP_ERR_CUDA, "staging kernel launch failed"));
        ctx->launches++;
        const uint32_t* d_list = consecutive ? nullptr : ctx->d_slots + i0;
        if (!pull) {
            ctx->stream = ctx->copy_stream;  // upload_by_runs enqueues on ctx->stream
            rc = upload_by_runs(ctx, reinterpret_cast<uint8_t*>(ctx->d_voxels), dsp::kChunkBytes, h_slots + i0, m,
                                reinterpret_cast<const uint8_t*>(blocks + (i0 * dsp::kChunkVoxels)));
            ctx->stream = main_stream;
            if (rc != DSP_OK) return bail(rc);
            if (cudaEventRecord(ctx->ev_sub[j], ctx->copy_stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->ev_sub[j], 0) != cudaSuccess)
                return bail(fail(ctx, DSP_ERR_CUDA, "event record / wait failed in the copy pipeline"));
        }
        rc = launch_mesh(ctx, m, d_list, h_slots[i0], mode, arena_offset, i0, j, false, false, pull ? dev_blocks + i0 * dsp::kChunkVoxels : nullptr);
        if (side) {
            ctx->stream = main_stream;
            if (rc == DSP_OK && (cudaEventRecord(ctx->ev_sub[j], ctx->copy_stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->ev_sub[j], 0) != cudaSuccess))
                rc = fail(ctx, DSP_ERR_CUDA, "event record / wait failed behind the side-stream launch");
        }
        if (rc != DSP_OK) return bail(rc);
        tr("sub-batch enqueued");
        if (early_entries && j + 1 < subs) {
            if (cudaEventRecord(ctx->ev_meshed[j], ctx->stream) != cudaSuccess || cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_meshed[j], 0) != cudaSuccess ||
                cudaMemcpyAsync(out_entries + i0, ctx->d_entries + i0, m * sizeof(dsp_mesh_entry), cudaMemcpyDeviceToHost, ctx->d2h_stream) != cudaSuccess)
                return bail(fail(ctx, DSP_ERR_CUDA, "entry download failed in the copy pipeline"));
            entries_out = i0 + m;
        }
        if (readback) {
            // where the cursor stands after this sub-batch: its vertices are the arena bytes [cursor(j-1), cursor(j))
            if (zero_copy_dst != nullptr) {
                // pinned destination: the device keeps the cursor history itself and a copy kernel on the third stream, ordered
                // behind this sub-batch by an event, stores the window into host memory -- the host enqueues and moves on
                dsp::publish_u64_kernel<<<1, 1, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(ctx->d_ctl + 16), d_cursors + j);
                if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ctx->ev_meshed[j], ctx->stream) != cudaSuccess ||
                    cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_meshed[j], 0) != cudaSuccess)
                    return bail(fail(ctx, DSP_ERR_CUDA, "event record / wait failed in the download pipeline"));
                dsp::copy_window_to_host_kernel<<<ctx->num_sms * 2, 256, 0, ctx->d2h_stream>>>(ctx->d_arena + arena_offset, zero_copy_dst, last_sub >= 0 ? d_cursors + last_sub : nullptr,
                                                                                               d_cursors + j, std::min<uint64_t>(room, host_capacity));
                if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, DSP_ERR_CUDA, "copy_window_to_host_kernel launch failed"));
                ctx->launches += 2;
                last_sub = j;
                continue;
            }
            // pageable destination: the copy engine needs sizes on the host -- read the cursor back, wait for the previous sub-batch
            dsp::publish_u64_kernel<<<1, 1, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(ctx->d_ctl + 16), reinterpret_cast<unsigned long long*>(h_cursor + j));
            if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ctx->ev_meshed[j], ctx->stream) != cudaSuccess)
                return bail(fail(ctx, DSP_ERR_CUDA, "cursor read-back failed in the copy pipeline"));
            // ... and the PREVIOUS sub-batch's vertices leave for the host now: its kernel has finished (this wait is short, the
            // copy of sub-batch j is what the host would otherwise idle behind), the next upload is already in the DMA queue
            if (last_sub >= 0 && (rc = copy_out(last_sub)) != DSP_OK) return bail(rc);
            last_sub = j;
        }
    }
    if (readback && zero_copy_dst == nullptr && last_sub >= 0 && (rc = copy_out(last_sub)) != DSP_OK) return bail(rc);
    if (out_slots) std::memcpy(out_slots, h_slots, n * sizeof(uint32_t));
    if ((rc = stage_release(ctx)) != DSP_OK) return rc;
    if (trace) { cudaStreamSynchronize(ctx->copy_stream); tr("copies done"); }
    uint64_t total = 0;
    rc = finish_mesh(ctx, out_entries, &total, entries_out);
    if (out_total_bytes) *out_total_bytes = total;
    if (rc != DSP_OK) {  // arena too small, seam time-out: the call did not happen -- the slots it created are given back (the bytes needed stay in *out_total_bytes)
        const std::string msg = ctx->err;
        cudaStreamSynchronize(ctx->copy_stream);
        if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
        release_fresh_slots(ctx, fresh_slots);
        ctx->err = msg;
        return rc;
    }
    if (early_entries) DSP_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));  // (long finished: those copies ran under the uploads)
    if (readback) {
        DSP_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
        if (rc == DSP_OK && total > host_capacity)
            return fail(ctx, DSP_ERR_BAD_ARG, "host vertex buffer too small: %llu bytes needed, %llu given (what fits was copied)", (unsigned long long)total, (unsigned long long)host_capacity);
    }
    tr("finished");
    return rc;
}

int dsp_mesh_host_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, uint32_t mode, uint64_t arena_offset,
                         uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes) {
    return mesh_host_chunks_impl(ctx, n, pos, blocks, nullptr, mode, arena_offset, out_slots, out_entries, out_total_bytes, nullptr, 0);
}

int dsp_mesh_host_chunks_ex(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, const uint32_t* palette_lens, uint32_t mode, uint64_t arena_offset,
                            uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes, void* host_vertices, uint64_t host_capacity) {
    return mesh_host_chunks_impl(ctx, n, pos, blocks, palette_lens, mode, arena_offset, out_slots, out_entries, out_total_bytes, host_vertices, host_capacity);
}

int dsp_mesh_host_chunks_to_host(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, uint32_t mode, uint64_t arena_offset,
                                 uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes, void* host_vertices, uint64_t host_capacity) {
    if (ctx && !host_vertices) return fail(ctx, DSP_ERR_BAD_ARG, "host_vertices is NULL");
    return mesh_host_chunks_impl(ctx, n, pos, blocks, nullptr, mode, arena_offset, out_slots, out_entries, out_total_bytes, host_vertices, host_capacity);
}

int dsp_apply_edits(dsp_ctx* ctx, uint64_t n, const dsp_edit* edits, uint32_t* out_dirty_slots, uint64_t cap_dirty, uint64_t* out_n_dirty) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (out_n_dirty) *out_n_dirty = 0;
    if (n == 0) return DSP_OK;
    if (!edits) return fail(ctx, DSP_ERR_BAD_ARG, "edits is NULL");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint64_t i = 0; i < n; ++i)
        if (edits[i].block > 0xFFFFu) return fail(ctx, DSP_ERR_BAD_ARG, "edit %llu: block id %u does not fit u16", (unsigned long long)i, edits[i].block);
    {
        bool handled = false;
        uint32_t nd = 0;
        int rc = apply_edits_device(ctx, n, edits, false, &handled, &nd);
        if (rc != DSP_OK) return rc;
        if (handled) {
            if (out_n_dirty) *out_n_dirty = nd;
            if (out_dirty_slots) {
                if (nd > cap_dirty) return fail(ctx, DSP_ERR_BAD_ARG, "dirty list needs %u entries, capacity %llu", nd, (unsigned long long)cap_dirty);
                if (nd) DSP_CUDA(ctx, cudaMemcpy(out_dirty_slots, ctx->d_slots, nd * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            }
            return DSP_OK;
        }
    }
    // host path (chunks addressed through the position map): world.rs:69-74 to_chunk_coord: chunk = div_euclid(w, 16), local = rem_euclid(w, 16)
    std::unordered_map<uint64_t, uint32_t> last;  // (slot << 12 | voxel index) -> position in `packed`
    std::vector<uint2> packed;
    packed.reserve(n);
    std::vector<uint32_t> dirty;
    PosKey cached{0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
    uint32_t cached_slot = DSP_NO_SLOT;
    for (uint64_t i = 0; i < n; ++i) {
        const dsp_edit& e = edits[i];
        if (e.block > 0xFFFFu) return fail(ctx, DSP_ERR_BAD_ARG, "edit %llu: block id %u does not fit u16", (unsigned long long)i, e.block);
        const int32_t c[3] = {e.wx >> 4, e.wy >> 4, e.wz >> 4};
        uint32_t slot;
        if (cached.x == c[0] && cached.y == c[1] && cached.z == c[2]) slot = cached_slot;
        else { if (!find_slot(ctx, c, &slot)) slot = DSP_NO_SLOT; cached = PosKey{c[0], c[1], c[2]}; cached_slot = slot; }
        if (slot == DSP_NO_SLOT) continue;
        const uint32_t idx = static_cast<uint32_t>(e.wx & 15) + 16u * static_cast<uint32_t>(e.wy & 15) + 256u * static_cast<uint32_t>(e.wz & 15);
        const uint64_t key = (static_cast<uint64_t>(slot) << 12) | idx;
        const uint2 rec = make_uint2(slot, idx | (e.block << 16));
        auto it = last.find(key);
        if (it == last.end()) { last.emplace(key, static_cast<uint32_t>(packed.size())); packed.push_back(rec); dirty.push_back(slot); }
        else packed[it->second] = rec;  // later edit to the same voxel wins
    }
    std::sort(dirty.begin(), dirty.end());
    dirty.erase(std::unique(dirty.begin(), dirty.end()), dirty.end());
    if (!packed.empty()) {
        int rc = ensure_scratch(ctx, packed.size() * sizeof(uint2));
        if (rc != DSP_OK) return rc;
        DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, packed.data(), packed.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        dsp::edit_scatter_kernel<<<static_cast<unsigned>((packed.size() + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->d_voxels, ctx->d_summary, ctx->d_bvalid, reinterpret_cast<const uint2*>(ctx->d_scratch), static_cast<uint32_t>(packed.size()));
        ctx->bmask_stale = true;
        DSP_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    if (out_n_dirty) *out_n_dirty = dirty.size();
    if (out_dirty_slots) {
        if (dirty.size() > cap_dirty) return fail(ctx, DSP_ERR_BAD_ARG, "dirty list needs %llu entries, capacity %llu",
                                                  (unsigned long long)dirty.size(), (unsigned long long)cap_dirty);
        std::memcpy(out_dirty_slots, dirty.data(), dirty.size() * sizeof(uint32_t));
    }
    return DSP_OK;
}

int dsp_apply_edits_and_remesh(dsp_ctx* ctx, uint64_t n, const dsp_edit* edits, uint32_t mode, uint64_t arena_offset, uint32_t* out_dirty_slots,
                               uint64_t cap_dirty, uint64_t* out_n_dirty, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (out_n_dirty) *out_n_dirty = 0;
    if (out_total_bytes) *out_total_bytes = 0;
    if (n == 0) return DSP_OK;
    if (!edits) return fail(ctx, DSP_ERR_BAD_ARG, "edits is NULL");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint64_t i = 0; i < n; ++i)
        if (edits[i].block > 0xFFFFu) return fail(ctx, DSP_ERR_BAD_ARG, "edit %llu: block id %u does not fit u16", (unsigned long long)i, edits[i].block);
    bool handled = false;
    uint32_t nd = 0;
    const bool halo = (mode & DSP_MESH_HALO) != 0;
    int rc = apply_edits_device(ctx, n, edits, halo, &handled, &nd);
    if (rc != DSP_OK) return rc;
    if (!handled) {  // host addressing, then the ordinary mesh call
        std::vector<uint32_t> dirty(std::min<uint64_t>(n, ctx->max_chunks));
        uint64_t nd64 = 0;
        if ((rc = dsp_apply_edits(ctx, n, edits, dirty.data(), dirty.size(), &nd64)) != DSP_OK) return rc;
        if (halo) {  // an edit on a chunk border also changes what the resident neighbour across that border may cull
            std::vector<uint32_t> more(dirty.begin(), dirty.begin() + nd64);
            for (uint64_t i = 0; i < n; ++i) {
                const dsp_edit& e = edits[i];
                const int32_t c[3] = {e.wx >> 4, e.wy >> 4, e.wz >> 4};
                uint32_t self;
                if (!find_slot(ctx, c, &self)) continue;
                const int l[3] = {e.wx & 15, e.wy & 15, e.wz & 15};
                for (int a = 0; a < 3; ++a) {
                    if (l[a] != 0 && l[a] != 15) continue;
                    int32_t q[3] = {c[0], c[1], c[2]};
                    q[a] += l[a] == 0 ? -1 : 1;
                    uint32_t s;
                    if (find_slot(ctx, q, &s)) more.push_back(s);
                }
            }
            std::sort(more.begin(), more.end());
            more.erase(std::unique(more.begin(), more.end()), more.end());
            dirty = more;
            nd64 = dirty.size();
        }
        if (out_n_dirty) *out_n_dirty = nd64;
        if (out_dirty_slots) {
            if (nd64 > cap_dirty) return fail(ctx, DSP_ERR_BAD_ARG, "dirty list needs %llu entries, capacity %llu", (unsigned long long)nd64, (unsigned long long)cap_dirty);
            std::memcpy(out_dirty_slots, dirty.data(), nd64 * sizeof(uint32_t));
        }
        return dsp_mesh_chunks(ctx, nd64, dirty.data(), mode, arena_offset, out_entries, out_total_bytes);
    }
    if (out_n_dirty) *out_n_dirty = nd;
    if (out_dirty_slots && nd > cap_dirty) return fail(ctx, DSP_ERR_BAD_ARG, "dirty list needs %u entries, capacity %llu", nd, (unsigned long long)cap_dirty);
    // the dirty list never leaves the device between the edit kernels and the mesh kernel
    if ((rc = enqueue_mesh(ctx, nd, ctx->d_slots, 0, mode, arena_offset)) != DSP_OK) return rc;
    if (out_dirty_slots && nd) DSP_CUDA(ctx, cudaMemcpyAsync(out_dirty_slots, ctx->d_slots, nd * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return finish_mesh(ctx, out_entries, out_total_bytes);
}

int dsp_get_block_world(dsp_ctx* ctx, int32_t wx, int32_t wy, int32_t wz, uint16_t* out_block) {
    if (!ctx || !out_block) return DSP_ERR_BAD_ARG;
    const int32_t c[3] = {wx >> 4, wy >> 4, wz >> 4};  // div_euclid(16)
    uint32_t slot;
    if (!find_slot(ctx, c, &slot)) return fail(ctx, DSP_ERR_NOT_RESIDENT, "chunk (%d, %d, %d) is not resident", c[0], c[1], c[2]);
    const uint32_t idx = static_cast<uint32_t>(wx & 15) + 16u * static_cast<uint32_t>(wy & 15) + 256u * static_cast<uint32_t>(wz & 15);  // rem_euclid(16), chunk.rs:44-46
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 128, ctx->d_voxels + static_cast<size_t>(slot) * dsp::kChunkVoxels + idx, sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_block = *reinterpret_cast<uint16_t*>(ctx->h_pinned + 128);
    return DSP_OK;
}

int dsp_pack_border_slabs(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, int face, void* dst_device) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (face < 0 || face > 5) return fail(ctx, DSP_ERR_BAD_ARG, "face %d out of range", face);
    if (n == 0) return DSP_OK;
    if (!slots || !dst_device) return fail(ctx, DSP_ERR_BAD_ARG, "NULL argument");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = validate_slots(ctx, n, slots);
    if (rc != DSP_OK) return rc;
    uint8_t* h;
    if ((rc = stage_host(ctx, n * sizeof(uint32_t), &h)) != DSP_OK) return rc;
    std::memcpy(h, slots, n * sizeof(uint32_t));
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_slots, h, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = stage_release(ctx)) != DSP_OK) return rc;
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(n, static_cast<uint64_t>(ctx->num_sms) * 16));
    dsp::pack_border_slabs_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_voxels, ctx->d_slots, static_cast<uint32_t>(n), face, static_cast<uint16_t*>(dst_device));
    DSP_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return DSP_OK;
}

int dsp_set_halo_slabs(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, int face, const void* src_device) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (face < 0 || face > 5) return fail(ctx, DSP_ERR_BAD_ARG, "face %d out of range", face);
    if (n == 0) return DSP_OK;
    if (!slots || !src_device) return fail(ctx, DSP_ERR_BAD_ARG, "NULL argument");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = validate_slots(ctx, n, slots);
    if (rc != DSP_OK) return rc;
    // A seam is exchanged again after every edit that touches it: the same (face, slot list) then REPLACES its earlier slabs in
    // place, so that repeated exchanges neither grow the slab store nor add launches to the neighbour-table rebuild.
    for (auto& hb : ctx->halo_batches) {
        if (hb.face != face || hb.n != n || std::memcmp(hb.slots.data(), slots, n * sizeof(uint32_t)) != 0) continue;
        DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_halo + static_cast<size_t>(hb.first_index) * 256, src_device, n * 512, cudaMemcpyDeviceToDevice, ctx->stream));
        return DSP_OK;  // same slabs indices: the neighbour table stays valid
    }
    // slabs of one call are stored contiguously (and their slots in a parallel array); a (slot, face) set again with a different
    // list uses the newer slab
    if (ctx->halo_n + n > ctx->halo_cap) {
        const uint64_t cap = std::max<uint64_t>(ctx->halo_n + n, std::max<uint64_t>(1024, ctx->halo_cap * 2));
        uint16_t* grown = nullptr;
        uint32_t* grown_slots = nullptr;
        DSP_CUDA(ctx, cudaMalloc(&grown, cap * 512));
        DSP_CUDA(ctx, cudaMalloc(&grown_slots, cap * sizeof(uint32_t)));
        if (ctx->halo_n) {
            DSP_CUDA(ctx, cudaMemcpyAsync(grown, ctx->d_halo, ctx->halo_n * 512, cudaMemcpyDeviceToDevice, ctx->stream));
            DSP_CUDA(ctx, cudaMemcpyAsync(grown_slots, ctx->d_halo_slots, ctx->halo_n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_halo);
        cudaFree(ctx->d_halo_slots);
        ctx->d_halo = grown;
        ctx->d_halo_slots = grown_slots;
        ctx->halo_cap = cap;
    }
    DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_halo + ctx->halo_n * 256, src_device, n * 512, cudaMemcpyDeviceToDevice, ctx->stream));
    dsp_ctx::HaloBatch hb{face, static_cast<uint32_t>(ctx->halo_n), static_cast<uint32_t>(n), std::vector<uint32_t>(slots, slots + n)};
    {
        uint8_t* h;
        if ((rc = stage_host(ctx, n * sizeof(uint32_t), &h)) != DSP_OK) return rc;
        std::memcpy(h, slots, n * sizeof(uint32_t));
        DSP_CUDA(ctx, cudaMemcpyAsync(ctx->d_halo_slots + ctx->halo_n, h, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = stage_release(ctx)) != DSP_OK) return rc;
    }
    ctx->halo_batches.push_back(std::move(hb));
    ctx->halo_n += n;
    ctx->nbr_dirty = true;
    return DSP_OK;
}

int dsp_clear_halo_slabs(dsp_ctx* ctx) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    ctx->halo_batches.clear();  // the slab and slot arrays are kept for the next exchange
    ctx->halo_n = 0;
    ctx->nbr_dirty = true;
    return DSP_OK;
}

int dsp_download_arena(dsp_ctx* ctx, uint64_t offset_bytes, uint64_t bytes, void* host_dst) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (offset_bytes + bytes > ctx->arena_bytes) return fail(ctx, DSP_ERR_BAD_ARG, "arena read [%llu,+%llu) out of range", (unsigned long long)offset_bytes, (unsigned long long)bytes);
    if (bytes == 0) return DSP_OK;
    if (!host_dst) return fail(ctx, DSP_ERR_BAD_ARG, "host_dst is NULL");
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    DSP_CUDA(ctx, cudaMemcpyAsync(host_dst, ctx->d_arena + offset_bytes, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DSP_OK;
}

int dsp_download_chunk(dsp_ctx* ctx, uint32_t slot, uint16_t* host_dst, int32_t* out_pos) {
    if (!ctx || !host_dst) return DSP_ERR_BAD_ARG;
    if (slot >= ctx->next_unused || !ctx->resident[slot]) return fail(ctx, DSP_ERR_NOT_RESIDENT, "slot %u is not resident", slot);
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    DSP_CUDA(ctx, cudaMemcpyAsync(host_dst, ctx->d_voxels + static_cast<size_t>(slot) * dsp::kChunkVoxels, dsp::kChunkBytes, cudaMemcpyDeviceToHost, ctx->stream));
    int4 p = make_int4(0, 0, 0, 0);
    if (out_pos) DSP_CUDA(ctx, cudaMemcpyAsync(&p, ctx->d_pos + slot, sizeof p, cudaMemcpyDeviceToHost, ctx->stream));
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_pos) { out_pos[0] = p.x; out_pos[1] = p.y; out_pos[2] = p.z; }
    return DSP_OK;
}

void* dsp_arena_device_ptr(dsp_ctx* ctx) { return ctx ? ctx->d_arena : nullptr; }
uint64_t dsp_arena_bytes(const dsp_ctx* ctx) { return ctx ? ctx->arena_bytes : 0; }
void* dsp_voxel_device_ptr(dsp_ctx* ctx) { return ctx ? ctx->d_voxels : nullptr; }
void* dsp_entries_device_ptr(dsp_ctx* ctx) { return ctx ? ctx->d_entries : nullptr; }
void* dsp_stream(dsp_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
uint64_t dsp_kernel_launches(const dsp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dsp_host_upload_stats(dsp_ctx* ctx, uint64_t out[8]) {
    if (!ctx || !out) return DSP_ERR_BAD_ARG;
    std::memcpy(out, ctx->upload_stats, sizeof(ctx->upload_stats));
    return DSP_OK;
}

int dsp_debug_pack_chunk(const uint16_t* blocks, uint8_t* packed, uint32_t* desc, int portable) {
    if (!blocks || !packed || !desc) return DSP_ERR_BAD_ARG;
    *desc = portable ? dsp_host::pack_chunk_portable(blocks, packed) : dsp_host::pack_chunk(blocks, packed);
    return DSP_OK;
}

// debug: reads and clears the 16 phase counters (all zero unless the library was built with -DDSP_PHASE_TIMERS)
int dsp_debug_phase_cycles(dsp_ctx* ctx, uint64_t* out16) {
    if (!ctx || !out16) return DSP_ERR_BAD_ARG;
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DSP_CUDA(ctx, cudaMemcpy(out16, ctx->d_phase, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    DSP_CUDA(ctx, cudaMemset(ctx->d_phase, 0, 32 * sizeof(uint64_t)));
    return DSP_OK;
}

#ifdef DSP_PHASE_TIMERS
// debug builds only (not in the header, not in the shipped library): the per-CTA timelines of the last quad-kernel launch
int dsp_debug_timeline(dsp_ctx* ctx, uint64_t* out, uint64_t words) {
    if (!ctx || !out || words > kPhaseWords - 32) return DSP_ERR_BAD_ARG;
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DSP_CUDA(ctx, cudaMemcpy(out, ctx->d_phase + 32, words * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    DSP_CUDA(ctx, cudaMemset(ctx->d_phase + 32, 0, words * sizeof(uint64_t)));
    return DSP_OK;
}
#endif

int dsp_set_kernel_timing(dsp_ctx* ctx, int on) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    ctx->kernel_timing = on != 0;
    return DSP_OK;
}

int dsp_mesh_kernel_stats(dsp_ctx* ctx, uint64_t* out_count, double* out_total_ms, float* out_last_ms) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    for (int i = 0; i < dsp_ctx::kEvRing; ++i) {  // oldest first so that `last` is the newest launch
        int rc = drain_kernel_event(ctx, (ctx->ev_next + i) % dsp_ctx::kEvRing);
        if (rc != DSP_OK) return rc;
    }
    if (out_count) *out_count = ctx->kernel_count;
    if (out_total_ms) *out_total_ms = ctx->kernel_ms_total;
    if (out_last_ms) *out_last_ms = ctx->kernel_ms_last;
    return DSP_OK;
}

}  // extern "C"

#include "dsp_scene.inl"
#include "dsp_seam.inl"
#include "dsp_export.inl"
