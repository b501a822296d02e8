// Fused face-cull + prefix-sum compaction + vertex emission (sm_100a).
//
// Replaces build_chunk_mesh, /root/reference/src/content/world/chunks/chunk_mesh.rs:13-129:
//   :32-43   voxel loop, coordinate derivation      -> one thread per x-row, 16-bit occupancy masks
//   :46-51   six neighbour-air tests                -> shifts / and-not on row masks (border = air)
//   :53-56   atlas lookup with (0,0)-(1,1) fallback -> pre-mapped per-block table (uv_map_kernel)
//   :61-62   pos_offset                             -> (f32)local + (f32)chunk * 16, then +-0.5
//   :65-106  emission order TOP..LEFT per voxel     -> order-preserving scan, so the stream is
//                                                     byte-identical, not merely set-equal
//   :109-128 one device buffer per chunk            -> one 128-byte aligned slot of a shared arena
//
// One persistent CTA (256 threads) takes chunks from a ticket counter (first chunk = block index).  Per chunk:
//   1. the 8 KiB voxel tile arrives in shared memory by a 1-D bulk (TMA) copy that was issued
//      while the previous chunk was being processed (2-deep mbarrier pipeline);
//   2. thread r builds the occupancy mask of x-row r, derives the six face masks and its face
//      count; a warp-shuffle + 8-way scan gives every row its first face ordinal and the CTA the
//      chunk's face total F (all-air chunks, and chunks with no visible face, leave here);
//   3. the chunk's arena slot of align128(120*F) bytes: one 64-bit atomicAdd on a cursor -- slots gap-free, in completion
//      order; with DSP_MESH_ORDERED a count + scan pass (count_chunks_scan_kernel) has fixed every chunk's offset in call
//      order beforehand and the kernel takes it from the chunk's entry (p.preset);
//      meanwhile all rows write 16-bit face records (voxel index, direction) in stream order;
//   4. faces are expanded one thread per face into a 30,720-byte staging tile (256 faces) that leaves with ONE bulk
//      shared->global copy (cp.async.bulk), so only full-line writes reach L2/HBM (MeshCfg variants: per-warp tiles);
//   5. the last CTA to finish publishes {status, total} and zeroes the control block for the next launch.
// Algorithmic HBM bytes per chunk: 8,192 (voxels) + 120*F (vertices) + 16 (position) + 16 (entry).
#pragma once
#include "dsp_common.cuh"

namespace dsp {

struct MeshEntry {  // == dsp_mesh_entry
    uint64_t offset_bytes;
    uint32_t vertex_count;
    uint32_t index_count;
};

struct MeshParams {
    const uint16_t* voxels;     // [max_chunks][4096]
    const int4* chunk_pos;      // per slot
    const uint32_t* slot_list;  // nullable
    const uint32_t* nbr_slots;  // [max_chunks][6] (halo mode) or nullptr: neighbour slot, kNbrExternal | slab index, or none
    const uint16_t* bmask;      // [max_chunks][96] border occupancy masks (halo mode): 6 one-voxel border layers x 16 masks per chunk
    const uint16_t* halo_slabs; // [n_slabs][256] border layers received from other devices (halo mode)
    const float4* uvmap;        // [n_blocks + 1], last row = fallback
    uint32_t n_blocks;
    uint32_t first_slot;
    uint32_t n;
    uint8_t* arena;             // already offset by arena_offset
    uint64_t arena_offset;
    uint64_t arena_room;        // bytes available from arena_offset
    MeshEntry* entries;         // [n]
    uint32_t* face_counts;      // [n] per-chunk face counts of count_chunks_scan_kernel (DSP_MESH_ORDERED)
    uint32_t* ticket;           // chunk ticket counter of this launch (a cache line of its own); zeroed before launch
    uint32_t* ctl;              // the context's control block (work-list counts live there)
    uint32_t* status;           // status bits (kStatus*), shared by the launches of one call
    uint64_t* total_bytes;      // out
    unsigned long long* phase_cycles;  // [16] debug: per-phase SM cycles of thread 0 summed over CTAs (DSP_PHASE_TIMERS builds)
    uint32_t prefetch_ahead;    // L2-prefetch the tile of ticket + prefetch_ahead when a tile load is issued (0 = off)
    // Self-resetting control block (single-launch calls): every CTA counts itself out on `done`; the last one copies
    // {status, total} to `result` and zeroes ticket / status / total / done, so the NEXT launch needs no memset in front
    // of it -- a step is exactly one kernel launch.  nullptr: the caller resets the block itself (multi-launch calls).
    uint32_t* done;
    uint64_t* result;           // [0] = status, [1] = total bytes
    // Work list produced on the device by classify_chunks_kernel (halo mode: all-air and buried chunks never reach the mesh
    // kernel): when n_dev != nullptr the number of chunks is read from it and entry_map[i] is the entry index of chunk i.
    const uint32_t* n_dev;
    const uint32_t* entry_map;
    const uint16_t* uid;        // per slot: the chunk's one block id where classify_greedy_kernel listed it as uniform
    // Chunk-local culling without a pre-pass (mesh_chunks_bump_kernel): per-slot summary bytes, or nullptr.  A chunk whose summary
    // says "all air" is finished by the thread that takes its ticket -- empty entry, no tile load, no CTA iteration; a chunk of one
    // known solid block id is meshed from that id without its tile being read (uid above).
    const uint8_t* summary;
    bool summary_air_skip;
    // Host-resident chunks (dsp_mesh_host_chunks*, pinned / registered memory): tile of ticket i = host_tiles + i * 4096, pulled over
    // PCIe by the same 1-D bulk (TMA) copy that otherwise reads the voxel store -- no copy engine, no staging buffer, no separate
    // upload kernel; once it is in shared memory the tile is also stored to the chunk's slot of the voxel store (voxels_out), so
    // the chunk is resident afterwards like any other.  nullptr: tiles come from the voxel store.
    const uint16_t* host_tiles;
    uint16_t* voxels_out;
    // Persistent launch over a batch that is still being uploaded (packed upload of host chunks): the batch arrives in pieces of
    // piece_chunks chunks; released[j] != 0 once piece j is in the voxel store -- tiles, summaries, positions and slot-list entries
    // written by an unpack kernel that ran beside this launch and set the word behind its last write.  Thread 0 waits (bounded)
    // before it uses a ticket of a piece it has not seen released; the per-chunk metadata is then read past the non-coherent L1
    // path (it was NOT constant for the lifetime of this kernel).  nullptr: everything was there before the launch.
    const uint32_t* released;
    uint32_t piece_chunks;
    // DSP_MESH_ORDERED: count_chunks_scan_kernel has already written every chunk's entry (slots packed in call order); the mesh
    // kernel takes each chunk's offset from its entry instead of claiming one, and writes neither entries nor the total.
    bool preset;
    // Work list layout: chunks that need nothing from another GPU fill it from the FRONT (count n_dev[0]), chunks with a
    // neighbour across a region seam from the BACK (count n_dev[3], list index n_cap - 1 - k).  Tickets run over the front
    // part first, so by the time a CTA reaches a seam chunk the neighbour GPU's border layers have usually arrived --
    // the exchange hides behind the interior of the region, inside ONE launch (seam_wait below).
    uint32_t n_front, n_cap;    // filled in by the kernel from n_dev
    // Region seams (dsp_seam_*): ext_base[0] = halo_slabs; ext_base[1 + s] = the inbox buffer of seam s that holds the current
    // epoch's border layers, written by the neighbour GPU over NVLink; seam_flag[s] = the inbox word the neighbour sets to the
    // epoch once its layers are complete.
    const uint16_t* ext_base[8];
    const uint32_t* seam_flag[7];
    uint32_t n_seams, seam_epoch;
};

// Chunk summaries (one byte per slot, written by the generators, reset to "unknown" by edits / uploads / dsp_invalidate_chunks;
// terrain_kernels.cuh has the full story) and the uniform-id side array.
constexpr uint16_t kUidMixed = 0xFFFFu;  // (uid side array) free of air, but more than one block id -- or not known
constexpr uint8_t kSummaryUnknown = 0, kSummaryAir = 1, kSummarySolid = 2, kSummaryLayer0 = 4, kSummaryAllSolid = 2 | (63 << 2);

constexpr uint32_t kNbrNone = 0xFFFFFFFFu;
constexpr uint32_t kNbrExternal = 0x80000000u;  // neighbour lives on another device: bits 28-30 = source (0: dsp_set_halo_slabs,
constexpr uint32_t kNbrSrcShift = 28u;          // 1 + s: seam s), bits 0-27 = slab index within that source
constexpr uint32_t kNbrIndexMask = 0x0FFFFFFFu;
constexpr uint32_t kStatusOverflow = 1u;
constexpr uint32_t kStatusSpin = 2u;  // (unused since the look-back organisation went; kept so that the bit is not reassigned)
constexpr uint32_t kStatusSeam = 4u;  // a neighbour GPU's border layers did not arrive (dsp_seam_exchange_async missing on the peer?)

// position in the work list of ticket `chunk` (identity without a device-built list)
__device__ __forceinline__ uint32_t work_index(const MeshParams& p, uint32_t chunk) {
    return (p.n_dev == nullptr || chunk < p.n_front) ? chunk : p.n_cap - 1u - (chunk - p.n_front);
}
// called at kernel start on the kernel's own copy of the parameters
__device__ __forceinline__ void adopt_work_list(MeshParams& p) {
    if (p.n_dev == nullptr) { p.n_front = 0u; return; }
    p.n_cap = p.n;
    p.n_front = p.n_dev[0];
    p.n = p.n_front + p.n_dev[3];
}
// A chunk of the seam part of the list may only be culled once every neighbour GPU has delivered this epoch's layers.
// Every thread spins for itself (acquire loads on a word in LOCAL memory that the peer sets with a system-scope release
// after its data); the spin is bounded so that a missing peer becomes an error code, not a hung GPU.
template <bool HALO>
__device__ __forceinline__ void seam_wait(const MeshParams& p, uint32_t chunk) {
    if (!HALO || p.n_seams == 0u || chunk < p.n_front) return;
#pragma unroll
    for (uint32_t s = 0; s < 7u; ++s) {  // (fully unrolled: a dynamic index would push the parameter block into local memory)
        if (s >= p.n_seams) break;
        uint32_t spins = 0;
        while (static_cast<int32_t>(ld_acquire_sys_u32(p.seam_flag[s]) - p.seam_epoch) < 0) {
            __nanosleep(200);
            if (++spins > (1u << 22)) { atomicOr(p.status, kStatusSeam); break; }
        }
    }
}

__device__ __forceinline__ uint32_t work_index(const MeshParams& p, uint32_t chunk);
__device__ __forceinline__ MeshEntry* entry_of(const MeshParams& p, uint32_t chunk) {
    return p.entries + (p.entry_map ? __ldg(p.entry_map + work_index(p, chunk)) : chunk);
}

// Last instructions of a persistent mesh kernel (thread 0 of every CTA, after its stores were issued).
__device__ __forceinline__ void control_block_epilogue(const MeshParams& p) {
    if (p.done == nullptr) return;
    __threadfence();  // this CTA's atomics on total / status are performed before it counts itself out
    if (atomicAdd(p.done, 1u) != gridDim.x - 1u) return;
    __threadfence();
    p.result[0] = atomicOr(p.status, 0u);
    p.result[1] = atomicAdd(reinterpret_cast<unsigned long long*>(p.total_bytes), 0ull);
    p.ticket[0] = 0u;
    if (p.preset) p.ticket[32] = 0u;  // the count pass's "CTAs done" counter (DSP_MESH_ORDERED)
    p.ctl[3] = 0u;  // the work-list counts of the classify kernels: front, back (seam chunks) ...
    p.ctl[6] = 0u;
    p.ctl[7] = 0u;  // ... and the uniform chunks of classify_greedy_kernel
    *p.status = 0u;
    *p.total_bytes = 0ull;
    *p.done = 0u;
}

constexpr int kNbrMasks = 96;           // 6 neighbours x 16 border masks (see fetch_nbr_mask)
constexpr int kBorderMaskStride = 96;   // u16 per chunk in MeshParams::bmask
constexpr int kTileFaces = 256;
constexpr int kTileBytes = kTileFaces * kFaceBytes;  // 30,720
constexpr int kWarpSlotWords = 32 * kFaceWords;      // 3,840 bytes: one warp's 32 faces

// Corner codes.  Hand-typed from /root/reference/src/engine/rendering/cube.rs:156-312.  Each face is
// 4 corners (c0,c1,c2,c3) emitted as c0,c1,c2,c2,c3,c0; a corner is 5 bits:
// bit0/1/2 = sign of the x/y/z offset (1 -> +0.5), bit3 = u, bit4 = v of the template tex_coords.
constexpr uint32_t corner(int sx, int sy, int sz, int u, int v) {
    return static_cast<uint32_t>((sx > 0 ? 1 : 0) | (sy > 0 ? 2 : 0) | (sz > 0 ? 4 : 0) | (u ? 8 : 0) | (v ? 16 : 0));
}
constexpr uint32_t face_code(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) { return c0 | (c1 << 5) | (c2 << 10) | (c3 << 15); }
constexpr uint32_t kCodeTop = face_code(corner(-1, 1, -1, 1, 1), corner(1, 1, -1, 0, 1), corner(1, 1, 1, 0, 0), corner(-1, 1, 1, 1, 0));        // cube.rs:287-312
constexpr uint32_t kCodeBottom = face_code(corner(-1, -1, 1, 1, 1), corner(1, -1, 1, 0, 1), corner(1, -1, -1, 0, 0), corner(-1, -1, -1, 1, 0));  // cube.rs:261-286
constexpr uint32_t kCodeRear = face_code(corner(-1, -1, -1, 0, 1), corner(1, -1, -1, 1, 1), corner(1, 1, -1, 1, 0), corner(-1, 1, -1, 0, 0));    // cube.rs:183-208
constexpr uint32_t kCodeFront = face_code(corner(-1, -1, 1, 1, 1), corner(-1, 1, 1, 1, 0), corner(1, 1, 1, 0, 0), corner(1, -1, 1, 0, 1));       // cube.rs:156-181
constexpr uint32_t kCodeRight = face_code(corner(-1, -1, -1, 0, 1), corner(-1, 1, -1, 0, 0), corner(-1, 1, 1, 1, 0), corner(-1, -1, 1, 1, 1));   // cube.rs:209-234 (-X)
constexpr uint32_t kCodeLeft = face_code(corner(1, -1, -1, 1, 1), corner(1, -1, 1, 0, 1), corner(1, 1, 1, 0, 0), corner(1, 1, -1, 1, 0));        // cube.rs:235-260 (+X)
// Emission order TOP, BOTTOM, REAR, FRONT, RIGHT, LEFT (chunk_mesh.rs:65-106): 3 codes per 64-bit word.
constexpr uint64_t kCodes0 = static_cast<uint64_t>(kCodeTop) | (static_cast<uint64_t>(kCodeBottom) << 20) | (static_cast<uint64_t>(kCodeRear) << 40);
constexpr uint64_t kCodes1 = static_cast<uint64_t>(kCodeFront) | (static_cast<uint64_t>(kCodeRight) << 20) | (static_cast<uint64_t>(kCodeLeft) << 40);

// occupancy bits of 8 consecutive u16 voxels: per 16-bit lane min(v, 1) is one VIMNMX.U16x2; the four words are then
// interleaved with three multiply-adds (lanes hold bits 0,2,4,6 / 16,18,20,22) and folded -- 9 instructions for 8 voxels
__device__ __forceinline__ uint32_t nonzero_mask8(const uint4 q) {
    const uint32_t r0 = __vminu2(q.x, 0x00010001u), r1 = __vminu2(q.y, 0x00010001u), r2 = __vminu2(q.z, 0x00010001u), r3 = __vminu2(q.w, 0x00010001u);
    const uint32_t c = r0 + 4u * r1 + 16u * r2 + 64u * r3;
    return (c | (c >> 15)) & 0xFFu;
}

// Expands one face record into its 6 vertices (30 words) at dst (8-byte aligned shared memory).
__device__ __forceinline__ void emit_face(float* dst, uint32_t rec, const uint16_t* vox, float cbx, float cby, float cbz,
                                          const float4* __restrict__ uvmap, uint32_t n_blocks) {
    const uint32_t d = rec & 7u, idx = rec >> 3;
    const uint32_t id = vox[idx];
    const float4 uv = __ldg(&uvmap[id < n_blocks ? id : n_blocks]);  // (u(0), v(0), u(1), v(1))
    // chunk_mesh.rs:61-62: pos_offset = (local as f32) + (chunk as f32 * 16.0); template +-0.5 added after.
    const float ox = __fadd_rn(__uint2float_rn(idx & 15u), cbx);
    const float oy = __fadd_rn(__uint2float_rn((idx >> 4) & 15u), cby);
    const float oz = __fadd_rn(__uint2float_rn(idx >> 8), cbz);
    const float lx = __fadd_rn(-0.5f, ox), hx = __fadd_rn(0.5f, ox);
    const float ly = __fadd_rn(-0.5f, oy), hy = __fadd_rn(0.5f, oy);
    const float lz = __fadd_rn(-0.5f, oz), hz = __fadd_rn(0.5f, oz);
    const uint32_t code = static_cast<uint32_t>((d < 3 ? kCodes0 >> (20 * d) : kCodes1 >> (20 * (d - 3))) & 0xFFFFFu);
    float cx[4], cy[4], cz[4], cu[4], cv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = code >> (5 * k);
        cx[k] = (c & 1u) ? hx : lx;
        cy[k] = (c & 2u) ? hy : ly;
        cz[k] = (c & 4u) ? hz : lz;
        cu[k] = (c & 8u) ? uv.z : uv.x;
        cv[k] = (c & 16u) ? uv.w : uv.y;
    }
    float2* o = reinterpret_cast<float2*>(dst);
    o[0] = make_float2(cx[0], cy[0]);   o[1] = make_float2(cz[0], cu[0]);   o[2] = make_float2(cv[0], cx[1]);
    o[3] = make_float2(cy[1], cz[1]);   o[4] = make_float2(cu[1], cv[1]);   o[5] = make_float2(cx[2], cy[2]);
    o[6] = make_float2(cz[2], cu[2]);   o[7] = make_float2(cv[2], cx[2]);   o[8] = make_float2(cy[2], cz[2]);
    o[9] = make_float2(cu[2], cv[2]);   o[10] = make_float2(cx[3], cy[3]);  o[11] = make_float2(cz[3], cu[3]);
    o[12] = make_float2(cv[3], cx[0]);  o[13] = make_float2(cy[0], cz[0]);  o[14] = make_float2(cu[0], cv[0]);
}

// Kernel variant: how many CTAs share an SM (sets the shared-memory budget), how many 30,720-byte staging
// tiles each CTA owns, and whether emission is per CTA tile (one bulk store per 256 faces, two CTA barriers per
// tile) or per warp (one bulk store per 32 faces, no CTA barrier in the emission loop).
template <int CTAS, int NSTAGE_, bool WARP_EMIT_>
struct MeshCfg {
    static constexpr int kCtasPerSm = CTAS;
    static constexpr int NSTAGE = NSTAGE_;
    static constexpr bool WARP_EMIT = WARP_EMIT_;
    // Face records are produced in windows of kRecCap faces; 4 CTAs/SM only fit with an 8 KiB window.  A chunk with
    // more faces than the window (dense noise, > 1/3 of the 12,288 maximum) takes several record/emit rounds.
    static constexpr int kRecCap = CTAS >= 4 ? 4096 : kMaxFaces;
    static constexpr int kVox = 0;                                   // 2 x 8192
    static constexpr int kStage = 2 * kChunkBytes;                   // NSTAGE x 30720
    static constexpr int kRec = kStage + NSTAGE * kTileBytes;        // kRecCap x u16
    static constexpr int kRowMask = kRec + kRecCap * 2;              // 256 x u16 row masks + 96 x u16 neighbour border masks
    static constexpr int kWarpTot = kRowMask + (kRows + kNbrMasks) * 2;  // 8 x u32 (+8 pad)
    static constexpr int kBar = kWarpTot + 64;                       // 2 x u64
    static constexpr int kBase = kBar + 16;                          // u64
    static constexpr int kNext = kBase + 8;                          // 2 x u32 next tickets + 2 x u32 their slots + 2 x u32 "uniform chunk of block id" words
    static constexpr int kBytes = kNext + 32;
};

// Border occupancy masks.  Every chunk has a 192-byte side record: for each of its six one-voxel border layers (face order top,
// bottom, rear, front, right, left) 16 masks of 16 bits -- +-Y layers: mask[z] bit x, +-Z layers: mask[y] bit x, +-X layers:
// mask[z] bit y (the layouts of dsp_pack_border_slabs, one bit per voxel instead of its u16 id).  Cross-chunk culling needs
// nothing else of a neighbour, so a chunk's six neighbours cost six 32-byte reads instead of strided reads all over their
// voxel tiles (round 1: the +-X neighbours alone touched 512 sectors for 64 useful bytes).

// Thread t < 96 of a CTA fetches mask k = t % 16 of the neighbour across face f = t / 16 of chunk `slot`: from the neighbour's
// side record when it is resident, from the received border layer (16 voxels -> 16 bits) when it lives on another GPU,
// 0 (air: the border face is emitted, chunk_mesh.rs:46-51) when there is none.
__device__ __forceinline__ uint32_t fetch_nbr_mask(const MeshParams& p, uint32_t slot, int t) {
    if (t >= kNbrMasks) return 0u;
    const int f = t >> 4, k = t & 15;
    const uint32_t code = __ldg(p.nbr_slots + static_cast<size_t>(slot) * 6 + f);
    if (code == kNbrNone) return 0u;
    if (code & kNbrExternal) {
        // (read through L2 only: a seam's inbox is written by the neighbour GPU while this kernel runs)
        const uint32_t src = (code >> kNbrSrcShift) & 7u;
        const uint16_t* base = p.ext_base[0];
#pragma unroll
        for (uint32_t j = 1; j < 8u; ++j) base = src == j ? p.ext_base[j] : base;  // selects, not a dynamic index (see seam_wait)
        const uint4* q = reinterpret_cast<const uint4*>(base + static_cast<size_t>(code & kNbrIndexMask) * 256 + k * 16);
        return nonzero_mask8(ld_cg_u4(q)) | (nonzero_mask8(ld_cg_u4(q + 1)) << 8);
    }
    return __ldg(p.bmask + static_cast<size_t>(code) * kBorderMaskStride + (f ^ 1) * 16 + k);  // the layer that faces us
}

// The same fetch in two stages, for prefetching the NEXT chunk's masks while the current one is processed (each stage is one
// global load whose latency hides behind a phase of the current chunk): the neighbour code first, the mask later.  Masks of
// neighbours on other GPUs are not prefetched -- their inbox may only be read after seam_wait -- `*ready` stays false then and the
// consumer falls back to fetch_nbr_mask.
__device__ __forceinline__ uint32_t prefetch_nbr_code(const MeshParams& p, uint32_t slot, int t) {
    return t < kNbrMasks ? __ldg(p.nbr_slots + static_cast<size_t>(slot) * 6 + (t >> 4)) : kNbrNone;
}
__device__ __forceinline__ uint32_t prefetch_nbr_mask(const MeshParams& p, uint32_t code, int t, bool* ready) {
    *ready = true;
    if (t >= kNbrMasks || code == kNbrNone) return 0u;
    if (code & kNbrExternal) { *ready = false; return 0u; }
    return __ldg(p.bmask + static_cast<size_t>(code) * kBorderMaskStride + ((t >> 4) ^ 1) * 16 + (t & 15));
}

// The six 16-bit face masks of x-row r = y + 16 z (bit x set: that face of voxel (x, y, z) is emitted), from the row
// occupancy masks of the chunk in shared memory.  chunk_mesh.rs:46-51.
struct FaceMasks {
    uint32_t top, bot, rear, front, right, left;
};
template <bool HALO>
__device__ __forceinline__ FaceMasks row_face_masks(const uint16_t* s_rowmask, int r, int y, int z, uint32_t m) {
    // chunk_mesh.rs:46-51: outside the chunk counts as air (mask 0) in parity mode
    uint32_t up = y < 15 ? s_rowmask[r + 1] : 0u;
    uint32_t dn = y > 0 ? s_rowmask[r - 1] : 0u;
    uint32_t re = z > 0 ? s_rowmask[r - 16] : 0u;
    uint32_t fr = z < 15 ? s_rowmask[r + 16] : 0u;
    uint32_t xm = 0u, xp = 0u;  // occupancy of the voxel beyond x = 0 / x = 15
    if (HALO) {
        // extension (DSP_MESH_HALO): the border rows consult the neighbour chunk across each face -- its border masks are in
        // shared memory behind the row masks (fetch_nbr_mask, stored by count_phase / the quad kernel before the barrier)
        const uint16_t* nb = s_rowmask + kRows;
        if (y == 15) up = nb[0 * 16 + z];
        if (y == 0) dn = nb[1 * 16 + z];
        if (z == 0) re = nb[2 * 16 + y];
        if (z == 15) fr = nb[3 * 16 + y];
        xm = (nb[4 * 16 + z] >> y) & 1u;
        xp = (nb[5 * 16 + z] >> y) & 1u;
    }
    FaceMasks f;
    f.top = m & ~up; f.bot = m & ~dn; f.rear = m & ~re; f.front = m & ~fr;
    f.right = m & ~((m << 1) | xm);
    f.left = m & ~((m >> 1) | (xp << 15));
    return f;
}

// Per-thread result of the count phase for one chunk (thread r owns x-row r = y + 16 z).
struct RowFaces {
    uint32_t top, bot, rear, front, right, left;  // 16-bit face masks of this row
    uint32_t cnt;                                  // faces in this row
    uint32_t first;                                // ordinal of this row's first face within the chunk
    uint32_t total;                                // faces in the chunk (same in every thread)
};

__device__ __forceinline__ uint64_t slot_bytes_of(uint32_t faces) {
    return (static_cast<uint64_t>(faces) * kFaceBytes + (kSlotAlign - 1)) & ~static_cast<uint64_t>(kSlotAlign - 1);
}

// ===== DSP_MESH_ORDERED, pass 1: face count of every chunk + exclusive scan of the slot sizes, in call order ==================
// CTA b counts chunks [16 b, 16 b + 16), one per warp: the 8 KiB tile is read with sixteen fully coalesced 128-bit loads per
// lane (none at all where the chunk's summary says all air or no air: 0 / 1,536 faces), the 256 row masks go to a 704-byte
// shared-memory strip of the warp and every lane counts the faces of eight rows (chunk_mesh.rs:46-51, the same row_face_masks as
// the mesh kernel).  Counts go to a u32 array; CTAs count themselves out, and the LAST one to finish scans the array (512
// threads, 2,048 chunks per sweep with a running carry) and writes every chunk's entry -- no CTA ever waits for another.  The
// mesh kernel that follows (mesh_chunks_bump_kernel with p.preset) emits every chunk at the offset found here: slots packed back
// to back in call order.  History: rounds 1-2 did this in ONE pass with a decoupled look-back inside the mesh kernel (97 us
// against 78 us for the atomic claim: a chunk's predecessors were busy emitting when it wanted their sizes); a look-back inside
// this count pass measured 14-16 us for the pass (every CTA finishes counting in the same microsecond, so the chain of waits
// IS the kernel); the last-CTA scan has no chain.
constexpr int kCountChunksPerCta = 16;
constexpr int kCountThreads = kCountChunksPerCta * 32;
template <bool HALO>
__global__ void __launch_bounds__(kCountThreads, 2) count_chunks_scan_kernel(const MeshParams p_in) {
    pdl_launch_dependents();
    pdl_wait();
    MeshParams p = p_in;
    p.n_front = 0u;  // no device-built work list in this mode
    __shared__ __align__(16) uint16_t s_masks[kCountChunksPerCta][kRows + kNbrMasks];
    __shared__ unsigned long long s_warp_tot[kCountChunksPerCta];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* const rm = s_masks[warp];
    {
        const uint64_t chunk64 = static_cast<uint64_t>(blockIdx.x) * kCountChunksPerCta + warp;
        if (chunk64 < p.n) {
            const uint32_t chunk = static_cast<uint32_t>(chunk64);
            const uint32_t slot = p.slot_list ? __ldg(p.slot_list + chunk) : p.first_slot + chunk;
            const uint32_t summary = (!HALO && p.summary != nullptr) ? __ldg(p.summary + slot) : 0u;
            uint32_t faces;
            if (summary & kSummaryAir) {
                faces = 0u;  // palette.len() == 1 (chunk_mesh.rs:18-20)
            } else if (summary & kSummarySolid) {
                faces = 6u * 256u;  // no air anywhere, chunk-local cull: exactly the six outer layers
            } else {
                seam_wait<HALO>(p, chunk);
                const uint4* tile = reinterpret_cast<const uint4*>(p.voxels + static_cast<size_t>(slot) * kChunkVoxels);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint4 q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) q[i] = __ldg(tile + (half * 8 + i) * 32 + lane);  // 16-byte unit u = 32 i + lane: half (u & 1) of row u >> 1
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t b = nonzero_mask8(q[i]);
                        const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, b, 1);
                        if (!(lane & 1)) rm[(half * 8 + i) * 16 + (lane >> 1)] = static_cast<uint16_t>(b | (o << 8));
                    }
                }
                if (HALO) {
#pragma unroll
                    for (int j = 0; j < kNbrMasks / 32; ++j) rm[kRows + lane + 32 * j] = static_cast<uint16_t>(fetch_nbr_mask(p, slot, lane + 32 * j));
                }
                __syncwarp();
                uint32_t cnt = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = lane + 32 * j;
                    const FaceMasks fm = row_face_masks<HALO>(rm, r, r & 15, r >> 4, rm[r]);
                    cnt += __popc(fm.top) + __popc(fm.bot) + __popc(fm.rear) + __popc(fm.front) + __popc(fm.right) + __popc(fm.left);
                }
                faces = __reduce_add_sync(0xFFFFFFFFu, cnt);
            }
            if (lane == 0) p.face_counts[chunk] = faces;
        }
    }
    // count this CTA out; the last one does the scan
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    unsigned long long carry = 0ull;
    bool overflow = false;
    for (uint64_t base = 0; base < p.n; base += kCountThreads * 4) {
        const uint64_t c0 = base + threadIdx.x * 4;
        uint4 f = make_uint4(0u, 0u, 0u, 0u);
        if (c0 + 3 < p.n) f = __ldcg(reinterpret_cast<const uint4*>(p.face_counts + c0));
        else if (c0 < p.n) {
            f.x = __ldcg(p.face_counts + c0);
            if (c0 + 1 < p.n) f.y = __ldcg(p.face_counts + c0 + 1);
            if (c0 + 2 < p.n) f.z = __ldcg(p.face_counts + c0 + 2);
        }
        const unsigned long long b0 = slot_bytes_of(f.x), b1 = slot_bytes_of(f.y), b2 = slot_bytes_of(f.z), b3 = slot_bytes_of(f.w);
        unsigned long long incl = b0 + b1 + b2 + b3;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        unsigned long long wpre = 0ull, sweep = 0ull;
#pragma unroll
        for (int w = 0; w < kCountChunksPerCta; ++w) {
            const unsigned long long v = s_warp_tot[w];
            wpre += w < warp ? v : 0ull;
            sweep += v;
        }
        unsigned long long off = carry + wpre + incl - (b0 + b1 + b2 + b3);
        const uint32_t fv[4] = {f.x, f.y, f.z, f.w};
        const unsigned long long bv[4] = {b0, b1, b2, b3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (c0 + k < p.n) {
                overflow |= off + bv[k] > p.arena_room;
                p.entries[c0 + k] = MeshEntry{p.arena_offset + off, fv[k] * 6u, 0u};
            }
            off += bv[k];
        }
        carry += sweep;
        __syncthreads();  // s_warp_tot is rewritten by the next sweep
    }
    if (overflow) atomicOr(p.status, kStatusOverflow);
    if (threadIdx.x == 0) *p.total_bytes = carry;
}

// Per-CTA working set of the mesh kernel.  LIVE: the launch runs beside the upload of its own batch (MeshParams::released) -- a separate
// instantiation, so that the kernel every other call uses carries none of the waiting or the coherent metadata loads.
template <class Cfg, bool HALO, bool LIVE = false>
struct MeshCta {
    static constexpr int NSTAGE = Cfg::NSTAGE;
    static constexpr uint32_t kRecCap = Cfg::kRecCap;
    const MeshParams& p;
    uint16_t* s_vox;
    float* s_stage;
    uint16_t* s_rec;
    uint16_t* s_rowmask;
    uint32_t* s_warp;
    uint64_t* s_bar;
    uint64_t* s_base;
    uint32_t* s_next;
    int tid, lane, warp, r, y, z;
    uint32_t stage_ctr = 0;
    // halo mode: the neighbour masks of the chunk that comes next, fetched in two stages during the current chunk
    uint32_t pf_code = kNbrNone, pf_val = 0u;
    bool pf_ready = false;
    uint32_t released_seen = 0;  // thread 0: bit j = piece j of the batch has been seen released (see MeshParams::released)

    __device__ __forceinline__ MeshCta(const MeshParams& params, uint8_t* smem) : p(params) {
        s_vox = reinterpret_cast<uint16_t*>(smem + Cfg::kVox);
        s_stage = reinterpret_cast<float*>(smem + Cfg::kStage);
        s_rec = reinterpret_cast<uint16_t*>(smem + Cfg::kRec);
        s_rowmask = reinterpret_cast<uint16_t*>(smem + Cfg::kRowMask);
        s_warp = reinterpret_cast<uint32_t*>(smem + Cfg::kWarpTot);
        s_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kBar);
        s_base = reinterpret_cast<uint64_t*>(smem + Cfg::kBase);
        s_next = reinterpret_cast<uint32_t*>(smem + Cfg::kNext);
        tid = threadIdx.x; lane = tid & 31; warp = tid >> 5;
        r = tid; y = r & 15; z = r >> 4;
    }

    // per-chunk metadata: read-only for the launch (__ldg) unless the launch runs beside the upload of its own batch (see MeshParams::released)
    template <class T>
    __device__ __forceinline__ T meta(const T* ptr) const {
        if constexpr (LIVE) return __ldcg(ptr);
        else return __ldg(ptr);
    }
    __device__ __forceinline__ uint32_t slot_of(uint32_t chunk) const { return p.slot_list ? meta(p.slot_list + work_index(p, chunk)) : p.first_slot + chunk; }

    // stage A / stage B of the neighbour-mask prefetch for ticket `nxt`, whose slot thread 0 left in s_next[2 + ...] when it issued the tile
    __device__ __forceinline__ void prefetch_codes(uint32_t nxt, int nxt_buf) {
        pf_ready = false;
        pf_code = kNbrNone;
        if (HALO && nxt < p.n && tid < kNbrMasks) pf_code = prefetch_nbr_code(p, s_next[2 + nxt_buf], tid);
    }
    __device__ __forceinline__ void prefetch_masks() {
        if (HALO) pf_val = prefetch_nbr_mask(p, pf_code, tid, &pf_ready);
    }

    // thread 0 only: the first ticket from `t` on that needs the CTA.  Where the chunk summaries are at hand (parity mode, no
    // pre-pass) an all-air chunk -- the reference's palette.len() == 1 early-out, chunk_mesh.rs:18-20 -- ends right here: empty
    // entry, next ticket.  Its tile is never read and the CTA spends no iteration (two barriers, a tile wait) on it.
    __device__ __forceinline__ uint32_t claim_ticket(uint32_t t) {
        if constexpr (!LIVE) {
            if (p.summary == nullptr || !p.summary_air_skip) return t;
            while (t < p.n && (__ldg(p.summary + slot_of(t)) & kSummaryAir)) {
                if (!p.preset) *entry_of(p, t) = MeshEntry{p.arena_offset, 0u, 0u};
                t = gridDim.x + atomicAdd(p.ticket, 1u);
            }
            return t;
        } else {
            for (;;) {
                if (t >= p.n) return t;
                const uint32_t piece = t / p.piece_chunks;  // (at most 32 pieces)
                if (!((released_seen >> piece) & 1u)) {
                    // the chunk is still on its way (host packers -> unpack kernel -> the word read here).  Bounded like every spin in
                    // this library: a host that never releases the rest of its batch becomes an error code, and this CTA takes no more tickets.
                    uint32_t spins = 0;
                    while (ld_acquire_gpu_u32(p.released + piece) == 0u) {
                        __nanosleep(100);
                        if (++spins > (1u << 23)) { atomicOr(p.status, kStatusSpin); return p.n; }
                    }
                    released_seen |= 1u << piece;
                    fence_proxy_async_global();  // the tile about to be fetched by the async proxy was written through the generic one
                }
                if (p.summary == nullptr || !p.summary_air_skip || !(meta(p.summary + slot_of(t)) & kSummaryAir)) return t;
                if (!p.preset) *entry_of(p, t) = MeshEntry{p.arena_offset, 0u, 0u};
                t = gridDim.x + atomicAdd(p.ticket, 1u);
            }
        }
    }

    // thread 0 only: 8 KiB voxel tile of `chunk` -> voxel buffer `buf`, completion on s_bar[buf]
    __device__ __forceinline__ void issue_load(uint32_t chunk, int buf) const {
        const uint32_t slot = slot_of(chunk);
        s_next[2 + buf] = slot;  // (read by prefetch_codes after the next barrier)
        // A chunk known to be one solid block id throughout needs no tile: every row mask is 0xFFFF and every voxel is that id.
        // count_phase rebuilds the tile in shared memory (two STS.128 per thread); the barrier phase completes without bytes.
        uint32_t uni = 0u;
        if (p.summary != nullptr && p.uid != nullptr && (meta(p.summary + slot) & kSummarySolid)) {
            const uint32_t id = meta(p.uid + slot);
            if (id != kUidMixed) uni = 0x10000u | id;
        }
        s_next[4 + buf] = uni;
        if (uni) { mbar_arrive(&s_bar[buf]); return; }
        mbar_arrive_expect_tx(&s_bar[buf], kChunkBytes);
        const uint16_t* src = p.host_tiles != nullptr ? p.host_tiles + static_cast<size_t>(chunk) * kChunkVoxels : p.voxels + static_cast<size_t>(slot) * kChunkVoxels;
        bulk_load(s_vox + buf * kChunkVoxels, src, kChunkBytes, &s_bar[buf]);
#ifdef DSP_L2_PREFETCH
        // tickets are handed out in order, so the tile of chunk + prefetch_ahead will be wanted by some CTA soon: pull it into L2
        const uint64_t ahead = static_cast<uint64_t>(chunk) + p.prefetch_ahead;
        if (p.prefetch_ahead && ahead < p.n) bulk_prefetch_l2(p.voxels + static_cast<size_t>(slot_of(static_cast<uint32_t>(ahead))) * kChunkVoxels, kChunkBytes);
#endif
    }

    // ---- count phase: row occupancy -> six face masks -> per-row count -> block scan.  2 CTA barriers.
    // `skip_empty`: when the chunk is all air (the reference's palette.len() == 1 early-out, chunk_mesh.rs:18-20) only the
    // first barrier is executed and total = cnt = 0 is returned with *empty = true; the caller then does nothing else.
    __device__ __forceinline__ RowFaces count_phase(uint32_t chunk, int buf, bool* empty = nullptr, bool use_prefetch = false) const {
        const uint16_t* vox = s_vox + buf * kChunkVoxels;
        uint32_t nbv = 0u;
        if (HALO) nbv = (use_prefetch && pf_ready) ? pf_val : fetch_nbr_mask(p, slot_of(chunk), tid);  // prefetched during the previous chunk, or fetched now
        (void)use_prefetch;
        uint32_t m;
        const uint32_t uni = s_next[4 + buf];  // written by thread 0 when it "loaded" this buffer, at least one barrier ago
        if (uni) {
            // one solid block id (by the chunk's summary): the tile was not loaded -- this thread writes its row of it for emit_face
            const uint32_t w = (uni & 0xFFFFu) * 0x00010001u;
            uint4* rowp = reinterpret_cast<uint4*>(s_vox + buf * kChunkVoxels + r * 16);
            const int h = (r >> 2) & 1;
            rowp[h] = make_uint4(w, w, w, w);
            rowp[h ^ 1] = make_uint4(w, w, w, w);
            fence_proxy_async_smem();  // the next tile load into this buffer is an async-proxy write behind these generic ones
            m = 0xFFFFu;
        } else {
            const uint4* rowp = reinterpret_cast<const uint4*>(vox + r * 16);
            const int h = (r >> 2) & 1;  // alternate which half is read first: conflict-free LDS.128
            const uint4 qa = rowp[h], qb = rowp[h ^ 1];
            m = (nonzero_mask8(qa) << (8 * h)) | (nonzero_mask8(qb) << (8 * (h ^ 1)));
        }
        s_rowmask[r] = static_cast<uint16_t>(m);
        if (HALO && tid < kNbrMasks) s_rowmask[kRows + tid] = static_cast<uint16_t>(nbv);
        if (empty != nullptr) {
            *empty = __syncthreads_or(m != 0u) == 0;
            if (*empty) return RowFaces{};
        } else {
            __syncthreads();
        }
        RowFaces f;
        {
            const FaceMasks fm = row_face_masks<HALO>(s_rowmask, r, y, z, m);
            f.top = fm.top; f.bot = fm.bot; f.rear = fm.rear; f.front = fm.front; f.right = fm.right; f.left = fm.left;
        }
        f.cnt = __popc(f.top) + __popc(f.bot) + __popc(f.rear) + __popc(f.front) + __popc(f.right) + __popc(f.left);
        uint32_t incl = f.cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wpre = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t v = s_warp[w];
            wpre += w < warp ? v : 0u;
            total += v;
        }
        f.first = wpre + incl - f.cnt;
        f.total = total;
        return f;
    }

    // ---- face records in stream order for the window [q0, q0 + kRecCap): rec = (voxel index << 3) | direction
    //      (chunk_mesh.rs:32, 65-106: voxel ascending; TOP, BOTTOM, REAR, FRONT, RIGHT, LEFT per voxel)
    __device__ __forceinline__ void write_records(const RowFaces& fc, uint32_t q0) const {
        if (!fc.cnt) return;
        const uint32_t rbase = static_cast<uint32_t>(r) << 7;
        if (kRecCap >= kMaxFaces || fc.total <= kRecCap) {
            uint16_t* rp = s_rec + fc.first;
#pragma unroll
            for (int x = 0; x < 16; ++x) {
                const uint32_t bit = 1u << x, v = rbase | (x << 3);
                if (fc.top & bit) *rp++ = static_cast<uint16_t>(v | 0u);
                if (fc.bot & bit) *rp++ = static_cast<uint16_t>(v | 1u);
                if (fc.rear & bit) *rp++ = static_cast<uint16_t>(v | 2u);
                if (fc.front & bit) *rp++ = static_cast<uint16_t>(v | 3u);
                if (fc.right & bit) *rp++ = static_cast<uint16_t>(v | 4u);
                if (fc.left & bit) *rp++ = static_cast<uint16_t>(v | 5u);
            }
        } else {  // windowed: only ordinals in [q0, q0 + kRecCap) are recorded this round
            uint32_t ord = fc.first - q0;  // wraps below the window; the unsigned compare rejects those
#pragma unroll 1
            for (int x = 0; x < 16; ++x) {
                const uint32_t v = rbase | (x << 3);
                const uint32_t six = ((fc.top >> x) & 1u) | (((fc.bot >> x) & 1u) << 1) | (((fc.rear >> x) & 1u) << 2) |
                                     (((fc.front >> x) & 1u) << 3) | (((fc.right >> x) & 1u) << 4) | (((fc.left >> x) & 1u) << 5);
#pragma unroll
                for (uint32_t d = 0; d < 6; ++d)
                    if ((six >> d) & 1u) { if (ord < kRecCap) s_rec[ord] = static_cast<uint16_t>(v | d); ++ord; }
            }
        }
    }

    // ---- expand the records of faces [q0, q1) into vertices and send them to the arena at gout (chunk base)
    __device__ __forceinline__ void emit_window(uint32_t q0, uint32_t q1, uint32_t F, const uint16_t* vox, float cbx, float cby, float cbz,
                                                uint8_t* gout, bool fits) {
        if (Cfg::WARP_EMIT) {
            // warp w expands faces [256 j + 32 w, +32) of every 256-face tile j into its own 3,840-byte slot(s) and
            // sends each slot with one bulk copy (30 full 128-byte lines, line-aligned in the arena); a warp only
            // ever waits for the drain of its own previous store -- no CTA-wide barrier in this loop
            float* wstage = s_stage + warp * (NSTAGE * kWarpSlotWords);
            for (uint32_t f0 = q0 + warp * 32u; f0 < q1; f0 += kTileFaces) {
                float* slot = wstage + (stage_ctr % NSTAGE) * kWarpSlotWords;
                if (lane == 0) bulk_wait_read<NSTAGE - 1>();
                __syncwarp();
                const uint32_t f = f0 + lane;
                if (f < F) {
                    emit_face(slot + lane * kFaceWords, s_rec[f - q0], vox, cbx, cby, cbz, p.uvmap, p.n_blocks);
                } else if (f == F && (F & 1u)) {  // odd face count: pad the last 16-byte unit with zeros
                    *reinterpret_cast<float2*>(slot + lane * kFaceWords) = make_float2(0.0f, 0.0f);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && fits) {
                    const uint32_t nf = min(32u, F - f0);
                    bulk_store(gout + static_cast<size_t>(f0) * kFaceBytes, slot, (nf * kFaceBytes + 15u) & ~15u);
                    bulk_commit();
                }
                ++stage_ctr;
            }
        } else {
            // 256 faces per tile, one thread per face, one bulk copy of up to 30,720 bytes per tile
            for (uint32_t f0 = q0; f0 < q1; f0 += kTileFaces) {
                float* stage = s_stage + (stage_ctr % NSTAGE) * (kTileBytes / 4);
                if (tid == 0) bulk_wait_read<NSTAGE - 1>();  // the store that last used this tile has drained it
                __syncthreads();
                const uint32_t f = f0 + tid;
                if (f < F) {
                    emit_face(stage + tid * kFaceWords, s_rec[f - q0], vox, cbx, cby, cbz, p.uvmap, p.n_blocks);
                } else if (f == F && (F & 1u)) {
                    *reinterpret_cast<float2*>(stage + tid * kFaceWords) = make_float2(0.0f, 0.0f);
                }
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0 && fits) {
                    const uint32_t nf = min(static_cast<uint32_t>(kTileFaces), F - f0);
                    bulk_store(gout + static_cast<size_t>(f0) * kFaceBytes, stage, (nf * kFaceBytes + 15u) & ~15u);
                    bulk_commit();
                }
                ++stage_ctr;
            }
        }
    }

    // records + emission of one counted chunk; `between` runs once after the (first round of) records and before the
    // barrier that publishes them (the halo-mode mask prefetch lives there)
#ifdef DSP_PHASE_TIMERS
    long long ph[16] = {};
    long long tlast = 0;
    __device__ __forceinline__ void tick(int k) { const long long t = clock64(); ph[k] += t - tlast; tlast = t; }
#else
    __device__ __forceinline__ void tick(int) {}
#endif

    // `claimed` (atomic-claim kernel, tile emission): thread 0's slot offset, the result of an atomicAdd issued BEFORE this
    // call and first touched AFTER the record barrier -- only thread 0 issues bulk stores, so nobody else needs the value, and
    // the atomic's round trip hides behind the record phase instead of holding the whole CTA at the barrier (7.8 % of the
    // stall samples sat on that atomic when its result was stored to shared memory before the barrier).
    template <class Between>
    __device__ __forceinline__ void records_and_emit(uint32_t chunk, const RowFaces& fc, const uint16_t* vox, Between between,
                                                     const uint64_t* claimed = nullptr) {
        const uint32_t F = fc.total;
        tick(0);
        // issued first so that its global-memory latency hides behind the record phase (it used to sit right after the
        // barrier, where ncu showed a quarter of all stall samples)
        const int4 cp = meta(&p.chunk_pos[slot_of(chunk)]);
        write_records(fc, 0);
        tick(1);
        between();
        tick(4);
        __syncthreads();  // s_rec, s_base visible
        tick(5);
        uint64_t base = 0;
        if (claimed != nullptr) {
            if (tid == 0) {
                base = *claimed;
                if (!p.preset) {
                    if (base + slot_bytes_of(F) > p.arena_room) atomicOr(p.status, kStatusOverflow);
                    *entry_of(p, chunk) = MeshEntry{p.arena_offset + base, F * 6u, 0u};
                }
            }
        } else {
            base = *s_base;
        }
        const bool fits = base + slot_bytes_of(F) <= p.arena_room;
        uint8_t* gout = p.arena + base;
        const float cbx = __fmul_rn(__int2float_rn(cp.x), 16.0f);  // math.rs:135-139,166-170
        const float cby = __fmul_rn(__int2float_rn(cp.y), 16.0f);
        const float cbz = __fmul_rn(__int2float_rn(cp.z), 16.0f);
        emit_window(0, min(F, kRecCap), F, vox, cbx, cby, cbz, gout, fits);
        if (kRecCap < kMaxFaces && F > kRecCap) more_rounds(fc, vox, cbx, cby, cbz, gout, fits);
        tick(6);
    }

    // rare: a chunk with more faces than the record window takes further record/emit rounds
    __device__ __forceinline__ void more_rounds(const RowFaces& fc, const uint16_t* vox, float cbx, float cby, float cbz, uint8_t* gout, bool fits) {
        const uint32_t F = fc.total;
        for (uint32_t q0 = kRecCap; q0 < F; q0 += kRecCap) {
            __syncthreads();  // the previous round's records are consumed before they are overwritten
            write_records(fc, q0);
            __syncthreads();
            emit_window(q0, min(F, q0 + kRecCap), F, vox, cbx, cby, cbz, gout, fits);
        }
    }
};

// ===== the mesh kernel: single pass, arena space claimed with one atomic per chunk ========================================
// No ordering between chunks at all: thread 0 claims align128(120*F) bytes from a global cursor as soon as the chunk
// is counted, tickets can be taken a whole chunk early (deep tile prefetch) and nothing ever waits on another CTA.  Every
// chunk's stream is byte-exact and the entry table says where it landed; only the ORDER of the slots in the arena follows
// completion order instead of call order.  DSP_MESH_ORDERED runs count_chunks_scan_kernel first and this kernel with p.preset:
// same emission, offsets taken from the entries.
template <class Cfg, bool HALO, bool LIVE = false>
__global__ void __launch_bounds__(256, Cfg::kCtasPerSm) mesh_chunks_bump_kernel(const MeshParams p_in) {
    // Back-to-back mesh calls (a frame loop; bench.py's steps): the next launch's CTAs take their SM slots while this launch's
    // last CTAs drain, and wait here -- before the first global access -- until it has completed (control block reset included).
    pdl_launch_dependents();
    pdl_wait();
    MeshParams p = p_in;
    adopt_work_list(p);  // device-built work list (written by an earlier kernel of the same stream)
    extern __shared__ __align__(128) uint8_t smem[];
    MeshCta<Cfg, HALO, LIVE> c(p, smem);
    const int tid = c.tid;
    if (tid == 0) {
        mbar_init(&c.s_bar[0], 1);
        mbar_init(&c.s_bar[1], 1);
        fence_mbar_init();
        const uint32_t t0 = c.claim_ticket(blockIdx.x);  // first chunk = block index: the tile load starts without waiting for an atomic
        c.s_next[0] = t0;
        if (t0 < p.n) c.issue_load(t0, 0);
    }
    __syncthreads();
    uint32_t cur = c.s_next[0];
    uint32_t it = 0;
#ifdef DSP_PHASE_TIMERS
    c.tlast = clock64();
#endif
    // (Measured alternatives that were NOT faster on B200: consuming the ticket atomic only after the record phase,
    // running tickets two chunks ahead of the tiles; see profiles/README.md.  A static first ticket is worth ~0.3 us.)
    while (cur < p.n) {
        const int buf = it & 1;
        if (tid == 0) {  // next ticket now: its tile lands while this chunk is processed
            const uint32_t t = c.claim_ticket(gridDim.x + atomicAdd(p.ticket, 1u));
            c.s_next[(it + 1) & 1] = t;
            if (p.host_tiles != nullptr) bulk_wait_read<0>();  // the previous chunk's tile has left buf^1 for the voxel store (see below)
            if (t < p.n) c.issue_load(t, buf ^ 1);  // buf^1 was released by the barrier that ended the previous iteration
        }
        c.tick(7);
        seam_wait<HALO>(p, cur);
        mbar_wait(&c.s_bar[buf], (it >> 1) & 1);
        c.tick(2);
        if (p.host_tiles != nullptr && tid == 0) {
            // the tile came straight from host memory: it also becomes the chunk's copy in the voxel store (shared -> global bulk copy,
            // async proxy on both sides; its read of the buffer is awaited before the buffer's next load, at the top of the next iteration)
            bulk_store(p.voxels_out + static_cast<size_t>(c.s_next[2 + buf]) * kChunkVoxels, c.s_vox + buf * kChunkVoxels, kChunkBytes);
            bulk_commit();
        }
#ifndef DSP_NO_EMPTY_SKIP
        bool empty;
        const RowFaces fc = c.count_phase(cur, buf, &empty, true);
        c.tick(3);
        // the next ticket (written by thread 0 at the top of this iteration) is visible behind count_phase's barrier: its
        // neighbour codes start their trip now, its masks after the records -- both land before the next iteration needs them
        c.prefetch_codes(c.s_next[(it + 1) & 1], buf ^ 1);
        if (empty || fc.total == 0u) {
            c.prefetch_masks();
            // Nothing to emit and no slot to claim (all air; or, with cross-chunk culling, a chunk buried among solid
            // neighbours).  No further barrier is needed: every read of vox(buf), the row masks and the warp totals precedes a
            // barrier inside count_phase that all threads have passed, and s_next[(it + 1) & 1] was written before it.
            if (tid == 0 && !p.preset) *entry_of(p, cur) = MeshEntry{p.arena_offset, 0u, 0u};
            ++it;
            cur = c.s_next[it & 1];
            continue;
        }
#else
        const RowFaces fc = c.count_phase(cur, buf);
        c.tick(3);
#endif
        if (!Cfg::WARP_EMIT) {
            // claim the chunk's arena slot now; the result is not touched until the records are written (see records_and_emit)
            uint64_t claimed = 0;
            if (tid == 0) {
                if (p.preset) claimed = ld_volatile_u64(&entry_of(p, cur)->offset_bytes) - p.arena_offset;  // packed in call order by count_chunks_scan_kernel
                else claimed = atomicAdd(reinterpret_cast<unsigned long long*>(p.total_bytes), static_cast<unsigned long long>(slot_bytes_of(fc.total)));
            }
            c.records_and_emit(cur, fc, c.s_vox + buf * kChunkVoxels, [&]() { c.prefetch_masks(); }, &claimed);
        } else {
            c.records_and_emit(cur, fc, c.s_vox + buf * kChunkVoxels, [&]() {
                c.prefetch_masks();
                if (tid == 0) {  // every warp issues its own stores: the offset goes through shared memory
                    const uint64_t bytes = slot_bytes_of(fc.total);
                    if (p.preset) {
                        *c.s_base = ld_volatile_u64(&entry_of(p, cur)->offset_bytes) - p.arena_offset;
                    } else {
                        const uint64_t base = atomicAdd(reinterpret_cast<unsigned long long*>(p.total_bytes), static_cast<unsigned long long>(bytes));
                        *c.s_base = base;
                        if (base + bytes > p.arena_room) atomicOr(p.status, kStatusOverflow);
                        *entry_of(p, cur) = MeshEntry{p.arena_offset + base, fc.total * 6u, 0u};
                    }
                }
            });
        }
        __syncthreads();  // all reads of vox(buf), s_rec, s_next done before they are overwritten
        c.tick(8);
        ++it;
        cur = c.s_next[it & 1];
    }
    if (tid == 0) control_block_epilogue(p);  // control words only: its atomic's round trip overlaps the drain of the last stores
    if (c.lane == 0) bulk_wait<0>();
#ifdef DSP_PHASE_TIMERS
    c.tick(9);
    if (tid == 0 && p.phase_cycles) for (int k = 0; k < 16; ++k) atomicAdd(&p.phase_cycles[k], static_cast<unsigned long long>(c.ph[k]));
    if (tid == 255 && p.phase_cycles) atomicAdd(&p.phase_cycles[15], 1ull);
#endif
}

}  // namespace dsp
