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
// One persistent CTA (256 threads) takes chunks from an atomic ticket counter.  Per chunk:
//   1. the 8 KiB voxel tile arrives in shared memory by a 1-D bulk (TMA) copy that was issued
//      while the previous chunk was being processed (2-deep mbarrier pipeline);
//   2. thread r builds the occupancy mask of x-row r, derives the six face masks and its face
//      count; a warp-shuffle + 8-way scan gives every row its first face ordinal and the CTA the
//      chunk's face total F;
//   3. thread 0 publishes align128(120*F) right after the count; the count phase of chunk k+1 runs BEFORE the
//      emission of chunk k, so every ticket's size is published early and nobody's look-back waits on a CTA
//      that is busy writing vertices.  Warp 0 then resolves the chunk's arena offset with a decoupled
//      look-back over the earlier tickets (single pass: voxels are read from HBM exactly once);
//      all rows write 16-bit face records (voxel index, direction) in stream order;
//   4. faces are expanded 256 at a time, one thread per face, into a 30 KiB staging tile and each
//      tile leaves with ONE bulk shared->global copy (cp.async.bulk), i.e. full-line writes only.
// Algorithmic HBM bytes per chunk: 8,192 (voxels) + 120*F (vertices) + 16 (position) + 16 (entry).
#pragma once
#include "dsp_common.cuh"

namespace dsp {

struct MeshEntry {  // == dsp_mesh_entry
    uint64_t offset_bytes;
    uint32_t vertex_count;
    uint32_t reserved;
};

struct MeshParams {
    const uint16_t* voxels;     // [max_chunks][4096]
    const int4* chunk_pos;      // per slot
    const uint32_t* slot_list;  // nullable
    const uint32_t* nbr_slots;  // [max_chunks][6] (halo mode) or nullptr
    const float4* uvmap;        // [n_blocks + 1], last row = fallback
    uint32_t n_blocks;
    uint32_t first_slot;
    uint32_t n;
    uint8_t* arena;             // already offset by arena_offset
    uint64_t arena_offset;
    uint64_t arena_room;        // bytes available from arena_offset
    MeshEntry* entries;         // [n]
    uint64_t* lookback;         // [n], zeroed before launch
    uint32_t* ctrl;             // [0] ticket counter, [1] status bits; zeroed before launch
    uint64_t* total_bytes;      // out
};

constexpr uint32_t kStatusOverflow = 1u;
constexpr uint32_t kStatusSpin = 2u;

constexpr int kTileFaces = 256;
constexpr int kTileBytes = kTileFaces * kFaceBytes;  // 30,720

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

// occupancy bits of 8 consecutive u16 voxels
__device__ __forceinline__ uint32_t nonzero_mask8(const uint4 q) {
    uint32_t m = 0;
    m |= (q.x & 0xFFFFu) ? 1u : 0u;   m |= (q.x >> 16) ? 2u : 0u;
    m |= (q.y & 0xFFFFu) ? 4u : 0u;   m |= (q.y >> 16) ? 8u : 0u;
    m |= (q.z & 0xFFFFu) ? 16u : 0u;  m |= (q.z >> 16) ? 32u : 0u;
    m |= (q.w & 0xFFFFu) ? 64u : 0u;  m |= (q.w >> 16) ? 128u : 0u;
    return m;
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

// Shared-memory carve-up (dynamic).  NSTAGE staging tiles of 30,720 B.
template <int NSTAGE>
struct MeshSmem {
    static constexpr int kVox = 0;                                   // 2 x 8192
    static constexpr int kStage = 2 * kChunkBytes;                   // NSTAGE x 30720
    static constexpr int kRec = kStage + NSTAGE * kTileBytes;        // 12288 x u16
    static constexpr int kRowMask = kRec + kMaxFaces * 2;            // 256 x u16
    static constexpr int kWarpTot = kRowMask + kRows * 2;            // 8 x u32 (+8 pad)
    static constexpr int kBar = kWarpTot + 64;                       // 2 x u64
    static constexpr int kBase = kBar + 16;                          // u64
    static constexpr int kNext = kBase + 8;                          // 2 x u32
    static constexpr int kBytes = kNext + 8;
};

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

constexpr uint64_t kLbAgg = 1ull << 62, kLbPrefix = 2ull << 62, kLbVal = (1ull << 62) - 1;

template <int NSTAGE, bool HALO>
__global__ void __launch_bounds__(256, NSTAGE == 1 ? 3 : 2) mesh_chunks_kernel(const MeshParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    using L = MeshSmem<NSTAGE>;
    uint16_t* s_vox = reinterpret_cast<uint16_t*>(smem + L::kVox);
    float* s_stage = reinterpret_cast<float*>(smem + L::kStage);
    uint16_t* s_rec = reinterpret_cast<uint16_t*>(smem + L::kRec);
    uint16_t* s_rowmask = reinterpret_cast<uint16_t*>(smem + L::kRowMask);
    uint32_t* s_warp = reinterpret_cast<uint32_t*>(smem + L::kWarpTot);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
    uint64_t* s_base = reinterpret_cast<uint64_t*>(smem + L::kBase);
    uint32_t* s_next = reinterpret_cast<uint32_t*>(smem + L::kNext);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = tid, y = r & 15, z = r >> 4;

    auto slot_of = [&](uint32_t ticket) { return p.slot_list ? __ldg(p.slot_list + ticket) : p.first_slot + ticket; };
    auto issue_load = [&](uint32_t ticket, int buf) {  // thread 0 only
        mbar_arrive_expect_tx(&s_bar[buf], kChunkBytes);
        bulk_load(s_vox + buf * kChunkVoxels, p.voxels + static_cast<size_t>(slot_of(ticket)) * kChunkVoxels, kChunkBytes, &s_bar[buf]);
    };

    // ---- count phase: row occupancy -> six face masks -> per-row count -> block scan; publishes the
    //      chunk's byte size for the look-back of later tickets as early as possible.  2 barriers.
    auto count_phase = [&](uint32_t ticket, int buf) -> RowFaces {
        const uint16_t* vox = s_vox + buf * kChunkVoxels;
        uint32_t m;
        {
            const uint4* rowp = reinterpret_cast<const uint4*>(vox + r * 16);
            const int h = (r >> 2) & 1;  // alternate which half is read first: conflict-free LDS.128
            const uint4 qa = rowp[h], qb = rowp[h ^ 1];
            m = (nonzero_mask8(qa) << (8 * h)) | (nonzero_mask8(qb) << (8 * (h ^ 1)));
        }
        s_rowmask[r] = static_cast<uint16_t>(m);
        __syncthreads();
        // chunk_mesh.rs:46-51: outside the chunk counts as air (mask 0) in parity mode
        uint32_t up = y < 15 ? s_rowmask[r + 1] : 0u;
        uint32_t dn = y > 0 ? s_rowmask[r - 1] : 0u;
        uint32_t re = z > 0 ? s_rowmask[r - 16] : 0u;
        uint32_t fr = z < 15 ? s_rowmask[r + 16] : 0u;
        uint32_t xm = 0u, xp = 0u;  // occupancy of the voxel beyond x = 0 / x = 15
        if (HALO) {
            // extension: consult the resident neighbour across each border (DSP_MESH_HALO)
            const uint32_t* nb = p.nbr_slots + static_cast<size_t>(slot_of(ticket)) * 6;
            auto row_mask_of = [&](uint32_t nslot, int nr) -> uint32_t {
                const uint4* q = reinterpret_cast<const uint4*>(p.voxels + static_cast<size_t>(nslot) * kChunkVoxels + nr * 16);
                return nonzero_mask8(__ldg(q)) | (nonzero_mask8(__ldg(q + 1)) << 8);
            };
            if (y == 15) { const uint32_t s = __ldg(nb + 0); if (s != 0xFFFFFFFFu) up = row_mask_of(s, 0 + 16 * z); }
            if (y == 0) { const uint32_t s = __ldg(nb + 1); if (s != 0xFFFFFFFFu) dn = row_mask_of(s, 15 + 16 * z); }
            if (z == 0) { const uint32_t s = __ldg(nb + 2); if (s != 0xFFFFFFFFu) re = row_mask_of(s, y + 16 * 15); }
            if (z == 15) { const uint32_t s = __ldg(nb + 3); if (s != 0xFFFFFFFFu) fr = row_mask_of(s, y); }
            { const uint32_t s = __ldg(nb + 4); if (s != 0xFFFFFFFFu) xm = __ldg(p.voxels + static_cast<size_t>(s) * kChunkVoxels + r * 16 + 15) != 0; }
            { const uint32_t s = __ldg(nb + 5); if (s != 0xFFFFFFFFu) xp = __ldg(p.voxels + static_cast<size_t>(s) * kChunkVoxels + r * 16) != 0; }
        }
        RowFaces f;
        f.top = m & ~up; f.bot = m & ~dn; f.rear = m & ~re; f.front = m & ~fr;
        f.right = m & ~((m << 1) | xm);
        f.left = m & ~((m >> 1) | (xp << 15));
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
        if (tid == 0) st_volatile_u64(&p.lookback[ticket], (ticket == 0 ? kLbPrefix : kLbAgg) | slot_bytes_of(total));
        return f;
    };

    // ---- decoupled look-back (warp 0): exclusive prefix of the byte sizes of all earlier tickets ----------
    auto look_back = [&](uint32_t ticket, uint32_t faces) {
        const uint64_t bytes = slot_bytes_of(faces);
        uint64_t excl = 0;
        if (ticket > 0) {
            int64_t j = static_cast<int64_t>(ticket) - 1;
            uint32_t spins = 0;
            for (;;) {
                const int64_t q = j - lane;
                const uint64_t v = q >= 0 ? ld_volatile_u64(&p.lookback[q]) : kLbPrefix;
                const uint32_t flag = static_cast<uint32_t>(v >> 62);
                const uint32_t pref = __ballot_sync(0xFFFFFFFFu, flag == 2u);
                const uint32_t inval = __ballot_sync(0xFFFFFFFFu, flag == 0u);
                const int first = pref ? __ffs(pref) - 1 : 31;  // lanes 0..first are needed
                const uint32_t need = first == 31 ? 0xFFFFFFFFu : ((2u << first) - 1u);
                if (inval & need) {  // a predecessor has not published yet
                    if (++spins > (1u << 24)) { if (lane == 0) atomicOr(&p.ctrl[1], kStatusSpin); break; }
                    __nanosleep(20);
                    continue;
                }
                uint64_t part = (need >> lane) & 1u ? (v & kLbVal) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
                excl += part;
                if (pref) break;
                j -= 32;
            }
            if (lane == 0) st_volatile_u64(&p.lookback[ticket], kLbPrefix | (excl + bytes));
        }
        if (lane == 0) {
            *s_base = excl;
            if (excl + bytes > p.arena_room) atomicOr(&p.ctrl[1], kStatusOverflow);
            p.entries[ticket] = MeshEntry{p.arena_offset + excl, faces * 6u, 0u};
            if (ticket == p.n - 1) *p.total_bytes = excl + bytes;
        }
    };

    // ---- prologue: first two tickets, first count phase ------------------------------------------------------
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
        const uint32_t t0 = atomicAdd(&p.ctrl[0], 1u);
        s_next[0] = t0;
        if (t0 < p.n) issue_load(t0, 0);
        uint32_t t1 = 0xFFFFFFFFu;
        if (t0 < p.n) { t1 = atomicAdd(&p.ctrl[0], 1u); if (t1 < p.n) issue_load(t1, 1); }
        s_next[1] = t1;
    }
    __syncthreads();
    uint32_t cur = s_next[0];
    if (cur >= p.n) return;
    mbar_wait(&s_bar[0], 0);
    RowFaces fc = count_phase(cur, 0);
    uint32_t it = 0, stage_ctr = 0;

    // Invariant at the top of iteration `it`: chunk `cur` is in voxel buffer it&1 with its count phase done and
    // its size published; the ticket after it is in s_next[(it+1)&1] and its tile is in flight to the other buffer.
    for (;;) {
        const int buf = it & 1;
        const uint16_t* vox = s_vox + buf * kChunkVoxels;
        const uint32_t nxt = s_next[(it + 1) & 1];
        const uint32_t F = fc.total;

        // face records in stream order: rec = (voxel index << 3) | direction (chunk_mesh.rs:32, 65-106)
        if (fc.cnt) {
            uint16_t* rp = s_rec + fc.first;
            const uint32_t rbase = static_cast<uint32_t>(r) << 7;
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
        }

        // count the NEXT chunk before emitting this one, so that its size is published before the long
        // emission phase and nobody's look-back ever waits on a CTA that is busy writing vertices
        RowFaces fn{};
        if (nxt < p.n) {
            mbar_wait(&s_bar[buf ^ 1], ((it + 1) >> 1) & 1);
            fn = count_phase(nxt, buf ^ 1);
        }
        if (warp == 0) look_back(cur, F);
        __syncthreads();  // s_rec, s_base visible

        // expand + bulk store, 256 faces per tile
        const uint64_t base = *s_base;
        const bool fits = base + slot_bytes_of(F) <= p.arena_room;
        uint8_t* gout = p.arena + base;
        const int4 cp = __ldg(&p.chunk_pos[slot_of(cur)]);
        const float cbx = __fmul_rn(__int2float_rn(cp.x), 16.0f);  // math.rs:135-139,166-170
        const float cby = __fmul_rn(__int2float_rn(cp.y), 16.0f);
        const float cbz = __fmul_rn(__int2float_rn(cp.z), 16.0f);
        for (uint32_t f0 = 0; f0 < F; f0 += kTileFaces) {
            float* stage = s_stage + (stage_ctr % NSTAGE) * (kTileBytes / 4);
            if (tid == 0) bulk_wait_read<NSTAGE - 1>();  // the store that last used this tile has drained it
            __syncthreads();
            const uint32_t f = f0 + tid;
            if (f < F) {
                emit_face(stage + tid * kFaceWords, s_rec[f], vox, cbx, cby, cbz, p.uvmap, p.n_blocks);
            } else if (f == F && (F & 1u)) {  // odd face count: pad the last 16-byte unit with zeros
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
        if (nxt >= p.n) break;
        uint32_t t2 = 0;
        if (tid == 0) { t2 = atomicAdd(&p.ctrl[0], 1u); s_next[it & 1] = t2; }
        __syncthreads();  // everyone is done with vox(buf), s_rec and s_base; s_next visible
        if (tid == 0 && t2 < p.n) issue_load(t2, buf);
        cur = nxt;
        fc = fn;
        ++it;
    }
    if (tid == 0) bulk_wait<0>();
}

}  // namespace dsp
