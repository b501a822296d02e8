// Quad modes of the mesher (sm_100a): greedy run-stack merge (DSP_MESH_GREEDY) and/or the indexed layout
// (DSP_MESH_INDEXED).  Extensions named by SURVEY.md section 8f-2; the reference emits one unit quad per visible face
// as six vertices (/root/reference/src/content/world/chunks/chunk_mesh.rs:65-106) and draws non-indexed
// (src/engine/scenes/scene_game.rs:118-122).  The spec lives with the oracle (oracle/mesher_oracle.cpp, "extension
// oracle: greedy run-stack merge"); in short:
//   * per direction the face plane has a run axis u and a stack axis v (TOP/BOTTOM: x,z; REAR/FRONT: x,y; RIGHT/LEFT: y,z);
//   * a RUN is a maximal interval along u of visible faces with equal block id; a QUAD is a maximal chain of
//     identical runs (same start, length, id) in consecutive rows along v;
//   * quads leave in the reference's order applied to their anchor (smallest voxel index): anchor ascending, then
//     TOP, BOTTOM, REAR, FRONT, RIGHT, LEFT; every corner uses the reference's arithmetic for the voxel under it.
// No decision depends on an earlier one, so all 256 rows of a chunk decide at once:
//   1. thread r (x-row r = y + 16 z) holds the six 16-bit face masks of its row and three id-equality masks
//      (towards x-1, y-1, z-1);
//   2. for RIGHT/LEFT the run axis is y: the 16 rows of a z-slice sit in 16 lanes of one warp, so a row learns about its
//      neighbours along y by warp shuffle -- run starts / ends are bit-parallel over the 16 x positions, both directions
//      packed in one register; whether a whole run (lanes y0..y1) agrees with the row one z below is a 4-step
//      segmented OR-scan of shuffles delivered to the run's first lane;
//   3. a run is a (start bit, end bit) pair of the row mask; comparing a row with its predecessor along v gives the
//      mask of CONTINUED runs, anchors = starts & ~continued; height = how many following rows continue the run;
//   4. anchors are counted, scanned (warp shuffle + 8-way), the slot is claimed with one atomic, records
//      (anchor, direction, w, h) are written in stream order and expanded one thread per quad into a staging tile that
//      leaves with one bulk (TMA) store -- the same back half as mesh_chunks_bump_kernel.
// Algorithmic HBM bytes per chunk: 8,192 + 120 Q (stream) or 8,192 + 104 Q (indexed: 4 x 20-byte vertices + 6 u32) + 32.
#pragma once
#include "mesh_kernels.cuh"

namespace dsp {

#ifndef DSP_QUAD_DIRECT_CTAS
#define DSP_QUAD_DIRECT_CTAS 5
#endif
constexpr int kQuadVertexBytesIdx = 80;   // 4 vertices
constexpr int kQuadIndexBytes = 24;       // 6 x u32

// Shared-memory plan.  Greedy output is small (typically 10x fewer quads than faces), so the merge kernel trades
// staging space for residency and tile prefetch depth: 128-quad tiles and a 2,048-record window make room for 4 CTAs per SM, which is what
// hides the per-chunk latencies (tile, barriers, the slot atomic) once there is little emission to hide them behind.
// The unit-quad indexed kernel keeps the 256-quad tile / 4,096-record window of the parity kernel (3 CTAs per SM).
template <bool GREEDY, bool DIRECT = false>  // DIRECT: merged quads in the indexed layout are stored straight to the arena -- no staging tile
struct QuadSmem {
    static constexpr int kCtasPerSm = DIRECT ? DSP_QUAD_DIRECT_CTAS : (GREEDY ? 4 : 3);
    static constexpr int kTileQuads = GREEDY ? 128 : 256;
    static constexpr uint32_t kRecCap = GREEDY ? 2048 : 4096;         // records per window (u32 each)
    static constexpr int kVoxBufs = 2;                                // voxel tile ring: kVoxBufs - 1 tiles in flight per CTA (3 measured slower)
    static constexpr int kT = GREEDY ? kRows * 2 : 0;                 // bytes of one [256] u16 table (none without merging)
    static constexpr int kVox = 0;                                    // kVoxBufs x 8192
    static constexpr int kStage = kVoxBufs * kChunkBytes;                    // kTileQuads x 120 (stream) or x 80 + x 24 (indexed)
    static constexpr int kRec = kStage + (DIRECT ? 0 : kTileQuads * kFaceBytes);  // kRecCap x u32
    static constexpr int kRowMask = kRec + kRecCap * 4;               // 256 x u16 row masks + 96 x u16 neighbour border masks (halo mode) + pad
    static constexpr int kF = kRowMask + kRows * 2 + 256;             // [4][256] u16: TOP, BOTTOM, REAR, FRONT masks per row
    static constexpr int kEqx = kF + 4 * kT;                          // [256] u16
    static constexpr int kSE45 = kEqx + kT;                           // [2][256] u32: run starts / ends along y of RIGHT (low half) and LEFT (high half)
    static constexpr int kC = kSE45 + 4 * kT;                         // [4][256] u16 continued-run masks per row (TOP, BOTTOM, REAR, FRONT)
    static constexpr int kC45 = kC + 4 * kT;                          // [256] u32 continued-run masks of RIGHT | LEFT << 16
    static constexpr int kWarpTot = kC45 + 2 * kT;                      // 8 x u32 (+ pad)
    static constexpr int kBar = kWarpTot + 64;                        // kVoxBufs (<= 4) x u64
    static constexpr int kBase = kBar + 32;                           // u64
    static constexpr int kNext = kBase + 8;                           // kVoxBufs (<= 4) x u32
    static constexpr int kBytes = kNext + 16;
};
static_assert(QuadSmem<true>::kBytes <= 57344 && QuadSmem<false>::kBytes <= 76800 && QuadSmem<true, true>::kBytes <= 45056, "4 / 3 / 5 CTAs per SM");

// bit x = ids equal at position x of two rows of 16 u16 (held as 8 words each): "differs" is nonzero_mask8 of the XOR
__device__ __forceinline__ uint32_t eq_mask16(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    const uint32_t ne = nonzero_mask8(make_uint4(a0.x ^ b0.x, a0.y ^ b0.y, a0.z ^ b0.z, a0.w ^ b0.w)) |
                        (nonzero_mask8(make_uint4(a1.x ^ b1.x, a1.y ^ b1.y, a1.z ^ b1.z, a1.w ^ b1.w)) << 8);
    return ~ne & 0xFFFFu;
}
// bit x (x >= 1) = id[x] == id[x-1] within one row: the row against itself shifted by one voxel (funnel shifts)
__device__ __forceinline__ uint32_t eq_prev_mask16(const uint4 a0, const uint4 a1) {
    const uint4 p0 = make_uint4(a0.x << 16, __funnelshift_l(a0.x, a0.y, 16), __funnelshift_l(a0.y, a0.z, 16), __funnelshift_l(a0.z, a0.w, 16));
    const uint4 p1 = make_uint4(__funnelshift_l(a0.w, a1.x, 16), __funnelshift_l(a1.x, a1.y, 16), __funnelshift_l(a1.y, a1.z, 16), __funnelshift_l(a1.z, a1.w, 16));
    return eq_mask16(a0, a1, p0, p1) & 0xFFFEu;
}

__device__ __forceinline__ uint32_t run_starts(uint32_t f, uint32_t eq) { return f & ~((f << 1) & eq); }
__device__ __forceinline__ uint32_t run_ends(uint32_t f, uint32_t eq) { return f & ~((f >> 1) & (eq >> 1)); }

// Start bits of row b whose run is identical (start, end, block id) to a run of the previous row a.  TWO directions at once:
// the low and high halves of every argument are two independent 16-bit problems that share the id-equality masks (eqa, eqb,
// eqv are the 16-bit masks replicated into both halves; their bit 0 / 16 is clear, so no shift below leaks across the halves,
// and bit 15 of a half is always a run end).  Branch-free: "the two rows' end masks agree on every cell of the run" is a
// segmented suffix-OR of the disagreement towards each run's first cell, four doubling steps over the 16 cells.
__device__ __forceinline__ uint32_t continued_runs2(uint32_t fa, uint32_t eqa, uint32_t fb, uint32_t eqb, uint32_t eqv) {
    const uint32_t eb = run_ends(fb, eqb);
    const uint32_t cand = run_starts(fa, eqa) & run_starts(fb, eqb) & eqv;
    uint32_t acc = (run_ends(fa, eqa) ^ eb) & fb;   // cells of b where the rows disagree about "a run ends here"
    uint32_t open = fb & ~eb;                       // cell x and cell x + 1 belong to the same run of b
    acc |= (acc >> 1) & open;  open &= open >> 1;
    acc |= (acc >> 2) & open;  open &= open >> 2;
    acc |= (acc >> 4) & open;  open &= open >> 4;
    acc |= (acc >> 8) & open;
    return cand & ~acc;
}

// (Q <= 12,288, so every size below fits 32 bits)
__device__ __forceinline__ uint32_t quad_vertex_region_bytes(uint32_t quads, bool indexed) {
    return (quads * static_cast<uint32_t>(indexed ? kQuadVertexBytesIdx : kFaceBytes) + (kSlotAlign - 1)) & ~static_cast<uint32_t>(kSlotAlign - 1);
}
__device__ __forceinline__ uint32_t quad_slot_bytes(uint32_t quads, bool indexed) {
    return quad_vertex_region_bytes(quads, indexed) + (indexed ? ((quads * kQuadIndexBytes + (kSlotAlign - 1)) & ~static_cast<uint32_t>(kSlotAlign - 1)) : 0u);
}

// Corners of one quad record: positions lo/hi per axis and the block's mapped uv row.
struct QuadGeom {
    float lx, hx, ly, hy, lz, hz;
    float4 uv;
    uint32_t code;
};
__device__ __forceinline__ QuadGeom quad_geom_of_id(uint32_t rec, uint32_t id, float cbx, float cby, float cbz,
                                                    const float4* __restrict__ uvmap, uint32_t n_blocks) {
    const uint32_t idx = rec & 0xFFFu, d = (rec >> 12) & 7u, w1 = (rec >> 15) & 15u, h1 = (rec >> 19) & 15u;
    QuadGeom g;
    const uint32_t row = id < n_blocks ? id : n_blocks;  // last row = the reference's fallback (chunk_mesh.rs:53-56)
    g.uv = __ldg(&uvmap[row]);
    const uint32_t ax = idx & 15u, ay = (idx >> 4) & 15u, az = idx >> 8;
    // far cell per axis: TOP/BOTTOM run x, stack z; REAR/FRONT run x, stack y; RIGHT/LEFT run y, stack z
    const uint32_t fx = ax + (d < 4u ? w1 : 0u);
    const uint32_t fy = ay + (d >= 4u ? w1 : ((d & 6u) == 2u ? h1 : 0u));
    const uint32_t fz = az + (d < 2u || d >= 4u ? h1 : 0u);
    // chunk_mesh.rs:61-62,67: template +-0.5 added to ((local as f32) + (chunk as f32 * 16.0)) of the voxel under the corner
    g.lx = __fadd_rn(-0.5f, __fadd_rn(__uint2float_rn(ax), cbx));  g.hx = __fadd_rn(0.5f, __fadd_rn(__uint2float_rn(fx), cbx));
    g.ly = __fadd_rn(-0.5f, __fadd_rn(__uint2float_rn(ay), cby));  g.hy = __fadd_rn(0.5f, __fadd_rn(__uint2float_rn(fy), cby));
    g.lz = __fadd_rn(-0.5f, __fadd_rn(__uint2float_rn(az), cbz));  g.hz = __fadd_rn(0.5f, __fadd_rn(__uint2float_rn(fz), cbz));
    g.code = static_cast<uint32_t>((d < 3 ? kCodes0 >> (20 * d) : kCodes1 >> (20 * (d - 3))) & 0xFFFFFu);
    return g;
}
__device__ __forceinline__ QuadGeom quad_geom(uint32_t rec, const uint16_t* vox, float cbx, float cby, float cbz,
                                              const float4* __restrict__ uvmap, uint32_t n_blocks) {
    return quad_geom_of_id(rec, vox[rec & 0xFFFu], cbx, cby, cbz, uvmap, n_blocks);
}

// Extent bits (w - 1) << 15 | (h - 1) << 19 of the quad anchored at voxel (x, row rr) in direction d, from the chunk's tables in
// shared memory: run length from the row's end mask (RIGHT / LEFT: from the end masks of the rows above it in its z-slice), height =
// how many following rows along the stack axis continue the run -- the continued-run tables hold the 16 entries along the stack
// axis contiguously, so that is a few 128-bit loads, a bit gather and a count of trailing ones; no walk over rows.
__device__ __forceinline__ uint32_t quad_extent(uint32_t rec, const uint16_t* s_F, const uint16_t* s_eqx, const uint32_t* s_SE45, const uint16_t* s_C,
                                                const uint32_t* s_C45) {
    const uint32_t rr = (rec >> 4) & 0xFFu, x = rec & 15u, d = (rec >> 12) & 7u;
    const int zz = rr >> 4, yy = rr & 15;
    uint32_t w1, h1;
    if (d < 4u) {
        w1 = __ffs(run_ends(s_F[d * kRows + rr], s_eqx[rr]) >> x) - 1;
        // the 16 continued-run masks along the stack axis of this quad's column, bit x of each -> one 16-bit word
        const uint4* col = reinterpret_cast<const uint4*>(s_C + d * kRows + (d < 2u ? yy : zz) * 16);
        const uint4 ca = col[0], cb = col[1];
        const uint32_t w[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        uint32_t T = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t b = (w[j] >> x) & 0x00010001u;
            T |= ((b | (b >> 15)) & 3u) << (2 * j);
        }
        h1 = __ffs(~(T >> ((d < 2u ? zz : yy) + 1))) - 1;  // rows after the anchor's that continue the run, up to the first that does not
    } else {
        // RIGHT / LEFT: the run climbs the rows of its z-slice until one carries its end bit (run length along y); it is stacked
        // along z while the rows 16 further on continue it
        const uint32_t bit = x + 16u * (d - 4u);
        const uint4* ecol = reinterpret_cast<const uint4*>(s_SE45 + kRows + zz * 16);
        const uint4* ccol = reinterpret_cast<const uint4*>(s_C45 + yy * 16);
        uint32_t Te = 0u, Tc = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 e = ecol[j], c = ccol[j];
            Te |= (((e.x >> bit) & 1u) | (((e.y >> bit) & 1u) << 1) | (((e.z >> bit) & 1u) << 2) | (((e.w >> bit) & 1u) << 3)) << (4 * j);
            Tc |= (((c.x >> bit) & 1u) | (((c.y >> bit) & 1u) << 1) | (((c.z >> bit) & 1u) << 2) | (((c.w >> bit) & 1u) << 3)) << (4 * j);
        }
        w1 = __ffs(Te >> yy) - 1;  // the run's last row always carries the end bit
        h1 = __ffs(~(Tc >> (zz + 1))) - 1;
    }
    return (w1 << 15) | (h1 << 19);
}

// A chunk of one solid block id under chunk-local culling is exactly six 16 x 16 quads, in stream order BOTTOM, REAR, RIGHT
// anchored at voxel 0, LEFT at (15,0,0), TOP at (0,15,0), FRONT at (0,0,15).  Slot: align128(6 x 120) or align128(6 x 80) + align128(6 x 24).
__host__ __device__ constexpr uint32_t uniform_chunk_slot_bytes(bool indexed) { return indexed ? 512u + 256u : 768u; }
__device__ __forceinline__ uint32_t uniform_chunk_record(int q) {
    return (q == 0 ? (0u | (1u << 12)) : q == 1 ? (0u | (2u << 12)) : q == 2 ? (0u | (4u << 12)) : q == 3 ? (15u | (5u << 12))
            : q == 4 ? (240u | (0u << 12)) : (3840u | (3u << 12))) | (15u << 15) | (15u << 19);
}
// One 32-bit word per thread (t < 180 stream, t < 120 + 36 indexed) straight to the chunk's arena slot `out`.
template <bool INDEXED>
__device__ __forceinline__ void emit_uniform_chunk(uint8_t* out, const int4 cp, uint32_t id, const float4* __restrict__ uvmap, uint32_t n_blocks, int t) {
    constexpr int kVerts = INDEXED ? 4 : 6, kWordsPerQuad = kVerts * 5, kWords = 6 * kWordsPerQuad;
    if (t < kWords) {
        const float cbx = __fmul_rn(__int2float_rn(cp.x), 16.0f), cby = __fmul_rn(__int2float_rn(cp.y), 16.0f), cbz = __fmul_rn(__int2float_rn(cp.z), 16.0f);
        const int q = t / kWordsPerQuad, rem = t - q * kWordsPerQuad, v = rem / 5, c = rem - v * 5;
        const QuadGeom g = quad_geom_of_id(uniform_chunk_record(q), id, cbx, cby, cbz, uvmap, n_blocks);
        const int k = INDEXED ? v : (v < 3 ? v : (v == 3 ? 2 : (v == 4 ? 3 : 0)));  // stream order c0, c1, c2, c2, c3, c0
        const uint32_t cc = g.code >> (5 * k);
        reinterpret_cast<float*>(out)[t] = c == 0 ? ((cc & 1u) ? g.hx : g.lx) : c == 1 ? ((cc & 2u) ? g.hy : g.ly) : c == 2 ? ((cc & 4u) ? g.hz : g.lz)
                                           : c == 3 ? ((cc & 8u) ? g.uv.z : g.uv.x) : ((cc & 16u) ? g.uv.w : g.uv.y);
    } else if (INDEXED && t < kWords + 36) {
        const int w = t - kWords, q = w / 6, j = w - q * 6;
        reinterpret_cast<uint32_t*>(out + 512)[w] = 4u * q + (j < 3 ? j : (j == 3 ? 2 : (j == 4 ? 3 : 0)));
    }
}

template <bool HALO, bool GREEDY, bool INDEXED>
__global__ void __launch_bounds__(256, QuadSmem<GREEDY, GREEDY && INDEXED>::kCtasPerSm) mesh_chunks_quads_kernel(const MeshParams p_in) {
    pdl_launch_dependents();  // (see mesh_chunks_bump_kernel)
    pdl_wait();
    MeshParams p = p_in;
    adopt_work_list(p);  // device-built work list (see classify_chunks_kernel)
    constexpr bool DIRECT = GREEDY && INDEXED;  // merged quads in the indexed layout are stored straight to the arena (see the expansion)
    using L = QuadSmem<GREEDY, DIRECT>;
    constexpr uint32_t kRecCap = L::kRecCap;
    constexpr int kTile = L::kTileQuads;
    extern __shared__ __align__(128) uint8_t smem[];
    uint16_t* const s_vox = reinterpret_cast<uint16_t*>(smem + L::kVox);
    float* const s_stage = reinterpret_cast<float*>(smem + L::kStage);
    uint32_t* const s_rec = reinterpret_cast<uint32_t*>(smem + L::kRec);
    uint16_t* const s_rowmask = reinterpret_cast<uint16_t*>(smem + L::kRowMask);
    uint16_t* const s_F = reinterpret_cast<uint16_t*>(smem + L::kF);
    uint16_t* const s_eqx = reinterpret_cast<uint16_t*>(smem + L::kEqx);
    uint32_t* const s_SE45 = reinterpret_cast<uint32_t*>(smem + L::kSE45);
    uint16_t* const s_C = reinterpret_cast<uint16_t*>(smem + L::kC);
    uint32_t* const s_C45 = reinterpret_cast<uint32_t*>(smem + L::kC45);
    uint32_t* const s_warp = reinterpret_cast<uint32_t*>(smem + L::kWarpTot);
    uint64_t* const s_bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
    uint64_t* const s_base = reinterpret_cast<uint64_t*>(smem + L::kBase);
    uint32_t* const s_next = reinterpret_cast<uint32_t*>(smem + L::kNext);
    float* const s_istage = s_stage + kTile * 20;  // indexed layout: the tile's indices follow its vertices

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = tid, y = r & 15, z = r >> 4;   // thread r owns x-row (y, z): masks over x
    auto slot_of = [&](uint32_t chunk) -> uint32_t { return p.slot_list ? __ldg(p.slot_list + work_index(p, chunk)) : p.first_slot + chunk; };
    auto issue_load = [&](uint32_t chunk, int buf) {
        mbar_arrive_expect_tx(&s_bar[buf], kChunkBytes);
        // (host-resident chunks: the tile of ticket i is pulled straight out of pinned host memory, see MeshParams::host_tiles)
        const uint16_t* src = p.host_tiles != nullptr ? p.host_tiles + static_cast<size_t>(chunk) * kChunkVoxels : p.voxels + static_cast<size_t>(slot_of(chunk)) * kChunkVoxels;
        bulk_load(s_vox + buf * kChunkVoxels, src, kChunkBytes, &s_bar[buf]);
    };

    // Tickets.  The first chunk of a CTA is its block index (no atomic: hundreds of CTAs hitting one counter in the same
    // microsecond serialise in L2); later tickets are gridDim.x + atomicAdd(counter).  The atomic for the ticket after
    // next is issued at the top of an iteration and its result is first touched at the top of the following one, so its
    // round trip is not waited for.  Measured alternatives, all slower or equal on B200 (profiles/README.md): fetching two
    // iterations ahead, a 3-deep tile ring (with static striding to take the atomics out of the picture: no gain, so tile
    // latency is not what bounds short chunks), static striding itself (no atomics, but 15 % slower on uneven regions).
    constexpr int NB = L::kVoxBufs;
    static_assert(NB == 2, "the prologue assigns one buffer");
    // thread 0 only: where the chunk summaries are at hand (no pre-pass), an all-air chunk ends at the ticket -- empty entry, next
    // ticket, no tile (chunk_mesh.rs:18-20; see MeshCta::claim_ticket)
    auto skip_air = [&](uint32_t t) -> uint32_t {
        if (p.summary == nullptr || !p.summary_air_skip) return t;
        while (t < p.n && (__ldg(p.summary + slot_of(t)) & kSummaryAir)) {
            *entry_of(p, t) = MeshEntry{p.arena_offset, 0u, 0u};
            t = gridDim.x + atomicAdd(p.ticket, 1u);
        }
        return t;
    };
    uint32_t pending = 0;  // thread 0: the ticket after the one whose tile is in flight
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) mbar_init(&s_bar[b], 1);
        fence_mbar_init();
        const uint32_t t0 = skip_air(blockIdx.x);
        s_next[0] = t0;
        if (t0 < p.n) issue_load(t0, 0);
        pending = gridDim.x + atomicAdd(p.ticket, 1u);
    }
    if (GREEDY && !HALO && p.n_dev != nullptr) {
        // chunks that classify_greedy_kernel found to be one solid block id (slot and entry are already there): their six quads,
        // a few chunks per CTA, while the first tile is on its way
        const uint32_t n_uni = p.n_dev[4];
        for (uint32_t u = blockIdx.x; u < n_uni; u += gridDim.x) {
            const uint32_t li = p.n_cap - 1u - u;
            const uint32_t slot = __ldg(p.slot_list + li);
            const uint64_t off = p.entries[__ldg(p.entry_map + li)].offset_bytes - p.arena_offset;
            if (off + uniform_chunk_slot_bytes(INDEXED) <= p.arena_room)
                emit_uniform_chunk<INDEXED>(p.arena + off, __ldg(&p.chunk_pos[slot]), p.uid[slot], p.uvmap, p.n_blocks, tid);
        }
    }
    __syncthreads();
    uint32_t cur = s_next[0];
    int buf = 0, par = 0;  // ring position of `cur` and the phase parity of its barrier
    // Thread 0 only: the LAST staging tile of a chunk is not stored the moment it is complete but one barrier into the next
    // iteration (or at loop exit).  The store needs the slot offset, i.e. the result of the claim atomic -- a hot address every
    // chunk of the launch hits, with a round trip that was the longest wait of this kernel (21-26 % of its stall samples, all
    // eight warps parked behind thread 0).  Deferred, the wait overlaps the next chunk's tile wait and occupancy phase; the tile
    // is safe meanwhile, because nothing writes the staging buffer before the next expansion, which starts with bulk_wait_read.
#ifdef DSP_PHASE_TIMERS
    long long ph[16] = {};
    long long tlast = clock64();
#define QTICK(k) do { const long long t_ = clock64(); ph[k] += t_ - tlast; tlast = t_; } while (0)
    // per-CTA timeline (ns, globaltimer): [0] start, [1] prologue done, [2..7] chunk k done, [8] exit, [9] SM, [10] chunks, [11] first tile in
    unsigned long long* const tl = (p.phase_cycles && blockIdx.x < 1024u && tid == 0) ? p.phase_cycles + 32 + blockIdx.x * 16 : nullptr;
    uint32_t tl_chunks = 0;
    auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; };
    if (tl) { tl[0] = gtime(); unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); tl[9] = sm; }
#define QTL(k) do { if (tl) tl[k] = gtime(); } while (0)
#else
#define QTICK(k) do { } while (0)
#define QTL(k) do { } while (0)
#endif
    bool pend = false;
    uint64_t pend_claimed = 0;
    uint32_t pend_Q = 0, pend_f0 = 0, pend_cur = 0;
    auto store_tile = [&](uint64_t claimed, uint32_t Q, uint32_t f0, uint32_t chunk) {
        if (f0 == 0) {
            if (claimed + quad_slot_bytes(Q, INDEXED) > p.arena_room) atomicOr(p.status, kStatusOverflow);
            *entry_of(p, chunk) = MeshEntry{p.arena_offset + claimed, INDEXED ? Q * 4u : Q * 6u, INDEXED ? Q * 6u : 0u};
        }
        if (claimed + quad_slot_bytes(Q, INDEXED) <= p.arena_room) {
            uint8_t* const gout = p.arena + claimed;
            uint8_t* const gidx = gout + quad_vertex_region_bytes(Q, true);
            const uint32_t nq = min(static_cast<uint32_t>(kTile), Q - f0);
            if (INDEXED) {
                bulk_store(gout + static_cast<size_t>(f0) * kQuadVertexBytesIdx, s_stage, nq * kQuadVertexBytesIdx);
                bulk_store(gidx + static_cast<size_t>(f0) * kQuadIndexBytes, s_istage, (nq * kQuadIndexBytes + 15u) & ~15u);
            } else {
                bulk_store(gout + static_cast<size_t>(f0) * kFaceBytes, s_stage, (nq * kFaceBytes + 15u) & ~15u);
            }
            bulk_commit();
        }
    };
    QTICK(8);
    QTL(1);
    while (cur < p.n) {
        const int nb = buf == 0 ? NB - 1 : buf - 1;  // the buffer the previous iteration has just finished with
        uint32_t tnext = 0;
        if (tid == 0) {
            tnext = skip_air(pending);
            s_next[nb] = tnext;
            pending = gridDim.x + atomicAdd(p.ticket, 1u);
        }
        const int4 cp = __ldg(&p.chunk_pos[slot_of(cur)]);
        seam_wait<HALO>(p, cur);
        uint32_t nbv = 0u;
        if (HALO) nbv = fetch_nbr_mask(p, slot_of(cur), tid);  // the neighbours' border masks (global loads in flight during the tile wait)
        mbar_wait(&s_bar[buf], par);
        QTICK(0);
#ifdef DSP_PHASE_TIMERS
        if (tl && tl_chunks == 0) tl[11] = gtime();
#endif
        const uint16_t* vox = s_vox + buf * kChunkVoxels;

        // ---- 1. row occupancy; is the whole chunk one block id?
        uint4 qa, qb;
        uint32_t m;
        {
            const uint4* rowp = reinterpret_cast<const uint4*>(vox + r * 16);
            const int h = (r >> 2) & 1;  // alternate which half is read first: conflict-free LDS.128
            const uint4 q0 = rowp[h], q1 = rowp[h ^ 1];
            qa = h ? q1 : q0;
            qb = h ? q0 : q1;
            m = nonzero_mask8(qa) | (nonzero_mask8(qb) << 8);
        }
        s_rowmask[r] = static_cast<uint16_t>(m);
        if (HALO && tid < kNbrMasks) s_rowmask[kRows + tid] = static_cast<uint16_t>(nbv);
        const uint32_t id0 = vox[0];
        const uint32_t id0w = id0 * 0x00010001u;
        const uint32_t differs = (qa.x ^ id0w) | (qa.y ^ id0w) | (qa.z ^ id0w) | (qa.w ^ id0w) | (qb.x ^ id0w) | (qb.y ^ id0w) | (qb.z ^ id0w) | (qb.w ^ id0w);
        const bool uniform = __syncthreads_and(differs == 0u) != 0;  // also publishes s_rowmask
        // The next ticket's tile goes into the buffer of the PREVIOUS iteration only now: merged quads are expanded straight
        // from the tables and the tile with no barrier behind them, so the barrier above is the first point at which every thread
        // is known to have finished with that buffer.
        if (tid == 0) {
            if (p.host_tiles != nullptr) {
                // the tile that has just arrived from host memory also becomes the chunk's copy in the voxel store; the previous
                // chunk's copy has left buffer nb before that buffer is loaded again
                bulk_wait_read<0>();
                bulk_store(p.voxels_out + static_cast<size_t>(slot_of(cur)) * kChunkVoxels, s_vox + buf * kChunkVoxels, kChunkBytes);
                bulk_commit();
            }
            if (tnext < p.n) issue_load(tnext, nb);
        }
        QTICK(1);
        if (tid == 0 && pend) { store_tile(pend_claimed, pend_Q, pend_f0, pend_cur); pend = false; }  // the previous chunk's last tile
        QTICK(9);
        if (uniform && id0 == 0u) {
            // all air: the reference's palette.len() == 1 early-out (chunk_mesh.rs:18-20)
            if (tid == 0) *entry_of(p, cur) = MeshEntry{p.arena_offset, 0u, 0u};
            if (++buf == NB) { buf = 0; par ^= 1; }
            cur = s_next[buf];
            continue;
        }
        // one solid block id throughout, chunk-local cull: exactly one 16x16 quad per side
        const bool solid_uniform = GREEDY && !HALO && uniform;

        uint32_t A[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        uint32_t eqx = 0u, E45 = 0u;
        FaceMasks fm{};
        uint32_t cnt = 0, first = 0, wpre = 0, Q;
        if (solid_uniform) {
            Q = 6u;
            if (tid < 6) {
                // anchors (x,y,z): BOTTOM, REAR, RIGHT at voxel 0; LEFT at (15,0,0); TOP at (0,15,0); FRONT at (0,0,15)
                const uint32_t rec = tid == 0 ? (0u | (1u << 12)) : tid == 1 ? (0u | (2u << 12)) : tid == 2 ? (0u | (4u << 12))
                                   : tid == 3 ? (15u | (5u << 12)) : tid == 4 ? (240u | (0u << 12)) : (3840u | (3u << 12));
                s_rec[tid] = rec | (15u << 15) | (15u << 19);
            }
        } else {
            // ---- 2. face masks (chunk_mesh.rs:46-51); anchors = faces unless merging
            fm = row_face_masks<HALO>(s_rowmask, r, y, z, m);
            A[0] = fm.top; A[1] = fm.bot; A[2] = fm.rear; A[3] = fm.front; A[4] = fm.right; A[5] = fm.left;
            if (HALO && GREEDY) {
                // cross-chunk culling buries most solid chunks completely: find that out before the merge machinery runs
                if (__syncthreads_or((A[0] | A[1] | A[2] | A[3] | A[4] | A[5]) != 0u) == 0) {
                    if (tid == 0) *entry_of(p, cur) = MeshEntry{p.arena_offset, 0u, 0u};
                    if (++buf == NB) { buf = 0; par ^= 1; }
                    cur = s_next[buf];
                    continue;
                }
            }
            if (GREEDY) {
                eqx = eq_prev_mask16(qa, qb);
                uint32_t eqy = 0u, eqz = 0u;
                if (y > 0) { const uint4* q = reinterpret_cast<const uint4*>(vox + (r - 1) * 16); eqy = eq_mask16(qa, qb, q[0], q[1]); }
                if (z > 0) { const uint4* q = reinterpret_cast<const uint4*>(vox + (r - 16) * 16); eqz = eq_mask16(qa, qb, q[0], q[1]); }
#pragma unroll
                for (int d = 0; d < 4; ++d) s_F[d * kRows + r] = static_cast<uint16_t>(A[d]);
                s_eqx[r] = static_cast<uint16_t>(eqx);
                // RIGHT / LEFT run along y: the row's neighbours along y are the adjacent lanes of its half warp.  Both directions
                // ride in one register (RIGHT low half, LEFT high half); everything is bit-parallel over the 16 x positions.
                const uint32_t f45 = A[4] | (A[5] << 16), eqy2 = eqy | (eqy << 16), eqz2 = eqz | (eqz << 16);
                uint32_t f_prev = __shfl_up_sync(0xFFFFFFFFu, f45, 1), f_next = __shfl_down_sync(0xFFFFFFFFu, f45, 1);
                uint32_t eqy_next = __shfl_down_sync(0xFFFFFFFFu, eqy2, 1);
                if (y == 0) f_prev = 0u;
                if (y == 15) { f_next = 0u; eqy_next = 0u; }
                const uint32_t S45 = f45 & ~(f_prev & eqy2);       // a run starts here: no visible equal-id face one step down in y
                E45 = f45 & ~(f_next & eqy_next);                   // ... ends here
                const uint32_t joins_next = f45 & f_next & eqy_next;  // lane y + 1 belongs to the same run as lane y
                s_SE45[r] = S45;
                s_SE45[kRows + r] = E45;
                QTICK(2);
                __syncthreads();
                QTICK(10);
                // ---- 3. continued runs -> anchors (TOP | BOTTOM << 16 against the row one z below, REAR | FRONT << 16 against y - 1)
                uint32_t Cm01 = 0u, Cm23 = 0u;
                const uint32_t eqx2 = eqx | (eqx << 16);
                if (z > 0) {
                    const uint32_t eqa = s_eqx[r - 16];
                    Cm01 = continued_runs2(s_F[0 * kRows + r - 16] | (static_cast<uint32_t>(s_F[1 * kRows + r - 16]) << 16), eqa | (eqa << 16),
                                           A[0] | (A[1] << 16), eqx2, eqz2);
                }
                if (y > 0) {
                    const uint32_t eqa = s_eqx[r - 1];
                    Cm23 = continued_runs2(s_F[2 * kRows + r - 1] | (static_cast<uint32_t>(s_F[3 * kRows + r - 1]) << 16), eqa | (eqa << 16),
                                           A[2] | (A[3] << 16), eqx2, eqy2);
                }
                const uint32_t Cm[4] = {Cm01 & 0xFFFFu, Cm01 >> 16, Cm23 & 0xFFFFu, Cm23 >> 16};
                // RIGHT / LEFT: a run (lanes y0..y1 of this z-slice) continues the row one z below when that row has a run with
                // the same start, the same id and the same end.  "Same end" = the end masks of the two rows agree on every lane
                // of the run: OR the disagreement over the run and deliver it to the run's first lane -- a segmented suffix
                // OR-scan in four shuffle steps (lane y absorbs lane y + k only while no run boundary lies between them).
                uint32_t C45 = 0u;
                {
                    uint32_t acc = z > 0 ? (E45 ^ s_SE45[kRows + r - 16]) : 0u, open = joins_next;
#pragma unroll
                    for (int k = 1; k < 16; k <<= 1) {
                        const uint32_t acc_k = __shfl_down_sync(0xFFFFFFFFu, acc, k), open_k = __shfl_down_sync(0xFFFFFFFFu, open, k);
                        const bool in_slice = y + k <= 15;
                        acc |= in_slice ? (acc_k & open) : 0u;
                        open &= in_slice ? open_k : 0u;
                    }
                    if (z > 0) C45 = S45 & s_SE45[r - 16] & eqz2 & ~acc;
                }
                // Continued-run tables, laid out so that the 16 entries along a direction's STACK axis are contiguous: a quad's
                // height is then two to four 128-bit loads and a count of trailing ones instead of a walk over up to 15 rows
                // (quad extent below).  TOP / BOTTOM / RIGHT / LEFT stack along z: index y * 16 + z; REAR / FRONT along y: index r.
                const int tz = y * 16 + z;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    s_C[d * kRows + (d < 2 ? tz : r)] = static_cast<uint16_t>(Cm[d]);
                    A[d] = run_starts(A[d], eqx) & ~Cm[d];
                }
                s_C45[tz] = C45;
                A[4] = (S45 & ~C45) & 0xFFFFu;
                A[5] = (S45 & ~C45) >> 16;
            }

            // ---- 4. count, scan
            cnt = __popc(A[0]) + __popc(A[1]) + __popc(A[2]) + __popc(A[3]) + __popc(A[4]) + __popc(A[5]);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_warp[warp] = incl;
            QTICK(3);
            __syncthreads();
            QTICK(11);
            Q = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const uint32_t v = s_warp[w];
                wpre += w < warp ? v : 0u;
                Q += v;
            }
            first = wpre + incl - cnt;
        }
        // claim the chunk's arena slot; the result is first touched after this thread's records are written
        if (Q == 0u) {  // e.g. a chunk buried among solid neighbours under cross-chunk culling: no slot, no records, no emission
            if (tid == 0) *entry_of(p, cur) = MeshEntry{p.arena_offset, 0u, 0u};
            if (++buf == NB) { buf = 0; par ^= 1; }
            cur = s_next[buf];
            continue;
        }
        // (everything about the slot -- size, claim, overflow check, destination pointers -- is thread 0's business alone)
        uint64_t claimed = 0;
        if (tid == 0) claimed = atomicAdd(reinterpret_cast<unsigned long long*>(p.total_bytes), static_cast<unsigned long long>(quad_slot_bytes(Q, INDEXED)));

        // ---- 5./6. records of a window of quads, then their expansion
        for (uint32_t q0 = 0; q0 == 0 || q0 < Q; q0 += kRecCap) {
            if (q0 > 0) __syncthreads();  // the previous window's records are consumed
            if (!GREEDY) {
                if (cnt && Q <= kRecCap) {  // everything fits one window: straight-line predicated appends
                    uint32_t* rp = s_rec + first;
                    const uint32_t rbase = static_cast<uint32_t>(r) << 4;
#pragma unroll
                    for (int x = 0; x < 16; ++x) {
                        const uint32_t bit = 1u << x, v = rbase | x;
#pragma unroll
                        for (uint32_t d = 0; d < 6; ++d)
                            if (A[d] & bit) *rp++ = v | (d << 12);
                    }
                } else if (cnt) {
                    uint32_t ord = first - q0;  // wraps below the window; the unsigned compare rejects those
                    const uint32_t rbase = static_cast<uint32_t>(r) << 4;
#pragma unroll 1
                    for (int x = 0; x < 16; ++x) {
#pragma unroll
                        for (uint32_t d = 0; d < 6; ++d)
                            if ((A[d] >> x) & 1u) { if (ord < kRecCap) s_rec[ord] = (rbase + x) | (d << 12); ++ord; }
                    }
                }
            } else if (!solid_uniform) {
                if (cnt) {
                    // anchors are sparse: visit only the voxels that carry one.  Only (anchor voxel, direction) is recorded here;
                    // the quad's extent is worked out by the thread that expands it (quad_extent), ONE quad per thread -- a row
                    // with many anchors no longer walks all their heights while the other 255 threads wait at the barrier.
                    uint32_t ord = first - q0;
                    const uint32_t rbase = static_cast<uint32_t>(r) << 4;
                    uint32_t any = A[0] | A[1] | A[2] | A[3] | A[4] | A[5];
                    while (any) {
                        const int x = __ffs(any) - 1;
                        any &= any - 1;
#pragma unroll
                        for (uint32_t d = 0; d < 6; ++d) {
                            if (!((A[d] >> x) & 1u)) continue;
                            if (ord < kRecCap) s_rec[ord] = (rbase + x) | (d << 12);
                            ++ord;
                        }
                    }
                }
            }
            QTICK(4);
            if (DIRECT) {
                // Merged quads are few (a couple of hundred per surface chunk): in the indexed layout they go STRAIGHT to the arena
                // from the thread that expands them (five 16-byte stores per quad + its indices) -- no staging tile, no bulk store to
                // wait for between tiles, no barrier per tile; 35.4-36.7 -> 34.3 us on the bench region in one A/B call.  (The
                // six-vertex stream form stays on the staging tile: its 120-byte quads are only 8-byte aligned, and fifteen scattered
                // 8-byte stores per thread cost 12 us more than the bulk store.)  All that is shared is the slot offset: thread 0
                // publishes it (first touch of the claim's result; the atomic has had the record phase to come back), writes the
                // entry, and the barrier that publishes the records publishes it too.
                if (tid == 0 && q0 == 0) {
                    const bool fits = claimed + quad_slot_bytes(Q, INDEXED) <= p.arena_room;
                    if (!fits) atomicOr(p.status, kStatusOverflow);
                    *entry_of(p, cur) = MeshEntry{p.arena_offset + claimed, INDEXED ? Q * 4u : Q * 6u, INDEXED ? Q * 6u : 0u};
                    *s_base = fits ? claimed : ~0ull;
                }
            } else if (tid == 0) {
                bulk_wait_read<0>();  // the stores that last used the staging tile have drained it
            }
            __syncthreads();  // records visible
            QTICK(13);
            const uint32_t q1 = min(Q, q0 + kRecCap);
            if (DIRECT) {
                const uint64_t base = *s_base;
                const float cbx = __fmul_rn(__int2float_rn(cp.x), 16.0f);  // math.rs:135-139,166-170
                const float cby = __fmul_rn(__int2float_rn(cp.y), 16.0f);
                const float cbz = __fmul_rn(__int2float_rn(cp.z), 16.0f);
                uint8_t* const gout = p.arena + base;
                uint8_t* const gidx = gout + quad_vertex_region_bytes(Q, true);
                for (uint32_t i = tid; i < q1 - q0 && base != ~0ull; i += 256) {
                    uint32_t rec = s_rec[i];
                    if (!solid_uniform) rec |= quad_extent(rec, s_F, s_eqx, s_SE45, s_C, s_C45);
                    const QuadGeom g = quad_geom(rec, vox, cbx, cby, cbz, p.uvmap, p.n_blocks);
                    float cx[4], cy[4], cz[4], cu[4], cv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t c = g.code >> (5 * k);
                        cx[k] = (c & 1u) ? g.hx : g.lx;
                        cy[k] = (c & 2u) ? g.hy : g.ly;
                        cz[k] = (c & 4u) ? g.hz : g.lz;
                        cu[k] = (c & 8u) ? g.uv.z : g.uv.x;
                        cv[k] = (c & 16u) ? g.uv.w : g.uv.y;
                    }
                    const uint32_t f = q0 + i;
                    float4* o = reinterpret_cast<float4*>(gout + static_cast<size_t>(f) * kQuadVertexBytesIdx);  // 80 bytes, 16-byte aligned
                    o[0] = make_float4(cx[0], cy[0], cz[0], cu[0]);
                    o[1] = make_float4(cv[0], cx[1], cy[1], cz[1]);
                    o[2] = make_float4(cu[1], cv[1], cx[2], cy[2]);
                    o[3] = make_float4(cz[2], cu[2], cv[2], cx[3]);
                    o[4] = make_float4(cy[3], cz[3], cu[3], cv[3]);
                    uint2* io = reinterpret_cast<uint2*>(gidx + static_cast<size_t>(f) * kQuadIndexBytes);      // 24 bytes, 8-byte aligned
                    const uint32_t v0 = f * 4u;
                    io[0] = make_uint2(v0, v0 + 1u);
                    io[1] = make_uint2(v0 + 2u, v0 + 2u);
                    io[2] = make_uint2(v0 + 3u, v0);
                }
                QTICK(5);
                continue;  // next record window (the barrier at its top separates this window's reads of s_rec from the next one's writes)
            }
            if (GREEDY && !solid_uniform) {
                // Extent of every quad of the window, one quad per thread over ALL 256 threads (phase timers: with the extent
                // worked out inside the 128-thread tile loop, expansion was 45 % of a CTA's time and half the warps idled).
                // Written back into the record.
                for (uint32_t i = tid; i < q1 - q0; i += 256) s_rec[i] |= quad_extent(s_rec[i], s_F, s_eqx, s_SE45, s_C, s_C45);
                __syncthreads();
            }
            QTICK(12);
            for (uint32_t f0 = q0; f0 < q1; f0 += kTile) {
                if (f0 > q0) {  // (for the window's first tile the barrier above did this)
                    if (tid == 0) bulk_wait_read<0>();
                    __syncthreads();
                }
                // one thread per quad (a thread PAIR per quad, each writing half the vertices so that all eight warps work on a
                // 128-quad tile, measured 3-5 % slower: the kernel is issue-bound and the pair recomputes the corner geometry)
                const int qt = tid;
                const bool on = tid < kTile;
                const uint32_t f = f0 + qt;
                if (on && f < Q) {
                    const uint32_t rec = s_rec[f - q0];
                    const float cbx = __fmul_rn(__int2float_rn(cp.x), 16.0f);  // math.rs:135-139,166-170
                    const float cby = __fmul_rn(__int2float_rn(cp.y), 16.0f);
                    const float cbz = __fmul_rn(__int2float_rn(cp.z), 16.0f);
                    const QuadGeom g = quad_geom(rec, vox, cbx, cby, cbz, p.uvmap, p.n_blocks);
                    float cx[4], cy[4], cz[4], cu[4], cv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t c = g.code >> (5 * k);
                        cx[k] = (c & 1u) ? g.hx : g.lx;
                        cy[k] = (c & 2u) ? g.hy : g.ly;
                        cz[k] = (c & 4u) ? g.hz : g.lz;
                        cu[k] = (c & 8u) ? g.uv.z : g.uv.x;
                        cv[k] = (c & 16u) ? g.uv.w : g.uv.y;
                    }
                    if (INDEXED) {
                        float4* o = reinterpret_cast<float4*>(s_stage + qt * 20);  // 80 bytes: 5 conflict-free STS.128
                        o[0] = make_float4(cx[0], cy[0], cz[0], cu[0]);
                        o[1] = make_float4(cv[0], cx[1], cy[1], cz[1]);
                        o[2] = make_float4(cu[1], cv[1], cx[2], cy[2]);
                        o[3] = make_float4(cz[2], cu[2], cv[2], cx[3]);
                        o[4] = make_float4(cy[3], cz[3], cu[3], cv[3]);
                        uint2* io = reinterpret_cast<uint2*>(s_istage + qt * 6);  // 24 bytes: 3 conflict-free STS.64
                        const uint32_t v0 = f * 4u;
                        io[0] = make_uint2(v0, v0 + 1u);
                        io[1] = make_uint2(v0 + 2u, v0 + 2u);
                        io[2] = make_uint2(v0 + 3u, v0);
                    } else {
                        float2* o = reinterpret_cast<float2*>(s_stage + qt * kFaceWords);
                        o[0] = make_float2(cx[0], cy[0]);   o[1] = make_float2(cz[0], cu[0]);   o[2] = make_float2(cv[0], cx[1]);
                        o[3] = make_float2(cy[1], cz[1]);   o[4] = make_float2(cu[1], cv[1]);   o[5] = make_float2(cx[2], cy[2]);
                        o[6] = make_float2(cz[2], cu[2]);   o[7] = make_float2(cv[2], cx[2]);   o[8] = make_float2(cy[2], cz[2]);
                        o[9] = make_float2(cu[2], cv[2]);   o[10] = make_float2(cx[3], cy[3]);  o[11] = make_float2(cz[3], cu[3]);
                        o[12] = make_float2(cv[3], cx[0]);  o[13] = make_float2(cy[0], cz[0]);  o[14] = make_float2(cu[0], cv[0]);
                    }
                } else if (on && f == Q && (Q & 1u)) {  // odd count: pad the last 16-byte unit with zeros (inside the slot's alignment gap)
                    if (INDEXED) *reinterpret_cast<uint2*>(s_istage + qt * 6) = make_uint2(0u, 0u);
                    else *reinterpret_cast<float2*>(s_stage + qt * kFaceWords) = make_float2(0.0f, 0.0f);
                }
                fence_proxy_async_smem();
                QTICK(5);
                __syncthreads();
                QTICK(14);
                if (tid == 0) {
                    if (f0 + kTile >= Q) { pend = true; pend_claimed = claimed; pend_Q = Q; pend_f0 = f0; pend_cur = cur; }  // last tile: deferred
                    else store_tile(claimed, Q, f0, cur);
                }
            }
        }
        // No barrier here: everything the next iteration overwrites (row masks, tables, records, the staging tile, the
        // other voxel buffer) was last read before a barrier that every thread has already passed, and each of those
        // writes sits behind at least one barrier of the next iteration.
        if (++buf == NB) { buf = 0; par ^= 1; }
        cur = s_next[buf];
        QTICK(6);
#ifdef DSP_PHASE_TIMERS
        if (tl) { if (tl_chunks < 6) tl[2 + tl_chunks] = gtime(); tl[10] = ++tl_chunks; }
#endif
    }
    if (tid == 0 && pend) store_tile(pend_claimed, pend_Q, pend_f0, pend_cur);
#ifdef DSP_PHASE_TIMERS
    QTICK(7);
    QTL(8);
    if ((tid == 0 || tid == 255) && p.phase_cycles) for (int k = 0; k < 15; ++k) atomicAdd(&p.phase_cycles[(tid ? 16 : 0) + k], static_cast<unsigned long long>(ph[k]));
    if (tid == 0 && p.phase_cycles) atomicAdd(&p.phase_cycles[15], 1ull);
#endif
    if (tid == 0) control_block_epilogue(p);  // control words only: overlaps the drain of the last stores
    if (lane == 0) bulk_wait<0>();
}

}  // namespace dsp
