// Terrain fill, palette remap, edit scatter and bookkeeping kernels (sm_100a).
//
// Reference behaviour being replaced (paths under /root/reference/src):
//   content/world/world.rs:35-42          generator chosen by chunk Y
//   content/world/chunks/chunk.rs:86-120  generate_flat / generate_full / generate_empty
// The NOISE / RANDOM / CHECKER kinds are builder-defined integer specs (DESIGN.md "Terrain spec").
//
// Roofline: HBM write bound, 8,192 algorithmic bytes per chunk (one u16 per voxel) plus 16 B of
// position read.  One CTA of 256 threads per chunk; thread t owns x-row r = t (y = r & 15,
// z = r >> 4) and stores its 32 bytes with two 128-bit stores, so a warp writes 1 KiB contiguous.
#pragma once
#include "dsp_common.cuh"
#include "mesh_kernels.cuh"  // MeshEntry
#include "quad_kernels.cuh"  // quad geometry for the uniform-chunk shortcut of classify_greedy_kernel

namespace dsp {

struct TerrainParams {
    uint16_t* voxels;           // [max_chunks][4096]
    const int4* chunk_pos;      // per slot (x,y,z,_)
    const uint32_t* slot_list;  // nullable: chunk i -> slot_list[i], else first_slot + i
    uint32_t first_slot;
    uint32_t n;
    int kind;
    uint32_t seed;
    // box form (dsp_generate_region): when dims.x != 0 chunk i of the box is at origin + (i % dx, (i / dx) % dy, i / (dx dy)),
    // no position is loaded and chunk_pos_out[slot] is written instead
    int3 origin;
    uint3 dims;
    int4* chunk_pos_out;
    uint8_t* summary;  // per slot: kSummaryUnknown / kSummaryAir / kSummarySolid, written where the generator knows (may be nullptr)
    uint16_t* uid;     // per slot, meaningful only while summary has kSummarySolid: the block id every voxel has, or kUidMixed
    uint8_t* bvalid;   // per slot: border masks (MeshParams::bmask) up to date?  Reset here: they are recomputed on demand
};

// Chunk summaries let the halo-mode mesh call drop all-air chunks and chunks buried among solid neighbours before the mesh
// kernel ever sees them (classify_chunks_kernel).  Anything that writes voxels keeps them honest: generators set them,
// edits and host uploads reset them to "unknown".
// Bits: air = every voxel is air; solid = no voxel is air; layer f = the one-voxel border layer at face f (order top,
// bottom, rear, front, right, left) has no air.  0 = nothing known.  A solid chunk is buried, i.e. has no visible face,
// when across every face a RESIDENT neighbour's facing layer is solid.
// (the constants kSummary*, kUidMixed are in mesh_kernels.cuh: the mesh kernels consult the summaries too)

// smoothstep-weighted bilinear value noise on a lattice whose values were hashed once per chunk
// into `lat` (3x3 per octave, [dz][dx]); identical arithmetic to oracle/mesher_oracle.cpp.
__device__ __forceinline__ uint32_t octave_from_lattice(const uint32_t* lat, int32_t wx, int32_t wz, int32_t ix0, int32_t iz0, int o) {
    const int s = 5 - o;
    const int dx = (wx >> s) - ix0, dz = (wz >> s) - iz0;
    const uint32_t tx = static_cast<uint32_t>(wx & ((1 << s) - 1)) << (8 - s);
    const uint32_t tz = static_cast<uint32_t>(wz & ((1 << s) - 1)) << (8 - s);
    const uint32_t sx = (tx * tx * (768U - 2U * tx)) >> 16;
    const uint32_t sz = (tz * tz * (768U - 2U * tz)) >> 16;
    const uint32_t v00 = lat[dz * 3 + dx], v10 = lat[dz * 3 + dx + 1];
    const uint32_t v01 = lat[(dz + 1) * 3 + dx], v11 = lat[(dz + 1) * 3 + dx + 1];
    const uint32_t a = v00 * (256U - sx) + v10 * sx;
    const uint32_t b = v01 * (256U - sx) + v11 * sx;
    return (a * (256U - sz) + b * sz) >> 16;
}

// Store shape: a chunk is 512 half-rows of 8 voxels (16 bytes).  Thread t writes half-rows t and t + 256, so every
// warp-wide store instruction covers 512 CONTIGUOUS bytes (four full 128-byte lines).  (The first version gave each
// thread one whole row = two 16-byte stores 32 bytes apart; each instruction then touched only half of every 32-byte
// sector and the kernel stalled at 3.8-4.0 TB/s of a ~6.8 TB/s write path, profiles/README.md.)
__global__ void __launch_bounds__(256) terrain_fill_kernel(const TerrainParams p) {
    __shared__ uint32_t s_lat[3][9];
    __shared__ int32_t s_h[256];  // [z][x] column heights
    __shared__ uint32_t s_red[8];
    const int tid = threadIdx.x;
    for (uint32_t i = blockIdx.x; i < p.n; i += gridDim.x) {
        const uint32_t slot = p.slot_list ? p.slot_list[i] : p.first_slot + i;
        int4 cp;
        if (p.dims.x != 0u) {
            const uint32_t a = i % p.dims.x, t = i / p.dims.x;
            cp = make_int4(p.origin.x + static_cast<int32_t>(a), p.origin.y + static_cast<int32_t>(t % p.dims.y), p.origin.z + static_cast<int32_t>(t / p.dims.y), 0);
            if (tid == 0) p.chunk_pos_out[slot] = cp;
        } else {
            cp = p.chunk_pos[slot];
        }
        const int32_t wx0 = cp.x * 16;
        if (p.kind == 1) {
            if (tid < 27) {  // lattice values for the three octaves, hashed once per chunk
                const int o = tid / 9, q = tid % 9, s = 5 - o;
                s_lat[o][q] = hash2(p.seed + static_cast<uint32_t>(o), (wx0 >> s) + q % 3, ((cp.z * 16) >> s) + q / 3) & 0xFFU;
            }
            __syncthreads();
            {
                const int x = tid & 15, zc = tid >> 4;
                const int32_t cwx = wx0 + x, cwz = cp.z * 16 + zc;
                uint32_t n = 0;
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    const int s = 5 - o;
                    n += octave_from_lattice(s_lat[o], cwx, cwz, wx0 >> s, (cp.z * 16) >> s, o) << (2 - o);
                }
                s_h[tid] = 16 + static_cast<int32_t>(n >> 3);
            }
            __syncthreads();
        }
        if (p.kind == 2) {
            // hash3(seed, wx, wy, wz) = mix32(hash2(seed, wx, wy) + wz * K): the inner hash2 (three of the four mix32 rounds)
            // is shared by the 16 voxels of a z-column, so it is computed once per (x, y) and kept in shared memory
            s_h[tid] = static_cast<int32_t>(hash2(p.seed, wx0 + (tid & 15), cp.y * 16 + (tid >> 4)));
            __syncthreads();
        }
        uint4* dst = reinterpret_cast<uint4*>(p.voxels + static_cast<size_t>(slot) * kChunkVoxels);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int hr = tid + 256 * part;          // half-row: voxels [8 hr, 8 hr + 8)
            const int row = hr >> 1, x0 = (hr & 1) * 8, y = row & 15, z = row >> 4;
            const int32_t wy = cp.y * 16 + y, wz = cp.z * 16 + z;
            uint32_t w[4];  // 8 voxels, two per word, x ascending
            if (p.kind == 0) {
                // world.rs:35-42: Y > 0 empty, Y == 0 bottom half dirt (y < 8), Y < 0 full.
                const uint32_t id = cp.y > 0 ? 0u : (cp.y == 0 ? (y < 8 ? 1u : 0u) : 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = id | (id << 16);
            } else if (p.kind == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int32_t h0 = s_h[z * 16 + x0 + 2 * k], h1 = s_h[z * 16 + x0 + 2 * k + 1];
                    const uint32_t a = wy < h0 - 1 ? 1u : (wy == h0 - 1 ? 2u : 0u);
                    const uint32_t b = wy < h1 - 1 ? 1u : (wy == h1 - 1 ? 2u : 0u);
                    w[k] = a | (b << 16);
                }
            } else if (p.kind == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t zk = static_cast<uint32_t>(wz) * 0x27D4EB2FU;
                    const uint32_t r0 = mix32(static_cast<uint32_t>(s_h[y * 16 + x0 + 2 * k]) + zk), r1 = mix32(static_cast<uint32_t>(s_h[y * 16 + x0 + 2 * k + 1]) + zk);
                    const uint32_t a = (r0 & 1U) ? 1U + ((r0 >> 1) & 1U) : 0U;
                    const uint32_t b = (r1 & 1U) ? 1U + ((r1 >> 1) & 1U) : 0U;
                    w[k] = a | (b << 16);
                }
            } else {
                // checkerboard: (wx+wy+wz) even -> 1.  wx0 and x0 are even, so parity is local.
                const uint32_t even_first = ((wy + wz) & 1) == 0 ? 1u : 0u;  // voxel at even x
                const uint32_t word = even_first | ((even_first ^ 1u) << 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = word;
            }
            dst[hr] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        uint32_t summary = kSummaryUnknown;
        if (p.kind == 0) {  // flat chunk (Y == 0): y < 8 is dirt, so only its bottom layer is free of air
            summary = cp.y > 0 ? kSummaryAir : (cp.y < 0 ? kSummaryAllSolid : (kSummaryLayer0 << 1));
        } else if (p.kind == 1) {
            // thread t holds the height h of column t = (x, z): voxel (x, y, z) is solid iff wy <= h - 1.  Each condition below
            // must hold for every column (for a side layer: every column of that side), i.e. an AND over the CTA.
            const int32_t h = s_h[tid], top = cp.y * 16 + 15, bot = cp.y * 16;
            const int cx_ = tid & 15, cz_ = tid >> 4;
            uint32_t bits = (bot > h - 1 ? kSummaryAir : 0u) | (top <= h - 1 ? kSummarySolid | kSummaryLayer0 : 0u) | (bot <= h - 1 ? kSummaryLayer0 << 1 : 0u);
            bits |= (cz_ != 0 || top <= h - 1 ? kSummaryLayer0 << 2 : 0u) | (cz_ != 15 || top <= h - 1 ? kSummaryLayer0 << 3 : 0u);
            bits |= (cx_ != 0 || top <= h - 1 ? kSummaryLayer0 << 4 : 0u) | (cx_ != 15 || top <= h - 1 ? kSummaryLayer0 << 5 : 0u);
            bits = __reduce_and_sync(0xFFFFFFFFu, bits);
            if ((tid & 31) == 0) s_red[tid >> 5] = bits;
            __syncthreads();
            summary = s_red[0] & s_red[1] & s_red[2] & s_red[3] & s_red[4] & s_red[5] & s_red[6] & s_red[7];
        }
        if (tid == 0 && p.summary != nullptr) {
            p.bvalid[slot] = 0;
            p.summary[slot] = static_cast<uint8_t>(summary);
            if (summary & kSummarySolid) {
                // reference rule: a full chunk is all "template:dirt" = 1; noise: id 1 below the surface voxel (id 2), so the chunk
                // is uniformly 1 when its top layer lies under every column's surface voxel: top <= h - 2 for all columns
                uint16_t uid = kUidMixed;
                if (p.kind == 0) uid = 1;
                else if (p.kind == 1) { int32_t lo = s_h[0]; for (int k = 1; k < 256; ++k) lo = min(lo, s_h[k]); if (cp.y * 16 + 15 <= lo - 2) uid = 1; }
                p.uid[slot] = uid;
            }
        }
        if (p.kind == 1 || p.kind == 2) __syncthreads();  // s_h / s_lat / s_red reused by the next chunk of this CTA
    }
}

// NOISE terrain for a whole box of chunks (dsp_generate_region).  The heightmap depends on (x, z) only, so one CTA
// takes a chunk column (a, c) and a run of `ny` vertically stacked chunks: the 27 lattice hashes and the 256 column
// heights are computed ONCE and reused for every chunk of the run; per chunk only the row compare/pack and the two
// 128-bit stores per thread remain (coalesced as in terrain_fill_kernel), and rows entirely below / above the surface of their z-slice skip the per-voxel
// compare.  (The per-chunk kernel above spends 2,500 warp-instructions per chunk on the heights and is issue-bound at
// 2.2 TB/s; this one is bound by the 8,192 bytes it writes per chunk.)  Same integer arithmetic as the oracle.
struct RegionFillParams {
    uint16_t* voxels;
    int4* chunk_pos;     // written here as well: slot -> position
    uint32_t first_slot;
    int3 origin;
    uint3 dims;
    uint32_t ny;         // chunks of a column handled by one CTA
    uint32_t seed;
    uint8_t* summary;    // see TerrainParams
    uint16_t* uid;
    uint16_t* bmask;     // [slot][96] border occupancy masks, written here from the heightmap (kBorderMaskStride)
    uint8_t* bvalid;
};

__global__ void __launch_bounds__(256) terrain_fill_region_noise_kernel(const RegionFillParams p) {
    __shared__ uint32_t s_lat[3][9];
    __shared__ int32_t s_h[256];
    __shared__ int32_t s_zmin[16], s_zmax[16];
    const int tid = threadIdx.x;
    const uint32_t ypieces = (p.dims.y + p.ny - 1) / p.ny;
    const uint64_t n_items = static_cast<uint64_t>(p.dims.x) * p.dims.z * ypieces;
    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t a = static_cast<uint32_t>(item % p.dims.x);
        const uint64_t t = item / p.dims.x;
        const uint32_t c = static_cast<uint32_t>(t % p.dims.z), piece = static_cast<uint32_t>(t / p.dims.z);
        const int32_t cx = p.origin.x + static_cast<int32_t>(a), cz = p.origin.z + static_cast<int32_t>(c);
        const int32_t wx0 = cx * 16, wz0 = cz * 16;
        if (tid < 27) {
            const int o = tid / 9, q = tid % 9, s = 5 - o;
            s_lat[o][q] = hash2(p.seed + static_cast<uint32_t>(o), (wx0 >> s) + q % 3, (wz0 >> s) + q / 3) & 0xFFU;
        }
        __syncthreads();
        {
            const int x = tid & 15, zc = tid >> 4;
            uint32_t n = 0;
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                const int s = 5 - o;
                n += octave_from_lattice(s_lat[o], wx0 + x, wz0 + zc, wx0 >> s, wz0 >> s, o) << (2 - o);
            }
            const int32_t h = 16 + static_cast<int32_t>(n >> 3);
            s_h[tid] = h;
            // min / max of the 16 heights of this z-slice: the 16 threads of a slice are half a warp
            int32_t lo = h, hi = h;
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, off));
                hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, off));
            }
            if (x == 0) { s_zmin[zc] = lo; s_zmax[zc] = hi; }
        }
        __syncthreads();
        const uint32_t b0 = piece * p.ny, b1 = min(p.dims.y, b0 + p.ny);
        for (uint32_t b = b0; b < b1; ++b) {
            const int32_t cy = p.origin.y + static_cast<int32_t>(b);
            const uint32_t slot = p.first_slot + a + p.dims.x * (b + p.dims.y * c);
            uint4* dst = reinterpret_cast<uint4*>(p.voxels + static_cast<size_t>(slot) * kChunkVoxels);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const int hr = tid + 256 * part;  // half-row: 512 contiguous bytes per warp-wide store (see terrain_fill_kernel)
                const int row = hr >> 1, x0 = (hr & 1) * 8, y = row & 15, z = row >> 4;
                const int32_t wy = cy * 16 + y;
                const int32_t zlo = s_zmin[z], zhi = s_zmax[z];
                uint32_t w[4];
                if (wy < zlo - 1) {            // whole row below every surface voxel of the slice: body block
#pragma unroll
                    for (int k = 0; k < 4; ++k) w[k] = 0x00010001u;
                } else if (wy > zhi - 1) {     // whole row above the surface: air
#pragma unroll
                    for (int k = 0; k < 4; ++k) w[k] = 0u;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int32_t h0 = s_h[z * 16 + x0 + 2 * k], h1 = s_h[z * 16 + x0 + 2 * k + 1];
                        const uint32_t v0 = wy < h0 - 1 ? 1u : (wy == h0 - 1 ? 2u : 0u);
                        const uint32_t v1 = wy < h1 - 1 ? 1u : (wy == h1 - 1 ? 2u : 0u);
                        w[k] = v0 | (v1 << 16);
                    }
                }
                dst[hr] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            {
                // border occupancy masks straight from the heightmap: voxel (x, y, z) is solid iff cy * 16 + y <= h(x, z) - 1.
                // +-Y layers: thread (x = tid & 15, z = tid >> 4), half-warp ballot = mask[z]; +-Z and +-X layers: thread
                // (bit = tid & 15, k = tid >> 4) = (x, y) resp. (y, z).
                const int lo4 = tid & 15, hi4 = tid >> 4, lane = tid & 31;
                const int32_t top = cy * 16 + 15, bot = cy * 16;
                const int32_t h_xz = s_h[tid];
                const uint32_t b_top = __ballot_sync(0xFFFFFFFFu, top <= h_xz - 1), b_bot = __ballot_sync(0xFFFFFFFFu, bot <= h_xz - 1);
                const int32_t wy_k = cy * 16 + hi4;   // (x = lo4, y = hi4)
                const uint32_t b_rear = __ballot_sync(0xFFFFFFFFu, wy_k <= s_h[lo4] - 1), b_front = __ballot_sync(0xFFFFFFFFu, wy_k <= s_h[15 * 16 + lo4] - 1);
                const int32_t wy_b = cy * 16 + lo4;   // (y = lo4, z = hi4)
                const uint32_t b_right = __ballot_sync(0xFFFFFFFFu, wy_b <= s_h[hi4 * 16] - 1), b_left = __ballot_sync(0xFFFFFFFFu, wy_b <= s_h[hi4 * 16 + 15] - 1);
                if ((lane & 15) == 0) {
                    uint16_t* bm = p.bmask + static_cast<size_t>(slot) * kBorderMaskStride + hi4;
                    const int sh = lane & 16;
                    bm[0] = static_cast<uint16_t>(b_top >> sh);    bm[16] = static_cast<uint16_t>(b_bot >> sh);
                    bm[32] = static_cast<uint16_t>(b_rear >> sh);  bm[48] = static_cast<uint16_t>(b_front >> sh);
                    bm[64] = static_cast<uint16_t>(b_right >> sh); bm[80] = static_cast<uint16_t>(b_left >> sh);
                }
                if (tid == 0) p.bvalid[slot] = 1;
            }
            if (tid == 0) {
                p.chunk_pos[slot] = make_int4(cx, cy, cz, 0);
                if (p.summary != nullptr) {
                    // column-wide extremes of the heightmap: whole chunk, and the four side layers (x = 0 / 15, z = 0 / 15)
                    int32_t lo = s_zmin[0], hi = s_zmax[0], lo_x0 = s_h[0], lo_x15 = s_h[15];
#pragma unroll
                    for (int k = 1; k < 16; ++k) {
                        lo = min(lo, s_zmin[k]); hi = max(hi, s_zmax[k]);
                        lo_x0 = min(lo_x0, s_h[k * 16]); lo_x15 = min(lo_x15, s_h[k * 16 + 15]);
                    }
                    const int32_t top = cy * 16 + 15, bot = cy * 16, lo_z0 = s_zmin[0], lo_z15 = s_zmin[15];
                    uint32_t bits = (bot > hi - 1 ? kSummaryAir : 0u) | (top <= lo - 1 ? kSummarySolid | kSummaryLayer0 : 0u) | (bot <= lo - 1 ? kSummaryLayer0 << 1 : 0u);
                    bits |= (top <= lo_z0 - 1 ? kSummaryLayer0 << 2 : 0u) | (top <= lo_z15 - 1 ? kSummaryLayer0 << 3 : 0u);
                    bits |= (top <= lo_x0 - 1 ? kSummaryLayer0 << 4 : 0u) | (top <= lo_x15 - 1 ? kSummaryLayer0 << 5 : 0u);
                    p.summary[slot] = static_cast<uint8_t>(bits);
                    if (bits & kSummarySolid) p.uid[slot] = top <= lo - 2 ? uint16_t(1) : kUidMixed;  // all body blocks under every surface voxel
                }
            }
        }
        __syncthreads();  // s_h / s_lat are rewritten for the next column
    }
}

// chunk_pos[first_slot + i] for a dims box: i = a + dims.x*(b + dims.y*c) -> origin + (a,b,c)
__global__ void region_positions_kernel(int4* chunk_pos, uint32_t first_slot, int3 origin, uint3 dims, uint64_t n) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t a = static_cast<uint32_t>(i % dims.x);
        const uint64_t t = i / dims.x;
        const uint32_t b = static_cast<uint32_t>(t % dims.y), c = static_cast<uint32_t>(t / dims.y);
        chunk_pos[first_slot + i] = make_int4(origin.x + static_cast<int>(a), origin.y + static_cast<int>(b), origin.z + static_cast<int>(c), 0);
    }
}

// Neighbour slots of every chunk of a box-shaped region (order top +Y, bottom -Y, rear -Z, front +Z, right -X, left +X):
// inside the box it is slot arithmetic; across the box boundary there is no resident neighbour (0xFFFFFFFF) unless a
// received halo slab is scattered over it afterwards.
__global__ void region_neighbours_kernel(uint32_t* nbr, uint32_t first_slot, uint3 dims, uint64_t n) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t a = static_cast<uint32_t>(i % dims.x);
        const uint64_t t = i / dims.x;
        const uint32_t b = static_cast<uint32_t>(t % dims.y), c = static_cast<uint32_t>(t / dims.y);
        const uint32_t s = first_slot + static_cast<uint32_t>(i), sy = dims.x, sz = dims.x * dims.y;
        uint32_t* o = nbr + static_cast<size_t>(s) * 6;
        o[0] = b + 1 < dims.y ? s + sy : 0xFFFFFFFFu;
        o[1] = b > 0 ? s - sy : 0xFFFFFFFFu;
        o[2] = c > 0 ? s - sz : 0xFFFFFFFFu;
        o[3] = c + 1 < dims.z ? s + sz : 0xFFFFFFFFu;
        o[4] = a > 0 ? s - 1 : 0xFFFFFFFFu;
        o[5] = a + 1 < dims.x ? s + 1 : 0xFFFFFFFFu;
    }
}
// positions and slot list of a sub-batch, read directly from pinned host memory (no copy engine involved)
__global__ void stage_positions_kernel(int4* chunk_pos, uint32_t* slot_list_out, uint8_t* summary, uint8_t* bvalid, const int4* host_pos, const uint32_t* host_slots, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = host_slots[i];
    chunk_pos[s] = host_pos[i];
    slot_list_out[i] = s;
    summary[s] = kSummaryUnknown;  // host voxels are about to land in this slot
    bvalid[s] = 0;
}
// The same for a sub-batch whose chunks carry the reference's palette length (dsp_mesh_host_chunks_ex).  A chunk
// with palette.len() == 1 has never held anything but air (chunk.rs:58-65; build_chunk_mesh returns None for it without looking at
// a voxel, chunk_mesh.rs:18-20): its host voxels are not fetched at all -- the slot is zero-filled here and its summary says so,
// which is what lets the mesh kernel finish it at the ticket.
constexpr int kStageChunksPerCta = 64;
__global__ void __launch_bounds__(256) stage_host_chunks_kernel(int4* chunk_pos, uint32_t* slot_list_out, uint8_t* summary, uint8_t* bvalid, uint16_t* voxels,
                                                                const int4* host_pos, const uint32_t* host_slots, const uint8_t* host_air, uint32_t n) {
    // 64 chunks per CTA: their positions / slots / flags are read from the pinned staging block with coalesced loads (one PCIe
    // request per warp, not per chunk), then the CTA zero-fills the slots of the flagged ones, 8 KiB each
    __shared__ uint32_t s_air_slot[kStageChunksPerCta];
    __shared__ uint32_t s_n_air;
    if (threadIdx.x == 0) s_n_air = 0u;
    __syncthreads();
    const uint32_t i = blockIdx.x * kStageChunksPerCta + threadIdx.x;
    if (threadIdx.x < kStageChunksPerCta && i < n) {
        const uint32_t s = host_slots[i];
        const bool air = host_air[i] != 0;
        chunk_pos[s] = host_pos[i];
        slot_list_out[i] = s;
        summary[s] = air ? kSummaryAir : kSummaryUnknown;
        bvalid[s] = 0;
        if (air) s_air_slot[atomicAdd(&s_n_air, 1u)] = s;
    }
    __syncthreads();
    const uint32_t n_air = s_n_air;
    for (uint32_t k = 0; k < n_air; ++k) {
        uint4* d = reinterpret_cast<uint4*>(voxels + static_cast<size_t>(s_air_slot[k]) * kChunkVoxels);
        d[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
        d[threadIdx.x + 256] = make_uint4(0u, 0u, 0u, 0u);
    }
}
// ---- packed upload of host-resident chunks (csrc/host_pack.hpp has the five forms) ------------------------------------------------
// The host's worker threads have packed each chunk into its slot of a pinned staging buffer; this kernel reads what is there over
// PCIe -- nothing for air and uniform chunks, 1 / 2 / 4 KiB for chunks whose ids fit 2 / 4 / 8 bits, 8 KiB otherwise -- and writes the
// 8 KiB u16 tile the mesh kernels read into the chunk's slot of the voxel store.  One warp per chunk, 16 chunks per CTA; the
// per-chunk words (position, slot, descriptor) are read coalesced by the CTA's first 16 threads.  Air and uniform chunks get the
// summary the generators would have left (all air / no air + the one block id), so the parity kernel finishes them at the ticket.
constexpr int kUnpackChunksPerCta = 16;
__device__ __forceinline__ uint4 ldcv_u4(const uint4* p) {  // system memory that the host rewrites between calls: never a stale line
    uint4 v;
    asm volatile("ld.global.cv.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldcv_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(kUnpackChunksPerCta * 32) unpack_host_chunks_kernel(int4* chunk_pos, uint32_t* slot_list_out, uint8_t* summary, uint16_t* uid, uint8_t* bvalid,
                                                                                       uint16_t* voxels, const int4* host_pos, const uint32_t* host_slots,
                                                                                       const uint32_t* host_desc, const uint8_t* host_packed, const uint8_t* host_blocks,
                                                                                       uint32_t n, uint32_t* done_warps, uint32_t* released_flag) {
    __shared__ uint32_t s_slot[kUnpackChunksPerCta], s_desc[kUnpackChunksPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const uint32_t i = blockIdx.x * kUnpackChunksPerCta + threadIdx.x;
        if (threadIdx.x < kUnpackChunksPerCta && i < n) {
            const uint32_t s = ldcv_u32(host_slots + i), d = ldcv_u32(host_desc + i);
            const uint4 pq = ldcv_u4(reinterpret_cast<const uint4*>(host_pos + i));
            chunk_pos[s] = make_int4(static_cast<int>(pq.x), static_cast<int>(pq.y), static_cast<int>(pq.z), 0);
            slot_list_out[i] = s;
            const uint32_t form = d & 7u;
            summary[s] = form == 0u ? kSummaryAir : form == 1u ? kSummaryAllSolid : kSummaryUnknown;
            if (form == 1u) uid[s] = static_cast<uint16_t>(d >> 16);
            bvalid[s] = 0;
            s_slot[threadIdx.x] = s;
            s_desc[threadIdx.x] = d;
        }
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * kUnpackChunksPerCta + warp;
    // released_flag != nullptr: a mesh launch running beside this kernel waits for the piece (MeshParams::released).  Every warp counts
    // itself out behind its last store; the last one of the grid sets the word.
    auto count_out = [&] {
        if (released_flag == nullptr) return;
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            if (atomicAdd(done_warps, 1u) + 1u == gridDim.x * kUnpackChunksPerCta) { __threadfence(); st_release_gpu_u32(released_flag, 1u); }
        }
    };
    if (i >= n) { count_out(); return; }
    const uint32_t d = s_desc[warp], form = d & 7u;
    uint4* tile = reinterpret_cast<uint4*>(voxels + static_cast<size_t>(s_slot[warp]) * kChunkVoxels);  // 512 x 16 bytes
    // (form 6: the host did not touch the chunk -- its 8 KiB are read where the caller holds them, host_blocks = the caller's array)
    const uint8_t* src = (form == 6u ? host_blocks : host_packed) + static_cast<size_t>(i) * 8192;
    if (form <= 1u) {
        const uint32_t w = form == 0u ? 0u : (d >> 16) * 0x00010001u;
#pragma unroll
        for (int j = 0; j < 16; ++j) tile[lane + 32 * j] = make_uint4(w, w, w, w);
    } else if (form == 2u) {  // 2 bits per voxel: 16 bytes = 64 voxels = 8 tile units
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int u = lane + 32 * j;
            const uint4 q = ldcv_u4(reinterpret_cast<const uint4*>(src) + u);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 32 bits = 16 voxels = two tile units
                uint32_t o[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) o[t] = ((w[k] >> (4 * t)) & 3u) | (((w[k] >> (4 * t + 2)) & 3u) << 16);
                tile[u * 8 + k * 2] = make_uint4(o[0], o[1], o[2], o[3]);
                tile[u * 8 + k * 2 + 1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    } else if (form == 3u) {  // 4 bits per voxel: 16 bytes = 32 voxels = 4 tile units
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int u = lane + 32 * j;
            const uint4 q = ldcv_u4(reinterpret_cast<const uint4*>(src) + u);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 32 bits = 8 voxels = one tile unit
                uint32_t o[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) o[t] = ((w[k] >> (8 * t)) & 15u) | (((w[k] >> (8 * t + 4)) & 15u) << 16);
                tile[u * 4 + k] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    } else if (form == 4u) {  // 8 bits per voxel: 16 bytes = 16 voxels = 2 tile units
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int u = lane + 32 * j;
            const uint4 q = ldcv_u4(reinterpret_cast<const uint4*>(src) + u);
            tile[u * 2] = make_uint4(__byte_perm(q.x, 0u, 0x4140), __byte_perm(q.x, 0u, 0x4342), __byte_perm(q.y, 0u, 0x4140), __byte_perm(q.y, 0u, 0x4342));
            tile[u * 2 + 1] = make_uint4(__byte_perm(q.z, 0u, 0x4140), __byte_perm(q.z, 0u, 0x4342), __byte_perm(q.w, 0u, 0x4140), __byte_perm(q.w, 0u, 0x4342));
        }
    } else {  // raw
#pragma unroll 4
        for (int j = 0; j < 16; ++j) tile[lane + 32 * j] = ldcv_u4(reinterpret_cast<const uint4*>(src) + lane + 32 * j);
    }
    count_out();
}

__global__ void scatter_pos_kernel(int4* chunk_pos, uint8_t* summary, uint8_t* bvalid, const uint32_t* slots, const int4* src, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    chunk_pos[slots[i]] = src[i];
    summary[slots[i]] = kSummaryUnknown;  // the slot is being (re)defined; a generator that follows sets it again
    bvalid[slots[i]] = 0;
}

// dsp_invalidate_chunks: the caller wrote voxels behind the library's back (dsp_voxel_device_ptr); forget what is known
__global__ void invalidate_summaries_kernel(uint8_t* summary, uint8_t* bvalid, const uint32_t* slots, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { summary[slots[i]] = kSummaryUnknown; bvalid[slots[i]] = 0; }
}

// ---- border occupancy masks (MeshParams::bmask) on demand ------------------------------------------------------------------
// Writers that know the terrain analytically fill a chunk's masks as they go (terrain_fill_region_noise_kernel); everybody else
// -- the other generators, host uploads, edits, dsp_invalidate_chunks -- clears the chunk's `bvalid` byte, and a halo-mode mesh
// call first brings the masks of the chunks it touches (its own and their resident neighbours) up to date:
//   find_stale_masks_kernel   one thread per chunk of the call: every chunk among {self, 6 neighbours} whose byte is 0 is queued
//                             once (atomicOr of a "queued" bit on the byte's word) into a list;
//   refresh_border_masks_kernel   one CTA per queued chunk: row occupancy of the 256 x-rows from the voxel tile -> the six layers.
// bvalid byte: bit 0 = masks valid, bit 1 = queued.
__global__ void find_stale_masks_kernel(uint8_t* bvalid, const uint32_t* nbr, const uint32_t* slot_list, uint32_t first_slot, uint32_t n, uint32_t* list,
                                        uint32_t* count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = slot_list ? slot_list[i] : first_slot + i;
#pragma unroll
    for (int f = -1; f < 6; ++f) {
        const uint32_t s = f < 0 ? slot : nbr[static_cast<size_t>(slot) * 6 + f];
        if (s == 0xFFFFFFFFu || (s & 0x80000000u)) continue;  // none / lives on another GPU (its layer arrives through the seam)
        if (bvalid[s] != 0) continue;                          // valid, or queued by somebody else
        uint32_t* word = reinterpret_cast<uint32_t*>(bvalid + (s & ~3u));
        const uint32_t bit = 2u << (8u * (s & 3u));
        if ((atomicOr(word, bit) & (3u << (8u * (s & 3u)))) == 0u) list[atomicAdd(count, 1u)] = s;
    }
}
// The kernel reads the whole tile for that, so it also leaves the chunk's SUMMARY (all air / no air / which border layers are free
// of air) and, for a chunk without air, its one block id: a world that came from the host (dsp_load_chunks, dsp_mesh_host_chunks*)
// or was edited has no summaries, and without them the halo pre-pass could drop nothing -- after the first halo-mode call over
// it, such a world is classified like a generated one (all-air and buried chunks never reach the mesh kernel).
__global__ void __launch_bounds__(256) refresh_border_masks_kernel(const uint16_t* voxels, uint16_t* bmask, uint8_t* bvalid, uint8_t* summary, uint16_t* uid,
                                                                   const uint32_t* list, const uint32_t* count) {
    __shared__ uint32_t s_red[8];
    const int tid = threadIdx.x, y = tid & 15, z = tid >> 4, lane = tid & 31;
    const uint32_t n = *count;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t slot = list[i];
        const uint4* rowp = reinterpret_cast<const uint4*>(voxels + static_cast<size_t>(slot) * kChunkVoxels + tid * 16);
        const uint4 q0 = rowp[0], q1 = rowp[1];
        const uint32_t m = nonzero_mask8(q0) | (nonzero_mask8(q1) << 8);  // occupancy of x-row (y, z)
        {
            const uint32_t id0 = voxels[static_cast<size_t>(slot) * kChunkVoxels], w = id0 * 0x00010001u;
            const uint32_t differs = (q0.x ^ w) | (q0.y ^ w) | (q0.z ^ w) | (q0.w ^ w) | (q1.x ^ w) | (q1.y ^ w) | (q1.z ^ w) | (q1.w ^ w);
            const bool full = m == 0xFFFFu;
            // one bit per statement that must hold for EVERY row: all air, no air, layer f (top, bottom, rear, front, right = x 0, left = x 15) free of air, one id
            uint32_t bits = (m == 0u ? kSummaryAir : 0u) | (full ? kSummarySolid : 0u) | ((y != 15 || full) ? kSummaryLayer0 : 0u) | ((y != 0 || full) ? kSummaryLayer0 << 1 : 0u) |
                            ((z != 0 || full) ? kSummaryLayer0 << 2 : 0u) | ((z != 15 || full) ? kSummaryLayer0 << 3 : 0u) | ((m & 1u) ? kSummaryLayer0 << 4 : 0u) |
                            ((m >> 15) & 1u ? kSummaryLayer0 << 5 : 0u) | (differs == 0u ? 0x100u : 0u);
            bits = __reduce_and_sync(0xFFFFFFFFu, bits);
            if (lane == 0) s_red[tid >> 5] = bits;
            __syncthreads();
            if (tid == 0) {
                const uint32_t all = s_red[0] & s_red[1] & s_red[2] & s_red[3] & s_red[4] & s_red[5] & s_red[6] & s_red[7];
                summary[slot] = static_cast<uint8_t>(all & 0xFFu);
                if (all & kSummarySolid) uid[slot] = (all & 0x100u) ? static_cast<uint16_t>(id0) : kUidMixed;
            }
            __syncthreads();  // s_red is rewritten for the CTA's next chunk
        }
        uint16_t* bm = bmask + static_cast<size_t>(slot) * kBorderMaskStride;
        if (y == 15) bm[0 * 16 + z] = static_cast<uint16_t>(m);
        if (y == 0) bm[1 * 16 + z] = static_cast<uint16_t>(m);
        if (z == 0) bm[2 * 16 + y] = static_cast<uint16_t>(m);
        if (z == 15) bm[3 * 16 + y] = static_cast<uint16_t>(m);
        const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, m & 1u), b15 = __ballot_sync(0xFFFFFFFFu, (m >> 15) & 1u);  // bit y of mask[z], two z per warp
        if ((lane & 15) == 0) {
            bm[4 * 16 + z] = static_cast<uint16_t>(b0 >> (lane & 16));
            bm[5 * 16 + z] = static_cast<uint16_t>(b15 >> (lane & 16));
        }
        if (tid == 0) bvalid[slot] = 1;
    }
}

// Halo-mode pre-pass: chunk i of the call needs no meshing when it is all air, or free of air with six RESIDENT neighbours
// whose facing border layers are free of air (every face culled); its entry is written here.  Everything else is appended to the work list that the
// mesh kernel then consumes straight from device memory (count in *n_work; order within the list does not matter, the arena
// slots are claimed in completion order anyway).  One warp-aggregated atomic per warp.
__global__ void classify_chunks_kernel(const uint8_t* summary, const uint32_t* nbr, const uint32_t* slot_list, uint32_t first_slot, uint32_t n,
                                       uint64_t arena_offset, MeshEntry* entries, uint32_t* work_slots, uint32_t* work_entry, uint32_t* n_work) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool work = false, seam = false;
    uint32_t slot = 0;
    if (i < n) {
        slot = slot_list ? slot_list[i] : first_slot + i;
        const uint8_t s = summary[slot];
        bool skip = (s & kSummaryAir) != 0;
        if (nbr != nullptr) {  // (nbr == nullptr: chunk-local culling, only all-air chunks can be dropped)
            bool buried = (s & kSummarySolid) != 0;
#pragma unroll
            for (int f = 0; f < 6; ++f) {  // the neighbour across face f shows us its layer at the opposite face f ^ 1
                const uint32_t q = nbr[static_cast<size_t>(slot) * 6 + f];
                const bool external = q != 0xFFFFFFFFu && (q & 0x80000000u) != 0u;
                seam = seam || (external && ((q >> kNbrSrcShift) & 7u) != 0u);  // a dsp_seam_* inbox: filled by another GPU, maybe not yet
                buried = buried && q != 0xFFFFFFFFu && !external && (summary[q] & (kSummaryLayer0 << (f ^ 1))) != 0;
            }
            skip = skip || buried;
        }
        if (skip) entries[i] = MeshEntry{arena_offset, 0u, 0u};
        work = !skip;
        seam = seam && work;
    }
    // two warp-aggregated appends: local chunks grow the list from the front (count n_work[0]), seam chunks from the back
    // (count n_work[3], position n - 1 - k): the mesh kernel's tickets reach the seam chunks last
    const int lane = threadIdx.x & 31;
    const uint32_t b_front = __ballot_sync(0xFFFFFFFFu, work && !seam), b_back = __ballot_sync(0xFFFFFFFFu, seam);
    if (b_front) {
        const int leader = __ffs(b_front) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(n_work, static_cast<uint32_t>(__popc(b_front)));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (work && !seam) {
            const uint32_t k = base + __popc(b_front & ((1u << lane) - 1u));
            work_slots[k] = slot;
            work_entry[k] = i;
        }
    }
    if (b_back) {
        const int leader = __ffs(b_back) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(n_work + 3, static_cast<uint32_t>(__popc(b_back)));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (seam) {
            const uint32_t k = n - 1u - (base + __popc(b_back & ((1u << lane) - 1u)));
            work_slots[k] = slot;
            work_entry[k] = i;
        }
    }
}
// Pre-pass of DSP_MESH_GREEDY calls with chunk-local culling.  Most chunks of a terrain are uniform: all air (no mesh, the
// reference's palette.len() == 1 early-out, chunk_mesh.rs:18-20) or one solid block id throughout (exactly six 16x16 quads, one
// per side, whatever the id).  Where the generators left a summary that says so, the chunk never reaches the mesh kernel and
// its 8 KiB tile is never read: air chunks get their empty entry here; a uniform solid chunk gets its arena slot and entry here
// and its six quads from the mesh kernel's prologue (emit_uniform_chunk: same order and arithmetic as the in-kernel uniform
// path).  Everything else -- surface chunks, and every chunk whose summary is unknown (host uploads, edits) -- goes to the
// front of the work list.
struct ClassifyGreedyParams {
    const uint8_t* summary;
    const uint16_t* uid;
    const int4* chunk_pos;
    const uint32_t* slot_list;
    uint32_t first_slot, n;
    uint64_t arena_offset, arena_room;
    uint8_t* arena;                 // already offset by arena_offset
    MeshEntry* entries;
    uint32_t* work_slots;
    uint32_t* work_entry;
    uint32_t* n_work;
    unsigned long long* total_bytes;  // the arena cursor of the call
    uint32_t* status;
    const float4* uvmap;
    uint32_t n_blocks;
};
template <bool INDEXED>
__global__ void __launch_bounds__(256) classify_greedy_kernel(const ClassifyGreedyParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int kind = 0;  // 0: nothing (i >= n or skipped), 1: work list, 2: uniform solid
    uint32_t slot = 0, uid = 0;
    if (i < p.n) {
        slot = p.slot_list ? p.slot_list[i] : p.first_slot + i;
        const uint8_t s = p.summary[slot];
        if (s & kSummaryAir) p.entries[i] = MeshEntry{p.arena_offset, 0u, 0u};
        else if ((s & kSummarySolid) && (uid = p.uid[slot]) != kUidMixed) kind = 2;
        else kind = 1;
    }
    const uint32_t b_work = __ballot_sync(0xFFFFFFFFu, kind == 1);
    if (b_work) {
        const int leader = __ffs(b_work) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(p.n_work, static_cast<uint32_t>(__popc(b_work)));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (kind == 1) {
            const uint32_t k = base + __popc(b_work & ((1u << lane) - 1u));
            p.work_slots[k] = slot;
            p.work_entry[k] = i;
        }
    }
    // Uniform solid chunks: their arena slots are claimed here (one atomic per warp on the cursor the mesh kernel uses too) and
    // their entries written; the chunks themselves go to the BACK of the work list (count n_work[4], position n - 1 - k) and
    // the mesh kernel's CTAs write their six quads in a short prologue, spread over the whole grid, without loading a tile.
    const uint32_t b_uni = __ballot_sync(0xFFFFFFFFu, kind == 2);
    if (b_uni == 0u) return;
    constexpr uint32_t kSlot = uniform_chunk_slot_bytes(INDEXED);
    const uint32_t n_uni = __popc(b_uni);
    const int leader = __ffs(b_uni) - 1;
    unsigned long long wbase = 0;
    uint32_t lbase = 0;
    if (lane == leader) {
        wbase = atomicAdd(p.total_bytes, static_cast<unsigned long long>(n_uni) * kSlot);
        lbase = atomicAdd(p.n_work + 4, n_uni);
        if (wbase + static_cast<unsigned long long>(n_uni) * kSlot > p.arena_room) atomicOr(p.status, kStatusOverflow);
    }
    wbase = __shfl_sync(0xFFFFFFFFu, wbase, leader);
    lbase = __shfl_sync(0xFFFFFFFFu, lbase, leader);
    if (kind == 2) {
        const uint32_t k = __popc(b_uni & ((1u << lane) - 1u));
        p.entries[i] = MeshEntry{p.arena_offset + wbase + static_cast<unsigned long long>(k) * kSlot, INDEXED ? 24u : 36u, INDEXED ? 36u : 0u};
        const uint32_t li = p.n - 1u - (lbase + k);
        p.work_slots[li] = slot;
        p.work_entry[li] = i;
    }
}

// neighbour table entry (slots[i], face) := external slab first_index + i of source `src` (0: dsp_set_halo_slabs store,
// 1 + s: inbox of seam s), unless a resident neighbour is already there
__global__ void apply_halo_batch_kernel(uint32_t* nbr, const uint32_t* slots, int face, uint32_t src, uint32_t first_index, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slots[i];
    if (s == 0xFFFFFFFFu) return;  // unloaded since
    uint32_t* e = nbr + static_cast<size_t>(s) * 6 + face;
    if (*e == 0xFFFFFFFFu) *e = kNbrExternal | (src << kNbrSrcShift) | (first_index + i);
}

// ---- region seams between GPUs (dsp_seam_*) ------------------------------------------------------------------------
// Inbox of one seam, allocated by the RECEIVING context and mapped into the sender (same process: peer access; another
// process: CUDA IPC).  The sender writes everything in it: its border layers for epoch e into data[e & 1], then `flag` = e
// (system-scope release); `ack` = the last epoch of OUR layers the peer has finished reading.
struct SeamInbox {
    uint32_t flag;       // newest complete epoch in data[]
    uint32_t pad0[31];
    uint32_t ack;        // the peer is done with the layers we sent it for every epoch <= ack
    uint32_t pad1[31];
    // uint16_t data[2][n][256] follows at byte 256
};
constexpr size_t kSeamInboxHeader = 256;

// One launch per seam and exchange.  (1) tell the peer which of its epochs we are done with: everything on this stream
// before this launch -- in particular every mesh launch that read the peer's layers of epoch e - 1 -- has completed, so
// ack = e - 1; (2) wait until the peer has released the buffer we are about to overwrite (it last held our epoch e - 2);
// (3) write this epoch's layers straight into the peer's inbox over NVLink (plain coalesced stores through the peer
// mapping: 512 B per chunk face); (4) the last CTA publishes flag = e behind a system-scope fence.
struct SeamPushParams {
    const uint16_t* voxels;
    const uint32_t* slots;    // this context's boundary chunks at the seam, canonical order
    uint32_t n;
    int face;
    uint32_t epoch;
    SeamInbox* peer;          // the neighbour's inbox (peer-mapped)
    const SeamInbox* mine;    // our inbox: the peer writes `ack` here
    uint32_t* done;           // CTA counter (local), zeroed by the last CTA
    uint32_t* status;         // context status word: kStatusSeam on time-out
    uint32_t epoch_for_ack_wait_off;  // non-zero: `peer` is a local send buffer for a stream-ordered transport (NCCL): no ack protocol
};
__global__ void seam_flag_kernel(SeamInbox* inbox, uint32_t epoch) { st_release_sys_u32(&inbox->flag, epoch); }
__global__ void __launch_bounds__(256) seam_push_kernel(const SeamPushParams p) {
    const int t = threadIdx.x, a = t >> 4, b = t & 15;
    if (blockIdx.x == 0 && t == 0) st_release_sys_u32(&p.peer->ack, p.epoch - 1u);
    if (p.epoch >= 3u && p.epoch_for_ack_wait_off == 0u) {
        uint32_t spins = 0;
        while (static_cast<int32_t>(ld_acquire_sys_u32(&p.mine->ack) - (p.epoch - 2u)) < 0) {
            __nanosleep(200);
            if (++spins > (1u << 22)) { if (t == 0) atomicOr(p.status, kStatusSeam); break; }
        }
    }
    int idx;
    switch (p.face) {  // as pack_border_slabs_kernel: +-Y [z][x], +-Z [y][x], +-X [z][y]
        case 0: idx = b + 16 * 15 + 256 * a; break;
        case 1: idx = b + 256 * a; break;
        case 2: idx = b + 16 * a; break;
        case 3: idx = b + 16 * a + 256 * 15; break;
        case 4: idx = 16 * b + 256 * a; break;
        default: idx = 15 + 16 * b + 256 * a; break;
    }
    uint16_t* out = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(p.peer) + kSeamInboxHeader) + static_cast<size_t>(p.epoch & 1u) * p.n * 256;
    for (uint32_t i = blockIdx.x; i < p.n; i += gridDim.x)
        out[static_cast<size_t>(i) * 256 + t] = p.voxels[static_cast<size_t>(p.slots[i]) * kChunkVoxels + idx];
    __threadfence_system();  // this thread's stores are performed at the peer before the CTA counts itself out
    __syncthreads();
    if (t == 0) {
        if (atomicAdd(p.done, 1u) == gridDim.x - 1u) {
            __threadfence_system();
            st_release_sys_u32(&p.peer->flag, p.epoch);
            *p.done = 0u;
        }
    }
}
// one 64-bit word device -> pinned host memory by a store over PCIe (UVA), not by a copy engine: a small copy would queue
// behind the bulk uploads that keep the DMA engines busy during a pipelined call
__global__ void publish_u64_kernel(const unsigned long long* src, unsigned long long* host_dst) {
    *host_dst = *src;
    __threadfence_system();
}
// A failing call lets a mesh launch that waits for the rest of its batch run out: every piece is declared released.
__global__ void release_all_pieces_kernel(uint32_t* released) { st_release_gpu_u32(released + threadIdx.x, 1u); }
// dsp_mesh_host_chunks_to_host, device-driven: the arena bytes [*lo_p, *hi_p) of a sub-batch -- bounds that only the device knows
// when the launch is enqueued (they are the cursor after the previous and after this sub-batch) -- stored straight into
// pinned host memory over PCIe (UVA), 16 bytes per thread and iteration.  No host round trip sits between the sub-batches.
__global__ void __launch_bounds__(256) copy_window_to_host_kernel(const uint8_t* arena, uint8_t* host, const unsigned long long* lo_p, const unsigned long long* hi_p,
                                                                  unsigned long long cap) {
    const unsigned long long lo = lo_p ? *lo_p : 0ull, hi = min(*hi_p, cap);
    if (hi <= lo) return;
    const uint4* src = reinterpret_cast<const uint4*>(arena + lo);
    uint4* dst = reinterpret_cast<uint4*>(host + lo);
    const unsigned long long n16 = (hi - lo) >> 4;  // cursors are multiples of 128
    for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n16; i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
        dst[i] = __ldcs(src + i);
}
__global__ void scatter_u32_kernel(uint32_t* dst, const uint2* idx_val, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx_val[i].x] = idx_val[i].y;
}

// chunk.rs:58-65: a chunk's voxels are indices into its own palette; the device store holds global
// ids, so after the raw host->device copy each voxel is replaced by palette_ids[off[i] + voxel].
__global__ void __launch_bounds__(256) palette_remap_kernel(uint16_t* voxels, const uint32_t* slot_list, const uint16_t* palette_ids,
                                                           const uint32_t* palette_offsets, uint32_t n) {
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t slot = slot_list[i], off = palette_offsets[i], len = palette_offsets[i + 1] - off;
        uint4* row = reinterpret_cast<uint4*>(voxels + static_cast<size_t>(slot) * kChunkVoxels) + threadIdx.x * 2;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            uint4 q = row[hlf];
            uint32_t* wq = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t lo = wq[k] & 0xFFFFu, hi = wq[k] >> 16;
                const uint32_t a = lo < len ? palette_ids[off + lo] : 0u;  // out-of-range palette index -> air
                const uint32_t b = hi < len ? palette_ids[off + hi] : 0u;
                wq[k] = a | (b << 16);
            }
            row[hlf] = q;
        }
    }
}

// Chunk::set_block on resident chunks (chunk.rs:49-68).  The host has already resolved world
// coordinates to (slot, voxel index) and kept only the last edit per voxel, so the scatter is
// race-free.  packed = slot (u32), idx | block << 16 (u32).
__global__ void edit_scatter_kernel(uint16_t* voxels, uint8_t* summary, uint8_t* bvalid, const uint2* edits, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 e = edits[i];
    voxels[static_cast<size_t>(e.x) * kChunkVoxels + (e.y & 0xFFFu)] = static_cast<uint16_t>(e.y >> 16);
    summary[e.x] = kSummaryUnknown;
    bvalid[e.x] = 0;
}

// ---- device-side edit pipeline for worlds made of box regions (dsp_apply_edits / dsp_apply_edits_and_remesh) ------------
// World::get_block_world addressing (world.rs:50-74): chunk = div_euclid(w, 16), local = rem_euclid(w, 16); Chunk::set_block
// (chunk.rs:49-68) on the voxel.  The edits of one call apply IN ORDER, so when several hit the same voxel the last
// one must win: pass 1 files every edit under its voxel key in an open-addressed table keeping the highest edit index
// (one atomicMax per edit), pass 2 lets only that edit write, and marks the chunk in a bitmap; pass 3 turns the
// bitmap into the ascending list of dirty slots that the mesh kernel then consumes straight from device memory.
struct EditRegion {
    int32_t ox, oy, oz;
    uint32_t dx, dy, dz;
    uint32_t first_slot;
};
constexpr int kMaxEditRegions = 8;
struct EditParams {
    const int4* edits;        // wx, wy, wz, block
    uint32_t n;
    EditRegion regions[kMaxEditRegions];
    int n_regions;
    unsigned long long* table;  // (key + 1) << 20 | edit index; zeroed before the call
    uint32_t table_mask;
    uint16_t* voxels;
    uint32_t* dirty_bitmap;   // one bit per slot; zeroed before the call
    const uint32_t* nbr;      // halo mode: [slot][6] neighbour table; an edit on a chunk border also dirties the resident neighbour
    uint8_t* summary;         // edited chunks become kSummaryUnknown
    uint8_t* bvalid;          // ... and their border masks stale
};

__device__ __forceinline__ bool edit_key(const EditParams& p, const int4 e, uint64_t* key) {
    const int32_t cx = e.x >> 4, cy = e.y >> 4, cz = e.z >> 4;
    for (int r = 0; r < p.n_regions; ++r) {
        const EditRegion& g = p.regions[r];
        const int64_t a = static_cast<int64_t>(cx) - g.ox, b = static_cast<int64_t>(cy) - g.oy, c = static_cast<int64_t>(cz) - g.oz;
        if (a < 0 || b < 0 || c < 0 || a >= g.dx || b >= g.dy || c >= g.dz) continue;
        const uint64_t slot = g.first_slot + static_cast<uint64_t>(a) + g.dx * (static_cast<uint64_t>(b) + static_cast<uint64_t>(g.dy) * c);
        *key = (slot << 12) | static_cast<uint32_t>((e.x & 15) + 16 * (e.y & 15) + 256 * (e.z & 15));
        return true;
    }
    return false;  // not resident: ignored, like an edit outside the loaded world
}
__device__ __forceinline__ uint32_t edit_hash(uint64_t key) {
    key *= 0x9E3779B97F4A7C15ull;
    return static_cast<uint32_t>(key >> 32);
}

__global__ void edit_file_kernel(const EditParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    uint64_t key;
    if (!edit_key(p, p.edits[i], &key)) return;
    const unsigned long long tag = (key + 1ull) << 20, mine = tag | i;
    for (uint32_t h = edit_hash(key) & p.table_mask;; h = (h + 1) & p.table_mask) {
        unsigned long long cur = p.table[h];
        if (cur == 0ull) {
            cur = atomicCAS(&p.table[h], 0ull, mine);
            if (cur == 0ull) return;
        }
        if ((cur >> 20) == (tag >> 20)) { atomicMax(&p.table[h], mine); return; }
    }
}

__global__ void edit_apply_kernel(const EditParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int4 e = p.edits[i];
    uint64_t key;
    if (!edit_key(p, e, &key)) return;
    const unsigned long long tag = (key + 1ull) << 20;
    for (uint32_t h = edit_hash(key) & p.table_mask;; h = (h + 1) & p.table_mask) {
        const unsigned long long cur = p.table[h];
        if ((cur >> 20) != (tag >> 20)) continue;  // filed by pass 1, so the probe always ends
        if ((cur & 0xFFFFFull) == i) {
            p.voxels[key] = static_cast<uint16_t>(e.w);  // key == slot * 4096 + voxel index
            const uint32_t slot = static_cast<uint32_t>(key >> 12);
            atomicOr(&p.dirty_bitmap[slot >> 5], 1u << (slot & 31u));
            p.summary[slot] = kSummaryUnknown;
            p.bvalid[slot] = 0;
            if (p.nbr != nullptr) {  // cross-chunk culling: the neighbour's border faces depend on this voxel
                const int lx = e.x & 15, ly = e.y & 15, lz = e.z & 15;
                const int faces[3] = {ly == 15 ? 0 : (ly == 0 ? 1 : -1), lz == 0 ? 2 : (lz == 15 ? 3 : -1), lx == 0 ? 4 : (lx == 15 ? 5 : -1)};
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    if (faces[a] < 0) continue;
                    const uint32_t s = p.nbr[static_cast<size_t>(slot) * 6 + faces[a]];
                    if (s != 0xFFFFFFFFu && !(s & 0x80000000u)) atomicOr(&p.dirty_bitmap[s >> 5], 1u << (s & 31u));
                }
            }
        }
        return;
    }
}

// One CTA of 1,024 threads: ascending list of the set bits of bitmap[0 .. n_words).
__global__ void __launch_bounds__(1024) bitmap_to_list_kernel(const uint32_t* bitmap, uint32_t n_words, uint32_t* out_list, uint32_t* out_count) {
    __shared__ uint32_t s_warp[32];
    const uint32_t t = threadIdx.x, per = (n_words + 1023u) / 1024u;
    const uint32_t w0 = min(n_words, t * per), w1 = min(n_words, w0 + per);
    uint32_t cnt = 0;
    for (uint32_t w = w0; w < w1; ++w) cnt += __popc(bitmap[w]);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((t & 31u) >= static_cast<uint32_t>(o)) incl += v;
    }
    if ((t & 31u) == 31u) s_warp[t >> 5] = incl;
    __syncthreads();
    uint32_t pre = 0, total = 0;
    for (uint32_t w = 0; w < 32; ++w) { const uint32_t v = s_warp[w]; pre += w < (t >> 5) ? v : 0u; total += v; }
    uint32_t o = pre + incl - cnt;
    for (uint32_t w = w0; w < w1; ++w) {
        uint32_t bits = bitmap[w];
        while (bits) { out_list[o++] = w * 32u + (__ffs(bits) - 1); bits &= bits - 1; }
    }
    if (t == 0) *out_count = total;
}

// Border layer of each listed chunk at `face` (0 top y=15, 1 bottom y=0, 2 rear z=0, 3 front z=15, 4 right x=0, 5 left
// x=15) as a 256-entry slab: +-Y faces [z][x], +-Z faces [y][x], +-X faces [z][y].  This is what crosses NVLink at a
// region seam in halo mode: 512 bytes per chunk face instead of the 8 KiB chunk.
__global__ void __launch_bounds__(256) pack_border_slabs_kernel(const uint16_t* voxels, const uint32_t* slots, uint32_t n, int face, uint16_t* out) {
    const int t = threadIdx.x, a = t >> 4, b = t & 15;
    int idx;
    switch (face) {
        case 0: idx = b + 16 * 15 + 256 * a; break;
        case 1: idx = b + 256 * a; break;
        case 2: idx = b + 16 * a; break;
        case 3: idx = b + 16 * a + 256 * 15; break;
        case 4: idx = 16 * b + 256 * a; break;
        default: idx = 15 + 16 * b + 256 * a; break;
    }
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x)
        out[static_cast<size_t>(i) * 256 + t] = voxels[static_cast<size_t>(slots[i]) * kChunkVoxels + idx];
}

// map_uv of chunk_mesh.rs:23-28 evaluated once per block id instead of once per vertex:
// out[id] = (map(0).x, map(0).y, map(1).x, map(1).y) with map(o) = uv_min + o * (uv_max - uv_min),
// each operation rounded separately (no FMA contraction) like the rustc-compiled reference.
__global__ void uv_map_kernel(const float4* uv_rows, float4* uvmap, uint32_t n_blocks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_blocks) return;
    const float4 a = i < n_blocks ? uv_rows[i] : make_float4(0.0f, 0.0f, 1.0f, 1.0f);  // row n_blocks = fallback (chunk_mesh.rs:53-56)
    const float dx = __fsub_rn(a.z, a.x), dy = __fsub_rn(a.w, a.y);
    uvmap[i] = make_float4(__fadd_rn(a.x, __fmul_rn(0.0f, dx)), __fadd_rn(a.y, __fmul_rn(0.0f, dy)),
                           __fadd_rn(a.x, __fmul_rn(1.0f, dx)), __fadd_rn(a.y, __fmul_rn(1.0f, dy)));
}

}  // namespace dsp
