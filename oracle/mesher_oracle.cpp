// TEST INFRASTRUCTURE ONLY -- CPU oracle for the chunk terrain -> cull -> mesh path.
//
// This file restates, statement by statement, the algorithm of the reference engine's
// hot path in plain C++17 so that the CUDA path can be checked bit-for-bit.  It is NOT
// part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load it.  The product library (libdespawn_b200.so) never
// links, loads or calls anything in oracle/.
//
// PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
// section 4 / 8c) and its Rust sources cannot be compiled in this image (no rustc/cargo), so there is
// no oracle/_ref.  The pin is the reference's SOURCE TEXT: oracle/reference_source.py parses cube.rs,
// chunk_mesh.rs, chunk.rs, world.rs and math.rs under /root/reference at test time and meshes chunks
// from the parsed tables / conditions / order (sharing nothing with this file);
// tests/test_oracle_vs_reference_source.py holds this oracle to it byte for byte, and the committed
// fixtures tests/golden/refsrc_* + flat_chunk_origin.bin are generated from that parser
// (tests/golden/make_golden_from_reference.py).  Further pins: the hand-derived known answers
// (tests/test_oracle_known_answers.py) and an independent numpy restatement (oracle/numpy_oracle.py).
//
// Reference anchors (paths relative to /root/reference):
//   src/content/world/chunks/chunk.rs:7-12,15-68,86-120   Chunk storage, palette, generators
//   src/content/world/world.rs:28-47                      terrain selection by chunk Y
//   src/content/world/chunks/chunk_mesh.rs:13-129         build_chunk_mesh
//   src/engine/rendering/cube.rs:156-312                  six face templates
//   src/engine/rendering/vertex.rs:5-12                   BlockVertex (20 bytes)
//   src/utils/math.rs:135-170                             Vec3 from/add/mul
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (see oracle/Makefile).  No fast-math:
// every f32 operation below must round exactly like the rustc-compiled reference (IEEE-754
// binary32, one rounding per operation, no FMA contraction).

#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <atomic>
#include <algorithm>
#include <chrono>

namespace {

// ---- chunk.rs:7-12 -------------------------------------------------------------------
constexpr size_t CHUNK_SIZE = 16;
constexpr size_t MAX_CHUNK_INDEX = 4095;
const char* const AIR_BLOCK_ID = "base:air";

// ---- vertex.rs:5-12, math.rs:17-19 ---------------------------------------------------
struct Vec3 {
    float v[3];
};
struct BlockVertex {  // #[repr(C)], 20 bytes, no padding
    Vec3 position;
    float tex_coords[2];
};
static_assert(sizeof(BlockVertex) == 20, "BlockVertex must be 20 bytes");

// math.rs:135-139  From<[i32;3]>
inline Vec3 vec3_from_i32(const int32_t p[3]) {
    return Vec3{{static_cast<float>(p[0]), static_cast<float>(p[1]), static_cast<float>(p[2])}};
}
// math.rs:155-164  Add
inline Vec3 vec3_add(Vec3 a, Vec3 b) { return Vec3{{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
// math.rs:166-170  Mul<f32>
inline Vec3 vec3_mul(Vec3 a, float s) { return Vec3{{a.v[0] * s, a.v[1] * s, a.v[2] * s}}; }

// ---- texture_atlas.rs:15-19 ----------------------------------------------------------
struct AtlasUV {
    float uv_min[2];
    float uv_max[2];
};

// ---- cube.rs:156-312 : six templates, each [c0,c1,c2,c2,c3,c0] -----------------------
// Emission order in build_chunk_mesh is TOP, BOTTOM, REAR, FRONT, RIGHT, LEFT
// (chunk_mesh.rs:65-106); the table below is stored in that order.
#define V(px, py, pz, u, w) BlockVertex{Vec3{{px, py, pz}}, {u, w}}
const BlockVertex FACE_TEMPLATES[6][6] = {
    // TOP_FACE  cube.rs:287-312
    {V(-0.5f, 0.5f, -0.5f, 1.0f, 1.0f), V(0.5f, 0.5f, -0.5f, 0.0f, 1.0f), V(0.5f, 0.5f, 0.5f, 0.0f, 0.0f),
     V(0.5f, 0.5f, 0.5f, 0.0f, 0.0f), V(-0.5f, 0.5f, 0.5f, 1.0f, 0.0f), V(-0.5f, 0.5f, -0.5f, 1.0f, 1.0f)},
    // BOTTOM_FACE  cube.rs:261-286
    {V(-0.5f, -0.5f, 0.5f, 1.0f, 1.0f), V(0.5f, -0.5f, 0.5f, 0.0f, 1.0f), V(0.5f, -0.5f, -0.5f, 0.0f, 0.0f),
     V(0.5f, -0.5f, -0.5f, 0.0f, 0.0f), V(-0.5f, -0.5f, -0.5f, 1.0f, 0.0f), V(-0.5f, -0.5f, 0.5f, 1.0f, 1.0f)},
    // REAR_FACE  cube.rs:183-208
    {V(-0.5f, -0.5f, -0.5f, 0.0f, 1.0f), V(0.5f, -0.5f, -0.5f, 1.0f, 1.0f), V(0.5f, 0.5f, -0.5f, 1.0f, 0.0f),
     V(0.5f, 0.5f, -0.5f, 1.0f, 0.0f), V(-0.5f, 0.5f, -0.5f, 0.0f, 0.0f), V(-0.5f, -0.5f, -0.5f, 0.0f, 1.0f)},
    // FRONT_FACE  cube.rs:156-181
    {V(-0.5f, -0.5f, 0.5f, 1.0f, 1.0f), V(-0.5f, 0.5f, 0.5f, 1.0f, 0.0f), V(0.5f, 0.5f, 0.5f, 0.0f, 0.0f),
     V(0.5f, 0.5f, 0.5f, 0.0f, 0.0f), V(0.5f, -0.5f, 0.5f, 0.0f, 1.0f), V(-0.5f, -0.5f, 0.5f, 1.0f, 1.0f)},
    // RIGHT_FACE (-X)  cube.rs:209-234
    {V(-0.5f, -0.5f, -0.5f, 0.0f, 1.0f), V(-0.5f, 0.5f, -0.5f, 0.0f, 0.0f), V(-0.5f, 0.5f, 0.5f, 1.0f, 0.0f),
     V(-0.5f, 0.5f, 0.5f, 1.0f, 0.0f), V(-0.5f, -0.5f, 0.5f, 1.0f, 1.0f), V(-0.5f, -0.5f, -0.5f, 0.0f, 1.0f)},
    // LEFT_FACE (+X)  cube.rs:235-260
    {V(0.5f, -0.5f, -0.5f, 1.0f, 1.0f), V(0.5f, -0.5f, 0.5f, 0.0f, 1.0f), V(0.5f, 0.5f, 0.5f, 0.0f, 0.0f),
     V(0.5f, 0.5f, 0.5f, 0.0f, 0.0f), V(0.5f, 0.5f, -0.5f, 1.0f, 0.0f), V(0.5f, -0.5f, -0.5f, 1.0f, 1.0f)},
};
#undef V

// ---- chunk.rs:15-68 ------------------------------------------------------------------
struct Chunk {
    int32_t position[3];
    std::vector<uint16_t> blocks;                       // per-chunk palette indices
    std::vector<std::string> palette;                   // index -> block id
    std::unordered_map<std::string, uint16_t> palette_map;  // block id -> index

    explicit Chunk(const int32_t pos[3]) {              // chunk.rs:23-37
        position[0] = pos[0]; position[1] = pos[1]; position[2] = pos[2];
        palette.push_back(AIR_BLOCK_ID);
        palette_map.emplace(AIR_BLOCK_ID, 0);
        blocks.assign(CHUNK_SIZE * CHUNK_SIZE * CHUNK_SIZE, 0);
    }
    bool is_air_at_idx(size_t idx) const { return blocks[idx] == 0; }          // chunk.rs:39-41
    static size_t index(size_t x, size_t y, size_t z) {                        // chunk.rs:44-46
        return x + y * CHUNK_SIZE + z * CHUNK_SIZE * CHUNK_SIZE;
    }
    void set_block(size_t x, size_t y, size_t z, const std::string& block_id) {  // chunk.rs:49-68
        const size_t idx = index(x, y, z);
        if (block_id == AIR_BLOCK_ID) {
            blocks[idx] = 0;
            return;
        }
        uint16_t palette_idx;
        auto it = palette_map.find(block_id);
        if (it != palette_map.end()) {
            palette_idx = it->second;
        } else {
            palette_idx = static_cast<uint16_t>(palette.size());
            palette.push_back(block_id);
            palette_map.emplace(block_id, palette_idx);
        }
        blocks[idx] = palette_idx;
    }
    void generate_flat(const std::string& dirt_id) {    // chunk.rs:86-100
        for (size_t x = 0; x < CHUNK_SIZE; ++x)
            for (size_t y = 0; y < CHUNK_SIZE; ++y)
                for (size_t z = 0; z < CHUNK_SIZE; ++z)
                    set_block(x, y, z, y < CHUNK_SIZE / 2 ? dirt_id : std::string(AIR_BLOCK_ID));
    }
    void generate_full(const std::string& dirt_id) {    // chunk.rs:102-110
        for (size_t x = 0; x < CHUNK_SIZE; ++x)
            for (size_t y = 0; y < CHUNK_SIZE; ++y)
                for (size_t z = 0; z < CHUNK_SIZE; ++z) set_block(x, y, z, dirt_id);
    }
    void generate_empty() {                             // chunk.rs:112-120
        for (size_t x = 0; x < CHUNK_SIZE; ++x)
            for (size_t y = 0; y < CHUNK_SIZE; ++y)
                for (size_t z = 0; z < CHUNK_SIZE; ++z) set_block(x, y, z, AIR_BLOCK_ID);
    }
};

// ---- world.rs:28-47 : terrain kind is a function of chunk Y only ----------------------
Chunk world_get_chunk(const int32_t pos[3]) {
    Chunk chunk(pos);
    if (pos[1] > 0) {
        chunk.generate_empty();
    } else if (pos[1] == 0) {
        chunk.generate_flat("template:dirt");
    } else {
        chunk.generate_full("template:dirt");
    }
    return chunk;
}

using BlockUvMap = std::unordered_map<std::string, AtlasUV>;  // RapidHashMap<String,AtlasUV>

// ---- chunk_mesh.rs:13-129 -------------------------------------------------------------
// `faithful` keeps the reference's per-voxel string-keyed lookup and un-reserved vector;
// the returned bool is false where the reference returns None.
bool build_chunk_mesh(const Chunk& chunk, const BlockUvMap& block_uvs, std::vector<BlockVertex>& vertices) {
    vertices.clear();
    if (chunk.palette.size() == 1) return false;  // :18-20

    auto map_uv = [](const float orig[2], const AtlasUV& atlas, float out[2]) {  // :23-28
        out[0] = atlas.uv_min[0] + orig[0] * (atlas.uv_max[0] - atlas.uv_min[0]);
        out[1] = atlas.uv_min[1] + orig[1] * (atlas.uv_max[1] - atlas.uv_min[1]);
    };
    auto is_air = [&](size_t idx) { return chunk.is_air_at_idx(idx); };  // :30

    for (size_t idx = 0; idx < MAX_CHUNK_INDEX + 1; ++idx) {  // :32
        const size_t palette_idx = chunk.blocks[idx];
        const std::string& block_id = chunk.palette[palette_idx];
        if (chunk.is_air_at_idx(idx)) continue;  // :36-38

        const size_t x = idx % CHUNK_SIZE;  // :41-43
        const size_t y = (idx / CHUNK_SIZE) % CHUNK_SIZE;
        const size_t z = idx / (CHUNK_SIZE * CHUNK_SIZE);

        // :46-51 -- chunk borders always count as "air"; RIGHT is -X, LEFT is +X
        const bool face_air[6] = {
            y == CHUNK_SIZE - 1 || is_air(idx + CHUNK_SIZE),               // top
            y == 0 || is_air(idx - CHUNK_SIZE),                            // bottom
            z == 0 || is_air(idx - CHUNK_SIZE * CHUNK_SIZE),               // rear
            z == CHUNK_SIZE - 1 || is_air(idx + CHUNK_SIZE * CHUNK_SIZE),  // front
            x == 0 || is_air(idx - 1),                                     // right
            x == CHUNK_SIZE - 1 || is_air(idx + 1),                        // left
        };

        AtlasUV atlas{{0.0f, 0.0f}, {1.0f, 1.0f}};  // :53-56 fallback
        auto it = block_uvs.find(block_id);
        if (it != block_uvs.end()) atlas = it->second;

        // :61-62  Vec3::from([x,y,z] as f32) + (Vec3::from(position) * 16.0)
        const Vec3 local{{static_cast<float>(x), static_cast<float>(y), static_cast<float>(z)}};
        const Vec3 pos_offset = vec3_add(local, vec3_mul(vec3_from_i32(chunk.position), static_cast<float>(CHUNK_SIZE)));

        for (int f = 0; f < 6; ++f) {  // :65-106, fixed order
            if (!face_air[f]) continue;
            for (int k = 0; k < 6; ++k) {
                BlockVertex v = FACE_TEMPLATES[f][k];
                v.position = vec3_add(v.position, pos_offset);  // v.position += pos_offset
                float uv[2];
                map_uv(v.tex_coords, atlas, uv);
                v.tex_coords[0] = uv[0];
                v.tex_coords[1] = uv[1];
                vertices.push_back(v);
            }
        }
    }
    return !vertices.empty();  // :109-128
}

// "fair" flavour for the CPU baseline: same arithmetic, but the palette->UV lookup is
// resolved once per palette entry and the output vector is reserved.  Must produce the
// same bytes as build_chunk_mesh (tested).
bool build_chunk_mesh_fair(const Chunk& chunk, const BlockUvMap& block_uvs, std::vector<BlockVertex>& vertices) {
    vertices.clear();
    if (chunk.palette.size() == 1) return false;
    std::vector<AtlasUV> pal_uv(chunk.palette.size(), AtlasUV{{0.0f, 0.0f}, {1.0f, 1.0f}});
    for (size_t p = 0; p < chunk.palette.size(); ++p) {
        auto it = block_uvs.find(chunk.palette[p]);
        if (it != block_uvs.end()) pal_uv[p] = it->second;
    }
    vertices.reserve(4096);
    const float cx = static_cast<float>(chunk.position[0]) * 16.0f;
    const float cy = static_cast<float>(chunk.position[1]) * 16.0f;
    const float cz = static_cast<float>(chunk.position[2]) * 16.0f;
    const uint16_t* b = chunk.blocks.data();
    for (size_t idx = 0; idx < 4096; ++idx) {
        if (b[idx] == 0) continue;
        const size_t x = idx & 15, y = (idx >> 4) & 15, z = idx >> 8;
        const bool face_air[6] = {y == 15 || b[idx + 16] == 0,  y == 0 || b[idx - 16] == 0,
                                  z == 0 || b[idx - 256] == 0,  z == 15 || b[idx + 256] == 0,
                                  x == 0 || b[idx - 1] == 0,    x == 15 || b[idx + 1] == 0};
        const AtlasUV& a = pal_uv[b[idx]];
        const float off[3] = {static_cast<float>(x) + cx, static_cast<float>(y) + cy, static_cast<float>(z) + cz};
        for (int f = 0; f < 6; ++f) {
            if (!face_air[f]) continue;
            for (int k = 0; k < 6; ++k) {
                const BlockVertex& t = FACE_TEMPLATES[f][k];
                BlockVertex v;
                v.position.v[0] = t.position.v[0] + off[0];
                v.position.v[1] = t.position.v[1] + off[1];
                v.position.v[2] = t.position.v[2] + off[2];
                v.tex_coords[0] = a.uv_min[0] + t.tex_coords[0] * (a.uv_max[0] - a.uv_min[0]);
                v.tex_coords[1] = a.uv_min[1] + t.tex_coords[1] * (a.uv_max[1] - a.uv_min[1]);
                vertices.push_back(v);
            }
        }
    }
    return !vertices.empty();
}

// ---- registry used by the C interface -------------------------------------------------
// The device path stores *global* block ids (0 = air).  The oracle goes the long way round:
// global id -> block-id string -> per-chunk palette (first-seen order, chunk.rs:58-65) ->
// string-keyed UV lookup, exactly as the reference would.
std::string block_name(uint32_t id) {
    if (id == 0) return AIR_BLOCK_ID;
    if (id == 1) return "template:dirt";      // assets/data/blocks/dirt.json5
    if (id == 2) return "template:engine";    // assets/data/blocks/engine.json5
    return "oracle:block_" + std::to_string(id);
}

Chunk chunk_from_global_ids(const int32_t pos[3], const uint16_t* ids) {
    Chunk c(pos);
    for (size_t x = 0; x < CHUNK_SIZE; ++x)
        for (size_t y = 0; y < CHUNK_SIZE; ++y)
            for (size_t z = 0; z < CHUNK_SIZE; ++z) c.set_block(x, y, z, block_name(ids[Chunk::index(x, y, z)]));
    return c;
}

BlockUvMap uv_map_from_table(const float* uv, uint32_t n_blocks) {
    BlockUvMap m;
    for (uint32_t id = 1; id < n_blocks; ++id) {  // id 0 is air; ids >= n_blocks are "missing"
        AtlasUV a{{uv[4 * id + 0], uv[4 * id + 1]}, {uv[4 * id + 2], uv[4 * id + 3]}};
        m.emplace(block_name(id), a);
    }
    return m;
}

// ---- builder-defined synthetic terrain (NOT in the reference; see DESIGN.md "Terrain spec") --------
// All-integer so the CUDA generator can be bit-exact.  Kinds mirror include/despawn_b200.h.
enum GenKind { GEN_REFERENCE = 0, GEN_NOISE = 1, GEN_RANDOM = 2, GEN_CHECKER = 3 };

inline uint32_t mix32(uint32_t h) {  // "lowbias32" integer finaliser
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return h;
}
inline uint32_t hash2(uint32_t seed, int32_t a, int32_t b) {
    return mix32(seed ^ mix32(static_cast<uint32_t>(a) * 0x9E3779B1U ^ mix32(static_cast<uint32_t>(b) * 0x85EBCA77U + 0xC2B2AE3DU)));
}
inline uint32_t hash3(uint32_t seed, int32_t a, int32_t b, int32_t c) {
    return mix32(hash2(seed, a, b) + static_cast<uint32_t>(c) * 0x27D4EB2FU);
}
// One octave of 2-D value noise, cell size 32>>o voxels, result in [0,255].
inline uint32_t value_noise_octave(uint32_t seed, int32_t wx, int32_t wz, int o) {
    const int s = 5 - o;                       // log2(cell size)
    const int32_t ix = wx >> s, iz = wz >> s;  // floor division (arithmetic shift)
    const uint32_t tx = static_cast<uint32_t>(wx & ((1 << s) - 1)) << (8 - s);  // [0,256)
    const uint32_t tz = static_cast<uint32_t>(wz & ((1 << s) - 1)) << (8 - s);
    const uint32_t sx = (tx * tx * (768U - 2U * tx)) >> 16;  // smoothstep, [0,256)
    const uint32_t sz = (tz * tz * (768U - 2U * tz)) >> 16;
    const uint32_t v00 = hash2(seed + static_cast<uint32_t>(o), ix, iz) & 0xFFU;
    const uint32_t v10 = hash2(seed + static_cast<uint32_t>(o), ix + 1, iz) & 0xFFU;
    const uint32_t v01 = hash2(seed + static_cast<uint32_t>(o), ix, iz + 1) & 0xFFU;
    const uint32_t v11 = hash2(seed + static_cast<uint32_t>(o), ix + 1, iz + 1) & 0xFFU;
    const uint32_t a = v00 * (256U - sx) + v10 * sx;
    const uint32_t b = v01 * (256U - sx) + v11 * sx;
    return (a * (256U - sz) + b * sz) >> 16;
}
inline int32_t terrain_height(uint32_t seed, int32_t wx, int32_t wz) {  // [16, 239]
    const uint32_t n = 4U * value_noise_octave(seed, wx, wz, 0) + 2U * value_noise_octave(seed, wx, wz, 1) +
                       value_noise_octave(seed, wx, wz, 2);
    return 16 + static_cast<int32_t>(n >> 3);
}

void generate_ids(int kind, uint32_t seed, const int32_t pos[3], uint16_t* ids) {
    if (kind == GEN_REFERENCE) {
        // Through the faithful Chunk/World path, then palette -> global ids.
        Chunk c = world_get_chunk(pos);
        for (size_t i = 0; i < 4096; ++i) {
            const std::string& name = c.palette[c.blocks[i]];
            ids[i] = name == AIR_BLOCK_ID ? 0 : (name == "template:dirt" ? 1 : 2);
        }
        return;
    }
    for (int z = 0; z < 16; ++z)
        for (int y = 0; y < 16; ++y)
            for (int x = 0; x < 16; ++x) {
                const int32_t wx = pos[0] * 16 + x, wy = pos[1] * 16 + y, wz = pos[2] * 16 + z;
                uint16_t id = 0;
                if (kind == GEN_NOISE) {
                    const int32_t h = terrain_height(seed, wx, wz);
                    id = wy < h - 1 ? 1 : (wy == h - 1 ? 2 : 0);
                } else if (kind == GEN_RANDOM) {
                    const uint32_t r = hash3(seed, wx, wy, wz);
                    id = (r & 1U) ? static_cast<uint16_t>(1U + ((r >> 1) & 1U)) : 0;
                } else if (kind == GEN_CHECKER) {
                    id = ((wx + wy + wz) & 1) == 0 ? 1 : 0;
                }
                ids[Chunk::index(x, y, z)] = id;
            }
}

// ---- extension oracle: cross-chunk (halo) culling ------------------------------------------------
// Same as build_chunk_mesh except that a border face is tested against the adjacent voxel of the
// neighbouring chunk (world.rs:50-74 style world addressing) instead of always being emitted.
// `nbr[f]` is the 16x16 border slab of the neighbour across face f (global ids), or nullptr when that
// neighbour is not resident (=> treated as air, i.e. the reference behaviour).  Slab layout: for +-Y
// faces [z][x], for +-Z faces [y][x], for +-X faces [z][y].
bool build_chunk_mesh_halo(const Chunk& chunk, const BlockUvMap& block_uvs, const uint16_t* const nbr[6],
                           std::vector<BlockVertex>& vertices) {
    vertices.clear();
    if (chunk.palette.size() == 1) return false;
    for (size_t idx = 0; idx < 4096; ++idx) {
        if (chunk.is_air_at_idx(idx)) continue;
        const std::string& block_id = chunk.palette[chunk.blocks[idx]];
        const size_t x = idx % 16, y = (idx / 16) % 16, z = idx / 256;
        auto slab_air = [&](int f, size_t a, size_t b) { return nbr[f] == nullptr || nbr[f][a * 16 + b] == 0; };
        const bool face_air[6] = {
            y == 15 ? slab_air(0, z, x) : chunk.is_air_at_idx(idx + 16),
            y == 0 ? slab_air(1, z, x) : chunk.is_air_at_idx(idx - 16),
            z == 0 ? slab_air(2, y, x) : chunk.is_air_at_idx(idx - 256),
            z == 15 ? slab_air(3, y, x) : chunk.is_air_at_idx(idx + 256),
            x == 0 ? slab_air(4, z, y) : chunk.is_air_at_idx(idx - 1),
            x == 15 ? slab_air(5, z, y) : chunk.is_air_at_idx(idx + 1),
        };
        AtlasUV atlas{{0.0f, 0.0f}, {1.0f, 1.0f}};
        auto it = block_uvs.find(block_id);
        if (it != block_uvs.end()) atlas = it->second;
        const Vec3 local{{static_cast<float>(x), static_cast<float>(y), static_cast<float>(z)}};
        const Vec3 pos_offset = vec3_add(local, vec3_mul(vec3_from_i32(chunk.position), 16.0f));
        for (int f = 0; f < 6; ++f) {
            if (!face_air[f]) continue;
            for (int k = 0; k < 6; ++k) {
                BlockVertex v = FACE_TEMPLATES[f][k];
                v.position = vec3_add(v.position, pos_offset);
                const float u = atlas.uv_min[0] + v.tex_coords[0] * (atlas.uv_max[0] - atlas.uv_min[0]);
                const float w = atlas.uv_min[1] + v.tex_coords[1] * (atlas.uv_max[1] - atlas.uv_min[1]);
                v.tex_coords[0] = u;
                v.tex_coords[1] = w;
                vertices.push_back(v);
            }
        }
    }
    return !vertices.empty();
}


// ---- extension oracle: greedy run-stack merge and indexed layout ----------------
// NOT in the reference (it emits one unit quad per visible face, chunk_mesh.rs:65-106, and draws non-indexed,
// scene_game.rs:118-122).  Builder-defined spec, pinned to the reference by two invariants that the tests check:
//   (1) the merged quads, re-expanded cell by cell, are exactly the reference's visible-face set;
//   (2) every corner of a merged quad is computed with the reference's own arithmetic for the voxel under that
//       corner (chunk_mesh.rs:61-62,67), so an unmerged quad is byte-identical to the reference's face and a
//       chunk in which nothing merges reproduces the reference stream.
// Spec.  For direction d the face plane has a run axis u and a stack axis v:
//       TOP/BOTTOM (normal y): u = x, v = z      REAR/FRONT (normal z): u = x, v = y      RIGHT/LEFT (normal x): u = y, v = z
//   A RUN is a maximal interval of cells along u (fixed slice, fixed v) whose faces are all visible and whose block
//   ids (palette entries, chunk.rs:58-65) are equal.  Two runs in adjacent rows v, v+1 of the same slice are IDENTICAL
//   when they have the same start, length and block id.  A QUAD is a maximal chain of identical runs in consecutive
//   rows: w = run length, h = chain length.  (Unlike the sequential "take the widest, then the tallest" greedy, no
//   decision depends on an earlier one, so every row can decide independently -- that is what the CUDA kernel
//   exploits with warp ballots.)  The quad's ANCHOR is its cell with the smallest voxel index; quads are emitted in
//   the reference's order applied to anchors: anchor voxel index ascending, then TOP, BOTTOM, REAR, FRONT, RIGHT, LEFT.
//   Corner k of the quad = template corner k of the unit face (cube.rs:156-312) placed on the quad's cell that lies
//   furthest in the direction of the corner's sign on each axis; tex_coords are the template's, through map_uv
//   (chunk_mesh.rs:23-28): the block's atlas tile stretches over the merged quad (the reference's fragment shader
//   samples atlas UVs directly, assets/shaders/first_triangle/fragment.glsl:10-12, so tiling would need a shader change).
//   Indexed layout: 4 vertices per quad (template corners c0,c1,c2,c3) and 6 u32 indices 4q + (0,1,2,2,3,0) -- the
//   expansion pattern of every template, cube.rs:156-312 -- local to the chunk.
struct Quad {
    uint16_t idx;  // anchor voxel
    uint8_t dir, w, h;
};
const int FACE_AXES[6][3] = {{1, 0, 2}, {1, 0, 2}, {2, 0, 1}, {2, 0, 1}, {0, 1, 2}, {0, 1, 2}};  // normal, u, v

// vis[d][idx]: face d of voxel idx is emitted.  nbr == nullptr: reference cull (chunk_mesh.rs:46-51).
void face_visibility(const Chunk& chunk, const uint16_t* const* nbr, std::vector<uint8_t>& vis) {
    vis.assign(6 * 4096, 0);
    for (size_t idx = 0; idx < 4096; ++idx) {
        if (chunk.is_air_at_idx(idx)) continue;
        const size_t x = idx % 16, y = (idx / 16) % 16, z = idx / 256;
        auto slab_air = [&](int f, size_t a, size_t b) { return nbr == nullptr || nbr[f] == nullptr || nbr[f][a * 16 + b] == 0; };
        vis[0 * 4096 + idx] = y == 15 ? slab_air(0, z, x) : chunk.is_air_at_idx(idx + 16);
        vis[1 * 4096 + idx] = y == 0 ? slab_air(1, z, x) : chunk.is_air_at_idx(idx - 16);
        vis[2 * 4096 + idx] = z == 0 ? slab_air(2, y, x) : chunk.is_air_at_idx(idx - 256);
        vis[3 * 4096 + idx] = z == 15 ? slab_air(3, y, x) : chunk.is_air_at_idx(idx + 256);
        vis[4 * 4096 + idx] = x == 0 ? slab_air(4, z, y) : chunk.is_air_at_idx(idx - 1);
        vis[5 * 4096 + idx] = x == 15 ? slab_air(5, z, y) : chunk.is_air_at_idx(idx + 1);
    }
}

std::vector<Quad> merge_quads(const Chunk& chunk, const std::vector<uint8_t>& vis, bool greedy) {
    std::vector<Quad> quads;
    for (int d = 0; d < 6; ++d) {
        const int an = FACE_AXES[d][0], au = FACE_AXES[d][1], av = FACE_AXES[d][2];
        auto at = [&](int s, int u, int v) {
            int c[3];
            c[an] = s; c[au] = u; c[av] = v;
            return Chunk::index(static_cast<size_t>(c[0]), static_cast<size_t>(c[1]), static_cast<size_t>(c[2]));
        };
        auto face = [&](int s, int u, int v) { return vis[static_cast<size_t>(d) * 4096 + at(s, u, v)] != 0; };
        auto id = [&](int s, int u, int v) { return chunk.blocks[at(s, u, v)]; };
        // is there a run with exactly this start, length and block id in row v?
        auto has_run = [&](int s, int v, int u0, int w, uint16_t bid) {
            for (int u = u0; u < u0 + w; ++u)
                if (!face(s, u, v) || id(s, u, v) != bid) return false;
            if (u0 > 0 && face(s, u0 - 1, v) && id(s, u0 - 1, v) == bid) return false;            // would start earlier
            if (u0 + w < 16 && face(s, u0 + w, v) && id(s, u0 + w, v) == bid) return false;        // would end later
            return true;
        };
        for (int s = 0; s < 16; ++s)
            for (int v = 0; v < 16; ++v)
                for (int u = 0; u < 16;) {
                    if (!face(s, u, v)) { ++u; continue; }
                    if (!greedy) {
                        quads.push_back(Quad{static_cast<uint16_t>(at(s, u, v)), static_cast<uint8_t>(d), 1, 1});
                        ++u;
                        continue;
                    }
                    const uint16_t bid = id(s, u, v);
                    int w = 1;
                    while (u + w < 16 && face(s, u + w, v) && id(s, u + w, v) == bid) ++w;
                    if (!(v > 0 && has_run(s, v - 1, u, w, bid))) {  // not a continuation: this run opens a quad
                        int h = 1;
                        while (v + h < 16 && has_run(s, v + h, u, w, bid)) ++h;
                        quads.push_back(Quad{static_cast<uint16_t>(at(s, u, v)), static_cast<uint8_t>(d), static_cast<uint8_t>(w), static_cast<uint8_t>(h)});
                    }
                    u += w;
                }
    }
    std::stable_sort(quads.begin(), quads.end(), [](const Quad& a, const Quad& b) { return a.idx != b.idx ? a.idx < b.idx : a.dir < b.dir; });
    return quads;
}

// Vertices (and indices) of the quads.  indexed == false: 6 vertices per quad, the reference's layout.
void emit_quads(const Chunk& chunk, const BlockUvMap& block_uvs, const std::vector<Quad>& quads, bool indexed,
                std::vector<BlockVertex>& vertices, std::vector<uint32_t>& indices) {
    vertices.clear();
    indices.clear();
    static const int kExpand[6] = {0, 1, 2, 2, 3, 0};
    static const int kCornerVertex[4] = {0, 1, 2, 4};  // template vertex that carries corner c0..c3
    uint32_t q = 0;
    for (const Quad& quad : quads) {
        const size_t idx = quad.idx;
        const std::string& block_id = chunk.palette[chunk.blocks[idx]];
        AtlasUV atlas{{0.0f, 0.0f}, {1.0f, 1.0f}};  // chunk_mesh.rs:53-56
        auto it = block_uvs.find(block_id);
        if (it != block_uvs.end()) atlas = it->second;
        const int anchor[3] = {static_cast<int>(idx % 16), static_cast<int>((idx / 16) % 16), static_cast<int>(idx / 256)};
        int ext[3];
        ext[FACE_AXES[quad.dir][0]] = 1;
        ext[FACE_AXES[quad.dir][1]] = quad.w;
        ext[FACE_AXES[quad.dir][2]] = quad.h;
        auto corner = [&](int k) {
            BlockVertex v = FACE_TEMPLATES[quad.dir][k];
            Vec3 cell;
            for (int a = 0; a < 3; ++a) cell.v[a] = static_cast<float>(anchor[a] + (v.position.v[a] > 0.0f ? ext[a] - 1 : 0));
            const Vec3 pos_offset = vec3_add(cell, vec3_mul(vec3_from_i32(chunk.position), 16.0f));  // chunk_mesh.rs:61-62
            v.position = vec3_add(v.position, pos_offset);
            const float u = atlas.uv_min[0] + v.tex_coords[0] * (atlas.uv_max[0] - atlas.uv_min[0]);
            const float w = atlas.uv_min[1] + v.tex_coords[1] * (atlas.uv_max[1] - atlas.uv_min[1]);
            v.tex_coords[0] = u;
            v.tex_coords[1] = w;
            return v;
        };
        if (indexed) {
            for (int c = 0; c < 4; ++c) vertices.push_back(corner(kCornerVertex[c]));
            for (int k = 0; k < 6; ++k) indices.push_back(4 * q + static_cast<uint32_t>(kExpand[k]));
        } else {
            for (int k = 0; k < 6; ++k) vertices.push_back(corner(k));
        }
        ++q;
    }
}

}  // namespace

// =====================================================================================================
// C interface (ctypes).  All buffers caller-owned.
// =====================================================================================================
extern "C" {

// Fill `ids` (4096 x u16, index x + 16y + 256z) with global block ids for chunk `pos`.
void orc_generate(int kind, uint32_t seed, const int32_t* pos, uint16_t* ids) { generate_ids(kind, seed, pos, ids); }

// Terrain height of the builder-defined noise spec (for tests of the generator alone).
int32_t orc_terrain_height(uint32_t seed, int32_t wx, int32_t wz) { return terrain_height(seed, wx, wz); }

// Mesh one chunk.  uv: n_blocks x 4 floats (uv_min.x, uv_min.y, uv_max.x, uv_max.y), entry 0 unused.
// flavour: 0 = faithful, 1 = fair.  Writes up to cap_vertices 20-byte vertices; returns the vertex
// count (0 <=> the reference returns None), or -1 if cap_vertices is too small.
int64_t orc_mesh_chunk(const int32_t* pos, const uint16_t* ids, const float* uv, uint32_t n_blocks, int flavour,
                       void* out_vertices, int64_t cap_vertices) {
    Chunk c = chunk_from_global_ids(pos, ids);
    BlockUvMap m = uv_map_from_table(uv, n_blocks);
    std::vector<BlockVertex> verts;
    const bool some = flavour == 0 ? build_chunk_mesh(c, m, verts) : build_chunk_mesh_fair(c, m, verts);
    if (!some) return 0;
    if (static_cast<int64_t>(verts.size()) > cap_vertices) return -1;
    std::memcpy(out_vertices, verts.data(), verts.size() * sizeof(BlockVertex));
    return static_cast<int64_t>(verts.size());
}

// Halo-mode extension.  nbr_slabs: 6 x 256 u16 (order top,bottom,rear,front,right,left),
// nbr_present: 6 flags.
int64_t orc_mesh_chunk_halo(const int32_t* pos, const uint16_t* ids, const uint16_t* nbr_slabs, const uint8_t* nbr_present,
                            const float* uv, uint32_t n_blocks, void* out_vertices, int64_t cap_vertices) {
    Chunk c = chunk_from_global_ids(pos, ids);
    BlockUvMap m = uv_map_from_table(uv, n_blocks);
    const uint16_t* nbr[6];
    for (int f = 0; f < 6; ++f) nbr[f] = nbr_present[f] ? nbr_slabs + 256 * f : nullptr;
    std::vector<BlockVertex> verts;
    if (!build_chunk_mesh_halo(c, m, nbr, verts)) return 0;
    if (static_cast<int64_t>(verts.size()) > cap_vertices) return -1;
    std::memcpy(out_vertices, verts.data(), verts.size() * sizeof(BlockVertex));
    return static_cast<int64_t>(verts.size());
}

// Extension modes (greedy merge and/or indexed layout, optionally with halo culling).  nbr_slabs / nbr_present may be
// NULL (reference cull).  out_quads (may be NULL, capacity = cap_quads) receives one record per quad:
// anchor | dir << 12 | (w-1) << 15 | (h-1) << 19.  Returns the vertex count (-1: a capacity was too small);
// *out_n_indices and *out_n_quads receive the other counts.
int64_t orc_mesh_chunk_ext(const int32_t* pos, const uint16_t* ids, const uint16_t* nbr_slabs, const uint8_t* nbr_present,
                           const float* uv, uint32_t n_blocks, int greedy, int indexed, void* out_vertices, int64_t cap_vertices,
                           uint32_t* out_indices, int64_t cap_indices, int64_t* out_n_indices, uint32_t* out_quads, int64_t cap_quads,
                           int64_t* out_n_quads) {
    Chunk c = chunk_from_global_ids(pos, ids);
    BlockUvMap m = uv_map_from_table(uv, n_blocks);
    const uint16_t* nbr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (nbr_slabs != nullptr && nbr_present != nullptr)
        for (int f = 0; f < 6; ++f) nbr[f] = nbr_present[f] ? nbr_slabs + 256 * f : nullptr;
    std::vector<uint8_t> vis;
    face_visibility(c, nbr, vis);
    const std::vector<Quad> quads = merge_quads(c, vis, greedy != 0);
    std::vector<BlockVertex> verts;
    std::vector<uint32_t> inds;
    emit_quads(c, m, quads, indexed != 0, verts, inds);
    if (out_n_indices) *out_n_indices = static_cast<int64_t>(inds.size());
    if (out_n_quads) *out_n_quads = static_cast<int64_t>(quads.size());
    if (static_cast<int64_t>(verts.size()) > cap_vertices || static_cast<int64_t>(inds.size()) > cap_indices) return -1;
    if (out_quads != nullptr) {
        if (static_cast<int64_t>(quads.size()) > cap_quads) return -1;
        for (size_t i = 0; i < quads.size(); ++i)
            out_quads[i] = quads[i].idx | (static_cast<uint32_t>(quads[i].dir) << 12) | (static_cast<uint32_t>(quads[i].w - 1) << 15) |
                           (static_cast<uint32_t>(quads[i].h - 1) << 19);
    }
    if (!verts.empty()) std::memcpy(out_vertices, verts.data(), verts.size() * sizeof(BlockVertex));
    if (!inds.empty()) std::memcpy(out_indices, inds.data(), inds.size() * sizeof(uint32_t));
    return static_cast<int64_t>(verts.size());
}

// Mesh a batch of n chunks (ids: n x 4096) on `threads` host threads and return the number of
// vertices per chunk in counts[n]; if out_vertices != NULL, vertices are written tightly packed
// in chunk order (the caller sizes it from a first counting call).  Used for bulk parity checks.
int64_t orc_mesh_batch(int64_t n, const int32_t* pos, const uint16_t* ids, const float* uv, uint32_t n_blocks,
                       int flavour, int threads, int64_t* counts, void* out_vertices, int64_t cap_vertices) {
    BlockUvMap m = uv_map_from_table(uv, n_blocks);
    std::vector<std::vector<BlockVertex>> all(static_cast<size_t>(n));
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            Chunk c = chunk_from_global_ids(pos + 3 * i, ids + 4096 * i);
            if (flavour == 0) build_chunk_mesh(c, m, all[i]); else build_chunk_mesh_fair(c, m, all[i]);
            counts[i] = static_cast<int64_t>(all[i].size());
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) total += counts[i];
    if (out_vertices != nullptr) {
        if (total > cap_vertices) return -1;
        auto* dst = static_cast<BlockVertex*>(out_vertices);
        for (int64_t i = 0; i < n; ++i) {
            std::memcpy(dst, all[i].data(), all[i].size() * sizeof(BlockVertex));
            dst += all[i].size();
        }
    }
    return total;
}

// CPU baseline: the reference's per-frame load loop (scene_game.rs:56-64) restated: for each chunk
// position, World::get_chunk-style generation (kind GEN_REFERENCE goes through the faithful
// string-palette generators; other kinds fill ids then replay them through set_block) followed by
// build_chunk_mesh and one memcpy standing in for the Vulkan upload (chunk_mesh.rs:111-124).
// Returns seconds of wall time; *out_vertices_total receives the vertex count.
double orc_time_pipeline(int64_t n, const int32_t* pos, int kind, uint32_t seed, const float* uv, uint32_t n_blocks,
                         int flavour, int threads, int include_generation, int64_t* out_vertices_total) {
    BlockUvMap m = uv_map_from_table(uv, n_blocks);
    std::vector<Chunk> pre;
    if (!include_generation) {
        pre.reserve(static_cast<size_t>(n));
        std::vector<uint16_t> ids(4096);
        for (int64_t i = 0; i < n; ++i) {
            if (kind == GEN_REFERENCE) pre.push_back(world_get_chunk(pos + 3 * i));
            else { generate_ids(kind, seed, pos + 3 * i, ids.data()); pre.push_back(chunk_from_global_ids(pos + 3 * i, ids.data())); }
        }
    }
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> total{0};
    if (threads < 1) threads = 1;
    auto work = [&]() {
        std::vector<uint16_t> ids(4096);
        std::vector<BlockVertex> upload(12288 * 6);  // stands in for the device allocation
        int64_t local = 0;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            std::vector<BlockVertex> verts;  // Vec::new() per call, chunk_mesh.rs:21
            bool some;
            if (include_generation) {
                Chunk c = kind == GEN_REFERENCE ? world_get_chunk(pos + 3 * i)
                                                : (generate_ids(kind, seed, pos + 3 * i, ids.data()),
                                                   chunk_from_global_ids(pos + 3 * i, ids.data()));
                some = flavour == 0 ? build_chunk_mesh(c, m, verts) : build_chunk_mesh_fair(c, m, verts);
            } else {
                some = flavour == 0 ? build_chunk_mesh(pre[i], m, verts) : build_chunk_mesh_fair(pre[i], m, verts);
            }
            if (some) {
                std::memcpy(upload.data(), verts.data(), verts.size() * sizeof(BlockVertex));
                local += static_cast<int64_t>(verts.size());
            }
        }
        total.fetch_add(local);
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (out_vertices_total) *out_vertices_total = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Prepared workload for repeated timing (bench.py): the Chunk objects (string palettes and all) are
// built once, outside any timed region, the way World keeps Arc<Chunk>s around (world.rs:11-12).
struct OrcWorkload {
    std::vector<Chunk> chunks;
    BlockUvMap uv;
};

void* orc_workload_create(int64_t n, const int32_t* pos, int kind, uint32_t seed, const float* uv, uint32_t n_blocks) {
    auto* w = new OrcWorkload();
    w->uv = uv_map_from_table(uv, n_blocks);
    w->chunks.reserve(static_cast<size_t>(n));
    std::vector<uint16_t> ids(4096);
    for (int64_t i = 0; i < n; ++i) {
        if (kind == GEN_REFERENCE) w->chunks.push_back(world_get_chunk(pos + 3 * i));
        else { generate_ids(kind, seed, pos + 3 * i, ids.data()); w->chunks.push_back(chunk_from_global_ids(pos + 3 * i, ids.data())); }
    }
    return w;
}
void orc_workload_destroy(void* h) { delete static_cast<OrcWorkload*>(h); }

// One pass of build_chunk_mesh over every chunk of the workload on `threads` host threads, each
// result followed by one memcpy standing in for the Vulkan upload (chunk_mesh.rs:111-124).
// Returns seconds; *out_vertices receives the vertex total.
double orc_workload_mesh(void* h, int flavour, int threads, int64_t* out_vertices) {
    auto* w = static_cast<OrcWorkload*>(h);
    const int64_t n = static_cast<int64_t>(w->chunks.size());
    std::atomic<int64_t> next{0}, total{0};
    if (threads < 1) threads = 1;
    auto work = [&]() {
        std::vector<BlockVertex> upload(12288 * 6);
        int64_t local = 0;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            std::vector<BlockVertex> verts;  // Vec::new() per call, chunk_mesh.rs:21
            const bool some = flavour == 0 ? build_chunk_mesh(w->chunks[i], w->uv, verts) : build_chunk_mesh_fair(w->chunks[i], w->uv, verts);
            if (some) {
                std::memcpy(upload.data(), verts.data(), verts.size() * sizeof(BlockVertex));
                local += static_cast<int64_t>(verts.size());
            }
        }
        total.fetch_add(local);
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (out_vertices) *out_vertices = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

int orc_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
