// despawn_world.hpp -- header-only C++17 host mirror of the reference's chunk / world / mesh interface over the
// C ABI (include/despawn_b200.h).  Same names and argument meaning as the Rust items it stands in for, so that a
// caller written against the reference reads the same (paths under /root/reference/src):
//
//   despawn::Chunk             content/world/chunks/chunk.rs:15-82     position, blocks, palette, palette_map
//   despawn::World             content/world/world.rs:10-88            get_chunk, load_chunk, unload_chunk
//   despawn::build_chunk_mesh  content/world/chunks/chunk_mesh.rs:13   (chunk, block_uvs) -> optional vertex buffer
//   despawn::AtlasUV           engine/rendering/texture_atlas.rs:15-19
//   despawn::visible_chunks    engine/scenes/scene_game.rs:170-223     get_all_chunk_pos_in_render
//
// The host only resolves block-id strings to numbers and keeps the position -> slot map; every voxel and every
// vertex is produced on the GPU by libdespawn_b200.so.  Errors are exceptions carrying dsp_last_error().
#pragma once
#include <array>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/despawn_b200.h"

namespace despawn {

constexpr std::size_t CHUNK_SIZE = 16;                 // chunk.rs:7
constexpr std::size_t MAX_CHUNK_INDEX = 4095;          // chunk.rs:9
inline const std::string AIR_BLOCK_ID = "base:air";    // chunk.rs:12

struct AtlasUV {  // texture_atlas.rs:15-19
    std::array<float, 2> uv_min{0.0f, 0.0f};
    std::array<float, 2> uv_max{1.0f, 1.0f};
};
using BlockUvs = std::unordered_map<std::string, AtlasUV>;  // RapidHashMap<String, AtlasUV>

struct BlockVertex {  // engine/rendering/vertex.rs:5-12
    float position[3];
    float tex_coords[2];
};
static_assert(sizeof(BlockVertex) == DSP_VERTEX_BYTES, "BlockVertex layout");

using ChunkPos = std::array<int32_t, 3>;

// Host-resident chunk data with the reference's per-chunk string palette (chunk.rs:15-68).
struct Chunk {
    ChunkPos position{};
    std::vector<uint16_t> blocks = std::vector<uint16_t>(CHUNK_SIZE * CHUNK_SIZE * CHUNK_SIZE, 0);
    std::vector<std::string> palette{AIR_BLOCK_ID};
    std::unordered_map<std::string, uint16_t> palette_map{{AIR_BLOCK_ID, 0}};

    explicit Chunk(ChunkPos pos) : position(pos) {}
    static std::size_t index(std::size_t x, std::size_t y, std::size_t z) { return x + y * CHUNK_SIZE + z * CHUNK_SIZE * CHUNK_SIZE; }
    bool is_air_at_idx(std::size_t idx) const { return blocks[idx] == 0; }
    void set_block(std::size_t x, std::size_t y, std::size_t z, const std::string& block_id) {  // chunk.rs:49-68
        const std::size_t idx = index(x, y, z);
        if (block_id == AIR_BLOCK_ID) { blocks[idx] = 0; return; }
        auto it = palette_map.find(block_id);
        uint16_t pi;
        if (it != palette_map.end()) pi = it->second;
        else { pi = static_cast<uint16_t>(palette.size()); palette.push_back(block_id); palette_map.emplace(block_id, pi); }
        blocks[idx] = pi;
    }
};

// What build_chunk_mesh returns in place of Subbuffer<[BlockVertex]>: a window of the device vertex arena.
struct ChunkMesh {
    uint64_t offset_bytes = 0;
    uint32_t vertex_count = 0;
    uint32_t index_count = 0;  // > 0 with DSP_MESH_INDEXED: u32 indices at offset_bytes + align128(20 * vertex_count)
    uint32_t len() const { return vertex_count; }  // mesh.len() as used by draw(), scene_game.rs:118-122
};

class World {
public:
    World(uint64_t max_chunks, uint64_t arena_bytes, int device = 0, int gen_kind = DSP_GEN_REFERENCE, uint32_t seed = 0)
        : gen_kind_(gen_kind), seed_(seed) {
        if (dsp_create(device, max_chunks, arena_bytes, &ctx_) != DSP_OK) throw std::runtime_error(dsp_last_error(nullptr));
        block_ids_.emplace(AIR_BLOCK_ID, 0);
        register_block("template:dirt");    // assets/data/blocks/dirt.json5
        register_block("template:engine");  // assets/data/blocks/engine.json5
    }
    ~World() { dsp_destroy(ctx_); }
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    uint16_t register_block(const std::string& id) {
        auto it = block_ids_.find(id);
        if (it != block_ids_.end()) return it->second;
        const uint16_t n = static_cast<uint16_t>(block_ids_.size());
        block_ids_.emplace(id, n);
        return n;
    }

    // world.rs:28-47 -- cached, else generated ON THE DEVICE by the rule for this chunk position; returns the slot
    uint32_t get_chunk(const ChunkPos& pos) { return get_chunks({pos})[0]; }
    std::vector<uint32_t> get_chunks(const std::vector<ChunkPos>& positions) {
        std::vector<ChunkPos> missing;
        for (const ChunkPos& p : positions)
            if (!chunks.count(p)) { chunks.emplace(p, DSP_NO_SLOT); missing.push_back(p); }
        if (!missing.empty()) {
            std::vector<uint32_t> slots(missing.size());
            check(dsp_generate_chunks(ctx_, missing.size(), missing[0].data(), gen_kind_, seed_, slots.data()));
            for (std::size_t i = 0; i < missing.size(); ++i) chunks[missing[i]] = slots[i];
        }
        std::vector<uint32_t> out;
        out.reserve(positions.size());
        for (const ChunkPos& p : positions) out.push_back(chunks.at(p));
        return out;
    }
    // host chunk data -> device store; the per-chunk palette is resolved to registry ids here, remapped on the device
    uint32_t put_chunk(const Chunk& chunk) {
        std::vector<uint16_t> pal;
        for (const std::string& b : chunk.palette) pal.push_back(register_block(b));
        const uint32_t off[2] = {0u, static_cast<uint32_t>(pal.size())};
        uint32_t slot = DSP_NO_SLOT;
        check(dsp_load_chunks(ctx_, 1, chunk.position.data(), chunk.blocks.data(), pal.data(), off, &slot));
        chunks[chunk.position] = slot;
        return slot;
    }
    void load_chunk(const ChunkPos& pos) { loaded_chunks[pos] = get_chunk(pos); }  // world.rs:80-84
    void unload_chunk(const ChunkPos& pos) { loaded_chunks.erase(pos); }          // world.rs:85-87

    void set_block_uvs(const BlockUvs& block_uvs) {
        for (const auto& kv : block_uvs) register_block(kv.first);
        std::vector<float> rows(block_ids_.size() * 4, 0.0f);
        for (std::size_t i = 0; i < block_ids_.size(); ++i) { rows[4 * i + 2] = 1.0f; rows[4 * i + 3] = 1.0f; }  // fallback, chunk_mesh.rs:53-56
        for (const auto& kv : block_uvs) {
            const std::size_t i = block_ids_.at(kv.first);
            rows[4 * i] = kv.second.uv_min[0]; rows[4 * i + 1] = kv.second.uv_min[1];
            rows[4 * i + 2] = kv.second.uv_max[0]; rows[4 * i + 3] = kv.second.uv_max[1];
        }
        check(dsp_set_uv_table(ctx_, static_cast<uint32_t>(block_ids_.size()), rows.data()));
    }

    std::vector<BlockVertex> download(const ChunkMesh& m) {
        std::vector<BlockVertex> v(m.vertex_count);
        if (m.vertex_count) check(dsp_download_arena(ctx_, m.offset_bytes, uint64_t(m.vertex_count) * DSP_VERTEX_BYTES, v.data()));
        return v;
    }

    dsp_ctx* ctx() { return ctx_; }
    void check(int rc) const { if (rc != DSP_OK) throw std::runtime_error(dsp_last_error(ctx_)); }

    std::map<ChunkPos, uint32_t> chunks;         // pos -> slot        (world.rs:11)
    std::map<ChunkPos, uint32_t> loaded_chunks;  // world.rs:12
    uint64_t arena_top = 0;

private:
    dsp_ctx* ctx_ = nullptr;
    int gen_kind_;
    uint32_t seed_;
    std::unordered_map<std::string, uint16_t> block_ids_;
};

// GameScene::update's load loop (scene_game.rs:56-64) as one batched call.
inline std::vector<std::optional<ChunkMesh>> build_chunk_meshes(World& world, const std::vector<ChunkPos>& positions, const BlockUvs& block_uvs,
                                                                uint32_t mode = DSP_MESH_PARITY) {
    world.set_block_uvs(block_uvs);
    const std::vector<uint32_t> slots = world.get_chunks(positions);
    std::vector<dsp_mesh_entry> entries(slots.size());
    uint64_t total = 0;
    world.check(dsp_mesh_chunks(world.ctx(), slots.size(), slots.data(), mode, world.arena_top, entries.data(), &total));
    world.arena_top += total;
    std::vector<std::optional<ChunkMesh>> out;
    for (const dsp_mesh_entry& e : entries)
        out.push_back(e.vertex_count ? std::optional<ChunkMesh>(ChunkMesh{e.offset_bytes, e.vertex_count, e.index_count}) : std::nullopt);
    return out;
}
// chunk_mesh.rs:13-17 -- nullopt where the reference returns None
inline std::optional<ChunkMesh> build_chunk_mesh(World& world, const ChunkPos& chunk_pos, const BlockUvs& block_uvs) {
    return build_chunk_meshes(world, {chunk_pos}, block_uvs)[0];
}
// GameScene::build_all_loaded_chunks (scene_game.rs:231-244) for chunks the caller holds on the HOST: every Chunk (blocks + string
// palette) is loaded and meshed by ONE dsp_mesh_host_chunks_ex call.  The host resolves palette indices to registry ids (what the
// reference's host does around the loop too) and hands over the palette lengths: a chunk whose palette is only "base:air" is never
// fetched (chunk_mesh.rs:18-20).  Hand in `staging` registered with cudaHostRegister and the mesh kernel pulls the tiles itself.
inline std::vector<std::optional<ChunkMesh>> build_all_host_chunks(World& world, const std::vector<Chunk>& host_chunks, const BlockUvs& block_uvs,
                                                                   uint32_t mode = DSP_MESH_PARITY, std::vector<uint16_t>* staging = nullptr) {
    world.set_block_uvs(block_uvs);
    const std::size_t n = host_chunks.size();
    std::vector<uint16_t> local;
    std::vector<uint16_t>& gids = staging ? *staging : local;
    gids.resize(n * 4096);
    std::vector<int32_t> pos(3 * n);
    std::vector<uint32_t> palette_lens(n), slots(n);
    for (std::size_t i = 0; i < n; ++i) {
        const Chunk& c = host_chunks[i];
        std::copy(c.position.begin(), c.position.end(), pos.begin() + 3 * i);
        palette_lens[i] = static_cast<uint32_t>(c.palette.size());
        std::vector<uint16_t> pal;
        for (const std::string& b : c.palette) pal.push_back(world.register_block(b));
        for (std::size_t v = 0; v < 4096; ++v) gids[i * 4096 + v] = pal[c.blocks[v]];  // per-chunk palette index -> registry id (chunk.rs:58-65)
    }
    std::vector<dsp_mesh_entry> entries(n);
    uint64_t total = 0;
    world.check(dsp_mesh_host_chunks_ex(world.ctx(), n, pos.data(), gids.data(), palette_lens.data(), mode, world.arena_top, slots.data(), entries.data(), &total,
                                        nullptr, 0));
    world.arena_top += total;
    std::vector<std::optional<ChunkMesh>> out;
    for (std::size_t i = 0; i < n; ++i) {
        world.chunks[host_chunks[i].position] = slots[i];
        out.push_back(entries[i].vertex_count ? std::optional<ChunkMesh>(ChunkMesh{entries[i].offset_bytes, entries[i].vertex_count, entries[i].index_count}) : std::nullopt);
    }
    return out;
}

// scene_game.rs:170-223 get_all_chunk_pos_in_render: cylinder around the camera chunk; loop bounds min-1 .. max+1 (exclusive)
inline std::vector<ChunkPos> visible_chunks(const ChunkPos& cam, int32_t hori_dist, int32_t vert_dist) {
    std::vector<ChunkPos> out;
    for (int32_t x = cam[0] - hori_dist - 1; x < cam[0] + hori_dist + 1; ++x)
        for (int32_t y = cam[1] - vert_dist - 1; y < cam[1] + vert_dist + 1; ++y)
            for (int32_t z = cam[2] - hori_dist - 1; z < cam[2] + hori_dist + 1; ++z) {
                const int64_t dx = std::llabs(int64_t(x) - cam[0]), dy = std::llabs(int64_t(y) - cam[1]), dz = std::llabs(int64_t(z) - cam[2]);
                if (dy <= vert_dist && dx * dx + dz * dz <= int64_t(hori_dist) * hori_dist) out.push_back({x, y, z});
            }
    return out;
}

// GameScene (engine/scenes/scene_game.rs:27-32, 39-82, 101-125) over dsp_scene_update: `update` is the whole per-frame
// cycle (early-out while the camera stays in its chunk, visible set, batched load + mesh, unload), `draw_list` the
// chunk_meshes.values().flatten() walk of draw().
class GameScene {
public:
    GameScene(World& world, uint32_t horizontal_render_distance, uint32_t vertical_render_distance, int gen_kind = DSP_GEN_REFERENCE,
              uint32_t seed = 0, uint32_t mesh_mode = DSP_MESH_PARITY)
        : world_(world), cfg_{horizontal_render_distance, vertical_render_distance, gen_kind, seed, mesh_mode, 0u} {}
    dsp_scene_stats update(const std::array<float, 3>& camera_pos, const BlockUvs& block_uvs) {
        world_.set_block_uvs(block_uvs);
        dsp_scene_stats st{};
        world_.check(dsp_scene_update(world_.ctx(), camera_pos.data(), &cfg_, &st));
        return st;
    }
    std::vector<std::pair<ChunkPos, ChunkMesh>> draw_list() {
        uint64_t n = 0;
        world_.check(dsp_scene_draw_list(world_.ctx(), 0, nullptr, nullptr, &n));
        std::vector<int32_t> pos(3 * n);
        std::vector<dsp_mesh_entry> entries(n);
        world_.check(dsp_scene_draw_list(world_.ctx(), n, pos.data(), entries.data(), &n));
        std::vector<std::pair<ChunkPos, ChunkMesh>> out;
        for (uint64_t i = 0; i < n; ++i)
            out.push_back({ChunkPos{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, ChunkMesh{entries[i].offset_bytes, entries[i].vertex_count, entries[i].index_count}});
        return out;
    }
    size_t amount_of_chunk_meshes() { return draw_list().size(); }  // scene_game.rs:272-274

private:
    World& world_;
    dsp_scene_config cfg_;
};

}  // namespace despawn
