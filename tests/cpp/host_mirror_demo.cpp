// The reference's first frame, through the C++ host mirror: GameScene::update with the shipped settings
// (settings.json5: horizontal 10, vertical 4) and the start camera (2.5, 22.5, -2.5) -> chunk (0,1,-1)
// (engine/core/app.rs:203, engine/scenes/scene_game.rs:39-82).  Prints "chunks meshes vertices" and dumps
// the vertex stream of chunk (0,0,0) to argv[1].
#include <cstdio>

#include "../../despawnengine_b200/host/despawn_world.hpp"

int main(int argc, char** argv) {
    using namespace despawn;
    try {
        World world(4096, 512ull << 20);
        const BlockUvs block_uvs = {{"template:dirt", AtlasUV{{0.0f, 0.0f}, {0.5f, 1.0f}}}, {"template:engine", AtlasUV{{0.5f, 0.0f}, {1.0f, 1.0f}}}};
        const std::vector<ChunkPos> visible = visible_chunks({0, 1, -1}, 10, 4);
        for (const ChunkPos& p : visible) world.load_chunk(p);                   // Load (scene_game.rs:56-58)
        const auto meshes = build_chunk_meshes(world, visible, block_uvs);       // build_chunk_mesh x N, one call
        uint64_t n_mesh = 0, n_vert = 0;
        for (const auto& m : meshes) if (m) { ++n_mesh; n_vert += m->len(); }
        std::printf("%zu %llu %llu\n", visible.size(), (unsigned long long)n_mesh, (unsigned long long)n_vert);
        // a host chunk with its own palette order, as Chunk::set_block would build it
        Chunk c({40, 0, 0});
        c.set_block(1, 2, 3, "template:engine");
        c.set_block(0, 0, 0, "template:dirt");
        world.put_chunk(c);
        const auto m = build_chunk_mesh(world, {40, 0, 0}, block_uvs);
        std::printf("%u\n", m ? m->len() : 0u);
        {   // the same first frame through GameScene::update (one terrain launch + one mesh launch), then one chunk to the +x
            World w2(4096, 512ull << 20);
            GameScene scene(w2, 10, 4);
            const dsp_scene_stats a = scene.update({2.5f, 22.5f, -2.5f}, block_uvs);
            const dsp_scene_stats b = scene.update({18.5f, 22.5f, -2.5f}, block_uvs);
            std::printf("%llu %llu %llu %llu %llu %zu\n", (unsigned long long)a.n_loaded_total, (unsigned long long)a.n_meshes, (unsigned long long)a.vertex_total,
                        (unsigned long long)b.n_loaded_now, (unsigned long long)b.n_unloaded_now, scene.amount_of_chunk_meshes());
        }
        {   // host-resident chunks in bulk (build_all_loaded_chunks over Chunk structs): flat / full / empty as the reference's
            // generators fill them (chunk.rs:86-120), one emptied again -- 1,024 + 1,536 + 0 + 0 faces
            World w3(64, 64ull << 20);
            std::vector<Chunk> host(4, Chunk({0, 0, 0}));
            for (int i = 0; i < 4; ++i) host[i].position = {100 + i, 0, 0};
            for (int x = 0; x < 16; ++x) for (int z = 0; z < 16; ++z) for (int y = 0; y < 16; ++y) {
                if (y < 8) host[0].set_block(x, y, z, "template:dirt");
                host[1].set_block(x, y, z, "template:engine");
            }
            host[3].set_block(5, 5, 5, "template:dirt");
            host[3].set_block(5, 5, 5, AIR_BLOCK_ID);
            const auto hm = build_all_host_chunks(w3, host, block_uvs);
            std::printf("%u %u %u %u\n", hm[0] ? hm[0]->len() : 0u, hm[1] ? hm[1]->len() : 0u, hm[2] ? hm[2]->len() : 0u, hm[3] ? hm[3]->len() : 0u);
        }
        if (argc > 1) {
            const auto origin = build_chunk_mesh(world, {0, 0, 0}, block_uvs);
            const auto v = world.download(*origin);
            FILE* f = std::fopen(argv[1], "wb");
            std::fwrite(v.data(), sizeof(BlockVertex), v.size(), f);
            std::fclose(f);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
