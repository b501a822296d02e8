// Scene layer: GameScene::update's visible set -> load -> mesh -> unload cycle on top of the batched calls
// (SURVEY.md section 8f-4).  Included at the end of dsp_api.cu (same translation unit, so it sees dsp_ctx and the
// helpers of the anonymous namespace).
//
// Reference (paths under /root/reference/src):
//   engine/scenes/scene_game.rs:39-82    update(): current chunk, early-out when it has not changed, visible set,
//                                        load + build_chunk_mesh of what is not loaded, unload of what left the set
//   engine/scenes/scene_game.rs:170-223  get_all_chunk_pos_in_render: x, y, z loops over [min-1, max], kept when
//                                        |dy| <= V and dx^2 + dz^2 <= H^2
//   engine/scenes/scene_game.rs:101-125  draw(): one draw per Some(mesh) of chunk_meshes
//   content/world/world.rs:80-87         load_chunk / unload_chunk
// The visible set is enumerated on the host, in the reference's order; the device does what scales with voxels -- ONE terrain
// launch and ONE mesh launch for all newly visible chunks of an update instead of the reference's chunk-at-a-time loop.
// After an update the loaded set IS the visible set of that update (everything visible was loaded, everything else unloaded,
// scene_game.rs:56-73), i.e. a fixed shape translated to the camera's chunk.  So "is this position loaded?" is arithmetic on
// (position - previous centre), not a hash lookup, and a chunk crossing costs two sweeps of that arithmetic plus map updates
// for the shell that actually changes: 0.5 / 7.0 / 40 ms per crossing at render distances 10,4 / 32,8 / 64,8 with the
// round-1 hash maps, see profiles/README.md for the figures now.  Per-chunk scene state lives in an array by voxel slot.
// Mesh storage: every update's batch lands in the largest free interval of the vertex arena (its size is the kernel's
// own total: a batch that does not fit reports the bytes it needs, nothing is truncated, the update is rolled back); an
// interval is returned to the free list when the last chunk of its batch is unloaded, and merged with its neighbours.

namespace {

// `x as i32` of Rust for f32: truncation toward zero, saturating, NaN -> 0
inline int32_t rust_f32_as_i32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return static_cast<int32_t>(v);
}
inline int32_t div_euclid16(int32_t a) { return a >> 4; }  // i32::div_euclid(16) == floor division

void current_chunk_of(const float cam[3], int32_t out[3]) {  // scene_game.rs:41-46
    for (int a = 0; a < 3; ++a) out[a] = div_euclid16(rust_f32_as_i32(cam[a]));
}

// scene_game.rs:170-223, same loop nest and order
void visible_chunks_of(const int32_t cur[3], uint32_t hori, uint32_t vert, std::vector<int32_t>& out) {
    out.clear();
    const int64_t h = hori, v = vert;
    for (int64_t x = cur[0] - h - 1; x < cur[0] + h + 1; ++x)
        for (int64_t y = cur[1] - v - 1; y < cur[1] + v + 1; ++y)
            for (int64_t z = cur[2] - h - 1; z < cur[2] + h + 1; ++z) {
                const uint64_t dx = static_cast<uint64_t>(x > cur[0] ? x - cur[0] : cur[0] - x);
                const uint64_t dy = static_cast<uint64_t>(y > cur[1] ? y - cur[1] : cur[1] - y);
                const uint64_t dz = static_cast<uint64_t>(z > cur[2] ? z - cur[2] : cur[2] - z);
                if (dy <= static_cast<uint64_t>(v) && dx * dx + dz * dz <= static_cast<uint64_t>(h) * static_cast<uint64_t>(h)) {
                    out.push_back(static_cast<int32_t>(x));
                    out.push_back(static_cast<int32_t>(y));
                    out.push_back(static_cast<int32_t>(z));
                }
            }
}

// is (cur + d) in the visible set around cur?  (the loop bounds of scene_game.rs:196-207 are implied: dx^2 + dz^2 <= H^2 gives |dx|, |dz| <= H)
inline bool in_visible_shape(int64_t dx, int64_t dy, int64_t dz, uint32_t hori, uint32_t vert) {
    const uint64_t ax = static_cast<uint64_t>(dx < 0 ? -dx : dx), ay = static_cast<uint64_t>(dy < 0 ? -dy : dy), az = static_cast<uint64_t>(dz < 0 ? -dz : dz);
    return ay <= vert && ax <= hori && az <= hori && ax * ax + az * az <= static_cast<uint64_t>(hori) * hori;
}

void scene_free_interval(dsp_ctx* ctx, uint64_t off, uint64_t bytes) {
    if (bytes == 0) return;
    auto next = ctx->scene_free.lower_bound(off);
    if (next != ctx->scene_free.begin()) {
        auto prev = std::prev(next);
        if (prev->first + prev->second == off) { off = prev->first; bytes += prev->second; ctx->scene_free.erase(prev); }
    }
    if (next != ctx->scene_free.end() && off + bytes == next->first) { bytes += next->second; ctx->scene_free.erase(next); }
    ctx->scene_free[off] = bytes;
}

void scene_fill_stats(dsp_ctx* ctx, const int32_t cur[3], uint32_t changed, uint64_t n_visible, uint64_t n_loaded_now, uint64_t n_unloaded_now,
                      dsp_scene_stats* out) {
    if (!out) return;
    out->current_chunk[0] = cur[0]; out->current_chunk[1] = cur[1]; out->current_chunk[2] = cur[2];
    out->changed = changed;
    out->n_visible = n_visible;
    out->n_loaded_now = n_loaded_now;
    out->n_unloaded_now = n_unloaded_now;
    out->n_loaded_total = ctx->scene_n_loaded;
    out->n_meshes = ctx->scene_meshes;
    out->vertex_total = ctx->scene_vertices;
    out->index_total = ctx->scene_indices;
    uint64_t free_bytes = 0;
    for (const auto& kv : ctx->scene_free) free_bytes += kv.second;
    out->arena_bytes_used = ctx->arena_bytes - free_bytes;
}

}  // namespace

extern "C" {

int dsp_visible_chunks(const float camera_pos[3], uint32_t horizontal, uint32_t vertical, uint64_t cap, int32_t* out_pos, uint64_t* out_n) {
    if (!camera_pos || !out_n) return DSP_ERR_BAD_ARG;
    int32_t cur[3];
    current_chunk_of(camera_pos, cur);
    std::vector<int32_t> vis;
    visible_chunks_of(cur, horizontal, vertical, vis);
    *out_n = vis.size() / 3;
    if (out_pos) {
        if (cap < vis.size() / 3) return DSP_ERR_BAD_ARG;
        std::memcpy(out_pos, vis.data(), vis.size() * sizeof(int32_t));
    }
    return DSP_OK;
}

int dsp_scene_reset(dsp_ctx* ctx) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    DSP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t s = 0; s < ctx->scene_chunks.size(); ++s) {
        if (!ctx->scene_chunks[s].loaded) continue;
        ctx->map.erase(ctx->slot_pos[s]);
        ctx->free_slots.push_back(s);
        ctx->resident[s] = 0;
        ctx->n_resident--;
    }
    // hand slots out again lowest first (free_slots is popped from the back): consecutive slots are one copy / no list
    std::sort(ctx->free_slots.begin(), ctx->free_slots.end(), std::greater<uint32_t>());
    if (ctx->scene_n_loaded) ctx->nbr_dirty = true;
    ctx->scene_chunks.clear();
    ctx->scene_n_loaded = 0;
    ctx->scene_batches.clear();
    ctx->scene_free_batch_ids.clear();
    ctx->scene_free.clear();
    ctx->scene_init = false;
    ctx->scene_has_last = false;
    ctx->scene_meshes = ctx->scene_vertices = ctx->scene_indices = 0;
    return DSP_OK;
}

int dsp_scene_update(dsp_ctx* ctx, const float camera_pos[3], const dsp_scene_config* cfg, dsp_scene_stats* out_stats) {
    if (!ctx) return DSP_ERR_BAD_ARG;
    if (!camera_pos || !cfg) return fail(ctx, DSP_ERR_BAD_ARG, "NULL argument");
    if (cfg->gen_kind < 0 || cfg->gen_kind > 3) return fail(ctx, DSP_ERR_BAD_ARG, "unknown gen_kind %d", cfg->gen_kind);
    if (cfg->mesh_mode & DSP_MESH_ORDERED) return fail(ctx, DSP_ERR_BAD_ARG, "DSP_MESH_ORDERED has no meaning for a scene (slots come and go)");
    if (!ctx->regions.empty())
        return fail(ctx, DSP_ERR_BAD_ARG, "a scene manages its own chunk slots: use a context without dsp_generate_region boxes");
    if (ctx->map.size() != ctx->scene_n_loaded)
        return fail(ctx, DSP_ERR_BAD_ARG, "a scene owns every chunk of its context: %llu chunks here were loaded by hand (dsp_generate_chunks / dsp_load_chunks); "
                    "unload them or use a separate context", (unsigned long long)(ctx->map.size() - ctx->scene_n_loaded));
    DSP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->scene_init) {
        ctx->scene_free.clear();
        ctx->scene_free[0] = ctx->arena_bytes & ~static_cast<uint64_t>(DSP_SLOT_ALIGN - 1);
        ctx->scene_init = true;
    }
    int32_t cur[3];
    current_chunk_of(camera_pos, cur);
    if (ctx->scene_has_last && cur[0] == ctx->scene_last[0] && cur[1] == ctx->scene_last[1] && cur[2] == ctx->scene_last[2]) {
        scene_fill_stats(ctx, cur, 0, 0, 0, 0, out_stats);  // scene_game.rs:48: nothing to do while the camera stays in its chunk
        return DSP_OK;
    }
    const uint32_t H = cfg->horizontal_render_distance, V = cfg->vertical_render_distance;
    const bool has_last = ctx->scene_has_last;
    const int64_t last[3] = {ctx->scene_last[0], ctx->scene_last[1], ctx->scene_last[2]};
    const uint32_t lastH = ctx->scene_last_h, lastV = ctx->scene_last_v;

    // ---- load (scene_game.rs:56-64): everything visible and not loaded, in the reference's order, as ONE batch.  Loaded = inside
    //      the previous update's shape around the previous centre.
    std::vector<int32_t> todo;
    uint64_t n_vis = 0;
    {
        const int64_t h = H, v = V;
        for (int64_t x = cur[0] - h - 1; x < cur[0] + h + 1; ++x)
            for (int64_t y = cur[1] - v - 1; y < cur[1] + v + 1; ++y)
                for (int64_t z = cur[2] - h - 1; z < cur[2] + h + 1; ++z) {
                    if (!in_visible_shape(x - cur[0], y - cur[1], z - cur[2], H, V)) continue;
                    ++n_vis;
                    if (has_last && in_visible_shape(x - last[0], y - last[1], z - last[2], lastH, lastV)) continue;
                    todo.push_back(static_cast<int32_t>(x)); todo.push_back(static_cast<int32_t>(y)); todo.push_back(static_cast<int32_t>(z));
                }
    }
    const uint64_t n_new = todo.size() / 3;
    if (n_new) {
        if (ctx->n_resident + n_new > ctx->max_chunks)
            return fail(ctx, DSP_ERR_STORE_FULL, "scene needs %llu more chunk slots, %llu free", (unsigned long long)n_new,
                        (unsigned long long)(ctx->max_chunks - ctx->n_resident));
        ctx->map.reserve(ctx->map.size() + n_new);  // one rehash at most, instead of several while thousands of chunks arrive
        std::vector<uint32_t> slots(n_new);
        int rc = dsp_generate_chunks(ctx, n_new, todo.data(), cfg->gen_kind, cfg->seed, slots.data());  // World::load_chunk -> get_chunk
        if (rc != DSP_OK) return rc;
        std::vector<dsp_mesh_entry> entries(n_new);
        uint64_t total = 0;
        bool done = false;
        // The batch's size is only known once the kernel has counted, so it is offered the largest free interval; if even
        // that is too small the call reports the bytes needed (nothing is truncated) and the update is rolled back.
        auto it = ctx->scene_free.begin();
        for (auto j = ctx->scene_free.begin(); j != ctx->scene_free.end(); ++j)
            if (j->second > it->second) it = j;
        if (it != ctx->scene_free.end()) {
            const uint64_t off = it->first, room = it->second;
            ctx->arena_limit = off + room;
            rc = dsp_mesh_chunks(ctx, n_new, slots.data(), cfg->mesh_mode, off, entries.data(), &total);
            ctx->arena_limit = 0;
            if (rc == DSP_OK) {
                ctx->scene_free.erase(it);
                if (room > total) ctx->scene_free[off + total] = room - total;
                uint32_t b;
                if (!ctx->scene_free_batch_ids.empty()) { b = ctx->scene_free_batch_ids.back(); ctx->scene_free_batch_ids.pop_back(); }
                else { b = static_cast<uint32_t>(ctx->scene_batches.size()); ctx->scene_batches.push_back({}); }
                ctx->scene_batches[b] = dsp_ctx::SceneBatch{off, total, static_cast<uint32_t>(n_new)};
                if (ctx->scene_chunks.size() < ctx->next_unused) ctx->scene_chunks.resize(ctx->next_unused, dsp_ctx::SceneChunk{0u, 0u, dsp_mesh_entry{}});
                ctx->scene_n_loaded += n_new;
                for (uint64_t i = 0; i < n_new; ++i) {
                    ctx->scene_chunks[slots[i]] = dsp_ctx::SceneChunk{1u, b, entries[i]};
                    if (entries[i].vertex_count) {
                        ctx->scene_meshes++;
                        ctx->scene_vertices += entries[i].vertex_count;
                        ctx->scene_indices += entries[i].index_count;
                    }
                }
                done = true;
            }
        }
        if (!done) {
            // give the voxel slots back: the update did not happen (whatever the reason -- every error path rolls back)
            const std::string why = ctx->err;
            cudaStreamSynchronize(ctx->stream);
            for (uint64_t i = 0; i < n_new; ++i) {
                ctx->map.erase(PosKey{todo[3 * i], todo[3 * i + 1], todo[3 * i + 2]});
                ctx->free_slots.push_back(slots[i]);
                ctx->resident[slots[i]] = 0;
                ctx->n_resident--;
            }
            ctx->nbr_dirty = true;
            if (rc != DSP_OK && rc != DSP_ERR_ARENA_FULL) { ctx->err = why; return rc; }
            return fail(ctx, DSP_ERR_ARENA_FULL, "no free arena interval of %llu bytes for the meshes of %llu newly visible chunks",
                        (unsigned long long)total, (unsigned long long)n_new);
        }
    }

    // ---- unload (scene_game.rs:67-73): loaded chunks that left the visible set = the previous shape around the previous centre,
    //      minus the new shape around the new centre
    uint64_t n_unloaded = 0;
    if (has_last) {
        const int64_t h = lastH, v = lastV;
        for (int64_t x = last[0] - h; x <= last[0] + h; ++x)
            for (int64_t z = last[2] - h; z <= last[2] + h; ++z) {
                if (!in_visible_shape(x - last[0], 0, z - last[2], lastH, lastV)) continue;
                for (int64_t y = last[1] - v; y <= last[1] + v; ++y) {
                    if (in_visible_shape(x - cur[0], y - cur[1], z - cur[2], H, V)) continue;
                    auto it = ctx->map.find(PosKey{static_cast<int32_t>(x), static_cast<int32_t>(y), static_cast<int32_t>(z)});
                    if (it == ctx->map.end()) continue;  // (cannot happen: the loaded set is the previous visible set)
                    const uint32_t slot = it->second;
                    dsp_ctx::SceneChunk& c = ctx->scene_chunks[slot];
                    if (c.entry.vertex_count) {
                        ctx->scene_meshes--;
                        ctx->scene_vertices -= c.entry.vertex_count;
                        ctx->scene_indices -= c.entry.index_count;
                    }
                    dsp_ctx::SceneBatch& b = ctx->scene_batches[c.batch];
                    if (--b.live == 0) {
                        scene_free_interval(ctx, b.offset, b.bytes);
                        ctx->scene_free_batch_ids.push_back(c.batch);
                    }
                    c.loaded = 0u;
                    ctx->map.erase(it);                // World::unload_chunk; the voxel slot is recycled (the reference keeps the Arc<Chunk> in
                    ctx->free_slots.push_back(slot);   // `chunks` forever, world.rs:11,44-46 -- regeneration is deterministic, so a chunk
                    ctx->resident[slot] = 0;           // that comes back into view is identical)
                    ctx->n_resident--;
                    ctx->scene_n_loaded--;
                    ctx->nbr_dirty = true;
                    ++n_unloaded;
                }
            }
    }
    ctx->scene_last_h = H; ctx->scene_last_v = V;
    ctx->scene_last[0] = cur[0]; ctx->scene_last[1] = cur[1]; ctx->scene_last[2] = cur[2];
    ctx->scene_has_last = true;
    scene_fill_stats(ctx, cur, 1, n_vis, n_new, n_unloaded, out_stats);
    return DSP_OK;
}

int dsp_scene_draw_list(dsp_ctx* ctx, uint64_t cap, int32_t* out_pos, dsp_mesh_entry* out_entries, uint64_t* out_n) {
    if (!ctx || !out_n) return DSP_ERR_BAD_ARG;
    *out_n = ctx->scene_meshes;
    if (!out_pos && !out_entries) return DSP_OK;
    if (cap < ctx->scene_meshes) return fail(ctx, DSP_ERR_BAD_ARG, "draw list holds %llu meshes, capacity %llu", (unsigned long long)ctx->scene_meshes, (unsigned long long)cap);
    uint64_t k = 0;
    for (uint32_t s = 0; s < ctx->scene_chunks.size(); ++s) {  // chunk_meshes.values().flatten(): the reference's order is its hash map's, i.e. unspecified; here by slot
        const dsp_ctx::SceneChunk& c = ctx->scene_chunks[s];
        if (!c.loaded || !c.entry.vertex_count) continue;
        if (out_pos) { out_pos[3 * k] = ctx->slot_pos[s].x; out_pos[3 * k + 1] = ctx->slot_pos[s].y; out_pos[3 * k + 2] = ctx->slot_pos[s].z; }
        if (out_entries) out_entries[k] = c.entry;
        ++k;
    }
    return DSP_OK;
}

}  // extern "C"
