/* despawn_b200.h -- C ABI of the B200-native chunk terrain -> cull -> mesh path.
 *
 * One shared library (libdespawn_b200.so), plain pointers and sizes, no C++/torch types.  The
 * reference engine has no FFI boundary today: the seam this library replaces
 * is three Rust items, cited per entry point below as file:line under /root/reference/src.
 *
 *   World::get_chunk / load_chunk / unload_chunk     content/world/world.rs:28-47, 80-87
 *   Chunk { position, blocks: Vec<u16>, palette }    content/world/chunks/chunk.rs:15-20
 *   build_chunk_mesh(allocator, &Chunk, &block_uvs)  content/world/chunks/chunk_mesh.rs:13-129
 *   BlockVertex { position: Vec3, tex_coords }       engine/rendering/vertex.rs:5-12   (20 bytes)
 *
 * Data contract
 *   voxels    u16 per voxel, 4096 per chunk, index x + 16*y + 256*z (chunk.rs:44-46); 0 = air
 *             (chunk.rs:39-41).  The device store holds GLOBAL block ids: the host shim resolves each
 *             chunk's string palette (chunk.rs:58-65) to registry ids before upload, or passes the
 *             palette so that dsp_load_chunks remaps on the device.
 *   uv table  row id = (uv_min.x, uv_min.y, uv_max.x, uv_max.y) of AtlasUV
 *             (engine/rendering/texture_atlas.rs:15-19); row 0 (air) is never read; ids >= n_blocks get the
 *             reference's fallback (0,0)-(1,1) (chunk_mesh.rs:53-56).
 *   vertices  tightly packed BlockVertex, 6 per face, in exactly the reference's emission order (voxel
 *             index ascending; TOP, BOTTOM, REAR, FRONT, RIGHT, LEFT per voxel, chunk_mesh.rs:32,65-106).
 *             Each chunk's stream starts at a 128-byte aligned offset of the arena.
 *             vertex_count == 0  <=>  the reference returns None (chunk_mesh.rs:18-20,109-128).
 *
 * Threading: a context is single-caller (the reference meshes on the winit thread only,
 * engine/scenes/scene_game.rs:39-82); one context per GPU.  No function panics or aborts; every
 * function returns a dsp_status and dsp_last_error() explains the last failure.
 * There is no CPU fallback: without a CUDA device dsp_create fails with DSP_ERR_NO_DEVICE.
 */
#ifndef DESPAWN_B200_H
#define DESPAWN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSP_ABI_VERSION 1

typedef enum dsp_status {
    DSP_OK = 0,
    DSP_ERR_BAD_ARG = 1,
    DSP_ERR_CUDA = 2,         /* a CUDA call failed; see dsp_last_error */
    DSP_ERR_NO_DEVICE = 3,    /* no usable CUDA device / driver */
    DSP_ERR_ARENA_FULL = 4,   /* vertex arena too small: nothing is truncated, *total_bytes = bytes needed */
    DSP_ERR_STORE_FULL = 5,   /* no free chunk slot (max_chunks reached) */
    DSP_ERR_NOT_RESIDENT = 6, /* chunk position / slot not loaded */
    DSP_ERR_INTERNAL = 7
} dsp_status;

/* Terrain generators.  REFERENCE is world.rs:35-42 (chunk Y > 0 empty, == 0 flat, < 0 full, all
 * "template:dirt" = id 1).  The others are builder-defined integer specs (DESIGN.md "Terrain spec");
 * the reference declares the `noise` crate but never calls it. */
typedef enum dsp_gen_kind {
    DSP_GEN_REFERENCE = 0,
    DSP_GEN_NOISE = 1,   /* 3-octave integer value-noise heightmap, ids 1 (body) and 2 (surface) */
    DSP_GEN_RANDOM = 2,  /* Bernoulli(0.5) per voxel from a 32-bit hash, ids 1 and 2 */
    DSP_GEN_CHECKER = 3  /* (wx+wy+wz) even -> id 1: the 12,288-face maximum */
} dsp_gen_kind;

/* Mesh modes (bit flags).  0 is the reference's behaviour and the only bit-exact-to-reference mode. */
#define DSP_MESH_PARITY 0u /* chunk-local cull, border faces always emitted (chunk_mesh.rs:46-51) */
#define DSP_MESH_HALO 1u   /* extension: border faces tested against resident neighbour chunks */
#define DSP_MESH_ORDERED 2u /* pack the chunks' arena slots back to back in the order of the call (reproducible layout: a count +
                             * scan pass fixes every offset first; 89 vs 72 us on the bench region); without it slots are claimed
                             * in completion order */

/* Quad modes (extensions; spec with the oracle, oracle/mesher_oracle.cpp).  Not available with
 * DSP_MESH_ORDERED. */
#define DSP_MESH_GREEDY 4u  /* greedy run-stack merge: maximal runs of equal-id visible faces along one in-plane axis, identical
                             * runs stacked along the other; re-expanded cell by cell the quads are exactly the faces the
                             * reference emits (chunk_mesh.rs:46-51,65-106).  Quads leave ordered by their anchor voxel. */
#define DSP_MESH_INDEXED 8u /* 4 vertices (template corners c0..c3, cube.rs:156-312) + 6 u32 indices 4q+(0,1,2,2,3,0) per quad
                             * instead of 6 vertices.  The chunk's slot holds the vertices, then (128-byte aligned) the indices:
                             * index bytes start at offset_bytes + align128(20 * vertex_count); indices are chunk-local. */

#define DSP_NO_SLOT 0xFFFFFFFFu
#define DSP_CHUNK_VOXELS 4096
#define DSP_VERTEX_BYTES 20
#define DSP_FACE_BYTES 120
#define DSP_QUAD_VERTEX_BYTES_INDEXED 80
#define DSP_QUAD_INDEX_BYTES 24
#define DSP_SLOT_ALIGN 128

typedef struct dsp_ctx dsp_ctx;

/* Result of meshing one chunk: where its vertex stream lives in the arena.  16 bytes. */
typedef struct dsp_mesh_entry {
    uint64_t offset_bytes;  /* from the start of the arena; multiple of DSP_SLOT_ALIGN */
    uint32_t vertex_count;  /* 6 * faces (4 * quads with DSP_MESH_INDEXED); 0 <=> None */
    uint32_t index_count;   /* 6 * quads with DSP_MESH_INDEXED, else 0 */
} dsp_mesh_entry;

/* One block edit in world voxel coordinates (world.rs:50-74 addressing: chunk = div_euclid(w,16),
 * local = rem_euclid(w,16)); block is a global id, 0 destroys. */
typedef struct dsp_edit {
    int32_t wx, wy, wz;
    uint32_t block;
} dsp_edit;

/* ---- lifetime ---------------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device`: a voxel store of max_chunks slots (8 KiB each) and a
 * vertex arena of arena_bytes.  Replaces World::new + World::set_allocator (world.rs:18-25, 76-78). */
int dsp_create(int device, uint64_t max_chunks, uint64_t arena_bytes, dsp_ctx** out_ctx);
/* dsp_create with options.  DSP_CREATE_EXPORTABLE_ARENA: the vertex arena is a CUDA virtual-memory allocation that
 * dsp_arena_export_fd can hand to another API or process (see "arena export" below); everything else is unchanged. */
#define DSP_CREATE_EXPORTABLE_ARENA 1u
int dsp_create_ex(int device, uint64_t max_chunks, uint64_t arena_bytes, uint32_t flags, dsp_ctx** out_ctx);
int dsp_destroy(dsp_ctx* ctx);
const char* dsp_last_error(const dsp_ctx* ctx); /* never NULL; ctx may be NULL for create failures */
const char* dsp_status_string(int status);
int dsp_abi_version(void);

/* ---- block_uvs ----------------------------------------------------------------------------------- */
/* Replaces the &RapidHashMap<String,AtlasUV> argument of build_chunk_mesh (chunk_mesh.rs:16). */
int dsp_set_uv_table(dsp_ctx* ctx, uint32_t n_blocks, const float* uv_rows /* n_blocks x 4 */);

/* ---- chunk residency (World) ----------------------------------------------------------------------- */
/* World::get_chunk + load_chunk with on-device generation (world.rs:28-47, 80-84; chunk.rs:86-120).
 * pos: n x 3 chunk coordinates.  As in the reference (world.rs:29-31: `if let Some(chunk_data) = self.chunks.get(&pos)
 * { return Ok(chunk_data.clone()) }`) a position that is already resident is returned AS IT IS -- its slot, its voxels,
 * edits included -- and only absent positions are generated; a position listed twice is one chunk.
 * out_slots (n, may be NULL) receives the slot of each chunk.  On DSP_ERR_STORE_FULL nothing was loaded. */
int dsp_generate_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, int gen_kind, uint32_t seed, uint32_t* out_slots);
/* Same, but resident positions are generated again in place, discarding edits (no counterpart in the reference, whose
 * `chunks` map never forgets; used by benches that want the generator inside the timed region). */
int dsp_regenerate_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, int gen_kind, uint32_t seed, uint32_t* out_slots);

/* Bulk form: a dims[0] x dims[1] x dims[2] box of chunks with minimum corner `origin`, generated into
 * consecutive slots; chunk (i,j,k) of the box gets slot *out_first_slot + i + dims[0]*(j + dims[1]*k).
 * A position holds one chunk: a box that overlaps another box or any resident position is DSP_ERR_BAD_ARG. */
int dsp_generate_region(dsp_ctx* ctx, const int32_t origin[3], const uint32_t dims[3], int gen_kind, uint32_t seed,
                        uint32_t* out_first_slot);
/* A second box with the voxels of an existing one (device-to-device copy) at another origin -- "paste this structure
 * there".  src_first_slot is the first slot of a fully resident dsp_generate_region / dsp_clone_region box. */
int dsp_clone_region(dsp_ctx* ctx, uint32_t src_first_slot, const int32_t new_origin[3], uint32_t* out_first_slot);

/* World::load_chunk for host-resident Chunk data (the reference's Chunk.blocks, chunk.rs:17).
 * blocks: n x 4096 u16 in host memory.  If palette_ids != NULL the values are per-chunk palette
 * indices and chunk i's palette is palette_ids[palette_offsets[i] .. palette_offsets[i+1]) (global id
 * per palette index, entry 0 = air = 0); the remap runs on the device.  If NULL, values are global ids.
 * A resident position's voxels are replaced.  A position may appear once per call (DSP_ERR_BAD_ARG otherwise; nothing is
 * loaded then). */
int dsp_load_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, const uint16_t* palette_ids,
                    const uint32_t* palette_offsets, uint32_t* out_slots);

/* World::unload_chunk (world.rs:85-87): frees the slots. */
int dsp_unload_chunks(dsp_ctx* ctx, uint64_t n, const uint32_t* slots);
int dsp_find_chunk(const dsp_ctx* ctx, const int32_t pos[3], uint32_t* out_slot); /* DSP_ERR_NOT_RESIDENT if absent */
/* Tells the library that the caller changed the voxels of these slots through dsp_voxel_device_ptr (interop writers):
 * the per-chunk summaries that let DSP_MESH_HALO / large DSP_MESH_GREEDY calls skip all-air and buried chunks are
 * reset to "unknown".  Every writer inside the library does this itself. */
int dsp_invalidate_chunks(dsp_ctx* ctx, uint64_t n, const uint32_t* slots);
uint64_t dsp_resident_chunks(const dsp_ctx* ctx);

/* ---- meshing (build_chunk_mesh) ---------------------------------------------------------------------- */
/* Batched build_chunk_mesh (chunk_mesh.rs:13-129); n = 1 is the reference call.  Meshes the chunks in
 * `slots` (n entries) into the arena starting at arena_offset (multiple of 128).  Every chunk gets a
 * 128-byte aligned slot of align128(20 * vertex_count) bytes; the slots are gap-free and together occupy
 * exactly *out_total_bytes.  With DSP_MESH_ORDERED they follow each other in the order given, otherwise in
 * completion order (the reference allocates one buffer per chunk, so slot order is not observable there).
 * out_entries (host, n entries, may be NULL) and *out_total_bytes (may be NULL) are filled
 * when the call returns; the vertex data stays on the device (the reference's result is a device
 * buffer too, chunk_mesh.rs:111-124).  Blocking. */
int dsp_mesh_chunks(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, uint32_t mode, uint64_t arena_offset,
                    dsp_mesh_entry* out_entries, uint64_t* out_total_bytes);
/* Same for the consecutive slots first_slot .. first_slot+n-1 (no slot list crosses the bus). */
int dsp_mesh_range(dsp_ctx* ctx, uint32_t first_slot, uint64_t n, uint32_t mode, uint64_t arena_offset,
                   dsp_mesh_entry* out_entries, uint64_t* out_total_bytes);
/* build_chunk_mesh for HOST-resident chunks in one call: World::load_chunk (host voxel data, global ids) followed
 * by the mesh of the same chunks, pipelined in sub-batches so that the host->device copy of sub-batch k+1 overlaps
 * the meshing of sub-batch k (pass pinned host memory for `blocks` to get a truly asynchronous copy).  This is the
 * end-to-end entry point for a caller that, like the reference, holds Chunk.blocks in host memory
 * (chunks/chunk.rs:17, scene_game.rs:56-64).  out_slots (n, may be NULL) receives the slot of each chunk.  Blocking.
 * A position may appear once per call.  On any error the call leaves no trace: work already enqueued is drained and
 * every slot it created is released. */
int dsp_mesh_host_chunks(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, uint32_t mode, uint64_t arena_offset,
                         uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes);

/* The same with the vertices delivered to HOST memory, for a consumer that is not on this GPU (B200 has no raster units:
 * whatever draws the meshes -- the reference's Vulkan device, chunk_mesh.rs:111-124 -- sits across PCIe unless it imports
 * the arena, see dsp_arena_export_fd).  host_vertices (pinned for a truly asynchronous copy) receives a byte-for-byte mirror of
 * the arena window [arena_offset, arena_offset + *out_total_bytes): chunk i's data is at host_vertices +
 * (out_entries[i].offset_bytes - arena_offset).  Three streams: the upload of sub-batch k+1, the meshing of sub-batch k and the
 * download of sub-batch k-1 overlap (PCIe is full duplex).  With 6-vertex unit quads the download is ~12x the upload and
 * bounds the call; DSP_MESH_GREEDY | DSP_MESH_INDEXED output is ~15x smaller and the call returns to the upload's pace. */
int dsp_mesh_host_chunks_to_host(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, uint32_t mode, uint64_t arena_offset,
                                 uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes, void* host_vertices,
                                 uint64_t host_capacity);

/* The general form of the two calls above, with what the reference's Chunk also carries: palette_lens[i] = Chunk::palette.len()
 * of chunk i (chunks/chunk.rs:11,58-65; NULL: not known).  A chunk whose palette has ONE entry has never held anything but air:
 * build_chunk_mesh returns None for it without looking at a voxel (chunk_mesh.rs:18-20), and so does this call -- the chunk's
 * `blocks` are not fetched (they must still be addressable; by the reference's invariant they are all zero), its slot of the voxel
 * store is zero-filled on the device and its entry is empty.  host_vertices may be NULL (then host_capacity is ignored and the
 * call is dsp_mesh_host_chunks with palette lengths).
 * When `blocks` is pinned or registered host memory that the device can address (cudaHostAlloc / cudaHostRegister), host_vertices
 * is NULL and the palette lengths spare at least a tenth of the chunks, no copy engine is involved: the mesh kernel itself pulls each remaining chunk's
 * 8 KiB tile over PCIe with the bulk (TMA) copy that otherwise reads the voxel store, meshes it and files it in the store on the
 * way -- two launches per call, nothing staged, nothing copied twice.  Otherwise (pageable memory, no palette lengths) the
 * copy-engine pipeline described above is used.  Both give identical results.
 * PACKED UPLOAD (the default for calls of 256 .. 131,072 chunks without DSP_MESH_ORDERED / DSP_MESH_HALO, with or without palette
 * lengths, pinned or pageable `blocks`; with host_vertices only when that buffer is pinned, where the download is device-driven): PCIe, not the GPU, bounds these calls, and the host has the voxels
 * in memory anyway -- so a few host worker threads (the CPUs the calling thread may run on, minus one, at most 15;
 * DSP_HOST_THREADS) first reduce every chunk to the smallest lossless form that holds it: nothing at all for a chunk of one id
 * (air included), 2 / 4 / 8 bits per voxel when its ids fit, the 8 KiB as they are otherwise.  A small kernel expands the forms
 * into the voxel store (and leaves the chunk summaries a generator would have left for air / one-id chunks), the ordinary mesh
 * kernel follows; three pieces per call, packing of piece k + 1 under the GPU work of piece k -- in chunk-local parity mode and up to 16,384 chunks ONE
 * mesh launch over the whole batch, started behind the first piece, whose CTAs wait (bounded) for the piece a chunk belongs to.  On the bench region 33.5 MB of voxels
 * become 1.3 MB on the bus and the call takes 0.21 ms instead of 0.51 ms.  Results are identical to the other paths (the chunks
 * are resident afterwards, bit for bit).  With eight worker threads or fewer and `blocks` in device-addressable memory the chunks
 * are shared out between the host cores and the bus (of every 16, 8 or 12 are packed, the others are fetched by the unpack kernel
 * as they are): eight ranks on a 32-CPU box, all reading host memory at once, pack slower than PCIe copies.  DSP_E2E_PACK=0 switches
 * the packed upload off, =2 packs everything whatever the core count; dsp_host_upload_stats reports what a call did.  The worker
 * threads belong to the context: created by its first such call, awake (spinning) for about a millisecond after a call so that
 * back-to-back calls find them ready, asleep after that, joined by dsp_destroy.  A process that runs several contexts side by side
 * should give each its own CPUs (sched_setaffinity before the first call) or set DSP_HOST_THREADS. */
int dsp_mesh_host_chunks_ex(dsp_ctx* ctx, uint64_t n, const int32_t* pos, const uint16_t* blocks, const uint32_t* palette_lens, uint32_t mode,
                            uint64_t arena_offset, uint32_t* out_slots, dsp_mesh_entry* out_entries, uint64_t* out_total_bytes,
                            void* host_vertices, uint64_t host_capacity);

/* Asynchronous forms: enqueue on the context's stream and return; results are read with
 * dsp_mesh_result after dsp_sync.  Let a caller overlap meshing with its own work (the reference is
 * synchronous inside GameScene::update, scene_game.rs:56-64). */
int dsp_mesh_range_async(dsp_ctx* ctx, uint32_t first_slot, uint64_t n, uint32_t mode, uint64_t arena_offset);
int dsp_mesh_chunks_async(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, uint32_t mode, uint64_t arena_offset);
int dsp_sync(dsp_ctx* ctx);
int dsp_mesh_result(dsp_ctx* ctx, dsp_mesh_entry* out_entries /* n of the last mesh call, may be NULL */,
                    uint64_t* out_total_bytes);

/* ---- edits (Chunk::set_block on resident chunks, chunk.rs:49-68) ---------------------------------- */
/* Applies n edits in order to the device voxel store.  Edits that fall in non-resident chunks are
 * ignored.  out_dirty_slots (capacity cap_dirty, may be NULL) receives the distinct slots touched, in
 * ascending slot order, and *out_n_dirty their number. */
int dsp_apply_edits(dsp_ctx* ctx, uint64_t n, const dsp_edit* edits, uint32_t* out_dirty_slots, uint64_t cap_dirty,
                    uint64_t* out_n_dirty);

/* Where the edits are resolved.  DEVICE pipeline (world -> (slot, voxel) addressing, last-edit-wins and the dirty list
 * all in kernels; 10,000 edits + remesh in ~0.4 ms) when ALL of these hold: every resident chunk belongs to a
 * dsp_generate_region / dsp_clone_region box, there are at most 8 boxes, no slot of a box has been unloaded, and the call
 * carries fewer than 2^20 edits.  Otherwise HOST addressing (one hash-map probe per edit, ~4x slower at 10,000 edits);
 * the results are identical. */
/* dsp_apply_edits followed by the whole-chunk remesh of the dirty chunks (the reference's only granularity:
 * build_chunk_mesh, chunk_mesh.rs:13-129) in one call.  When every resident chunk belongs to a dsp_generate_region box,
 * world -> (slot, voxel) addressing, last-edit-wins de-duplication, the voxel writes and the ascending dirty list all run
 * on the device and the mesh kernel takes the list from device memory; otherwise addressing runs on the host.
 * With DSP_MESH_HALO an edit on a chunk's border layer also dirties the resident neighbour across that border (its border
 * faces are culled against the edited voxel); neighbours on other devices are the caller's to re-exchange (dsp_pack_border_slabs).
 * out_entries[i] is the mesh of out_dirty_slots[i].  Blocking. */
int dsp_apply_edits_and_remesh(dsp_ctx* ctx, uint64_t n, const dsp_edit* edits, uint32_t mode, uint64_t arena_offset,
                               uint32_t* out_dirty_slots, uint64_t cap_dirty, uint64_t* out_n_dirty, dsp_mesh_entry* out_entries,
                               uint64_t* out_total_bytes);

/* World::get_block_world (world.rs:50-65) / Chunk::get_block (chunk.rs:71-82) on the device store: the global block id at world
 * voxel coordinates (chunk = div_euclid(w, 16), local = rem_euclid(w, 16), world.rs:69-74).  DSP_ERR_NOT_RESIDENT where the
 * reference returns None because the chunk is not in `chunks`.  Blocking (one 2-byte read-back). */
int dsp_get_block_world(dsp_ctx* ctx, int32_t wx, int32_t wy, int32_t wz, uint16_t* out_block);

/* ---- region seams (multi-GPU halo mode) ----------------------------------------------------- */
/* Faces are numbered in the reference's emission order (chunk_mesh.rs:46-51, 65-106). */
#define DSP_FACE_TOP 0    /* +Y, layer y = 15 */
#define DSP_FACE_BOTTOM 1 /* -Y, layer y = 0 */
#define DSP_FACE_REAR 2   /* -Z, layer z = 0 */
#define DSP_FACE_FRONT 3  /* +Z, layer z = 15 */
#define DSP_FACE_RIGHT 4  /* -X, layer x = 0 */
#define DSP_FACE_LEFT 5   /* +X, layer x = 15 */
/* Extracts the one-voxel-thick border layer at `face` of each listed chunk into dst_device (device memory owned by
 * the caller, n x 256 u16 = 512 bytes per chunk; layout +-Y: [z][x], +-Z: [y][x], +-X: [z][y]).  Asynchronous on the
 * context's stream (dsp_sync before handing dst_device to NCCL / a peer copy on another stream). */
int dsp_pack_border_slabs(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, int face, void* dst_device);
/* Registers, for each listed resident chunk, the border layer of its NON-resident neighbour across `face` (as packed
 * by the owning device with the opposite face).  DSP_MESH_HALO then culls against it exactly as against a resident
 * neighbour.  src_device: n x 512 bytes of device memory, copied.
 * Lifetime: registered slabs stay until dsp_clear_halo_slabs or until their chunk is unloaded.  Registering the SAME
 * (face, slot list) again -- the re-exchange after an edit on the seam -- replaces the slabs in place; a different list
 * adds a batch (newest wins where they overlap), so call dsp_clear_halo_slabs when the seam's chunk set changes. */
int dsp_set_halo_slabs(dsp_ctx* ctx, uint64_t n, const uint32_t* slots, int face, const void* src_device);
int dsp_clear_halo_slabs(dsp_ctx* ctx); /* forgets every registered slab (the buffers are kept for the next exchange) */

/* Seams inside the library: peer memory instead of pack -> copy -> register.  Each side of a seam creates an INBOX in its own
 * HBM and hands the other side a handle to it; an exchange is one small kernel per seam that writes this GPU's border
 * layers straight into the neighbour's inbox over NVLink (peer access within a process, CUDA IPC between processes) and
 * flags the epoch; the DSP_MESH_HALO call that follows meshes the interior of the region first and lets a CTA that reaches
 * a seam chunk wait on the (local) inbox flag -- the transfer hides behind the interior inside ONE mesh launch, with no
 * host synchronisation anywhere.  Nothing like this exists in the reference (single process, chunk-local cull,
 * chunk_mesh.rs:46-51); semantics are those of DSP_MESH_HALO: results equal a single context holding the whole world.
 *
 *   rank r: dsp_seam_create(face, slots)  ->  seam id + handle          (slots: r's boundary chunks at `face`, in an order
 *           hand the 128-byte handle to the neighbour by any means          both sides agree on; the neighbour lists ITS
 *           dsp_seam_connect(seam, neighbour's handle)                      boundary chunks at the opposite face, same order)
 *   per frame / epoch, on every rank:  dsp_seam_exchange_async(ctx);  dsp_mesh_*(..., DSP_MESH_HALO, ...);
 *
 * Exchanges are collective over a seam (both sides, once per epoch, before the halo-mode mesh calls of that epoch); a side
 * that never shows up turns into DSP_ERR_INTERNAL after a bounded wait (~1 s), not a hang.  Seam chunks cannot be unloaded.
 * Two contexts on the SAME device: enqueue the exchange of both before either mesh call (a spinning mesh kernel can keep a
 * later push kernel of the same GPU from being scheduled). */
typedef struct dsp_seam_handle { uint8_t bytes[128]; } dsp_seam_handle;
int dsp_seam_create(dsp_ctx* ctx, int face, uint64_t n, const uint32_t* slots, uint32_t* out_seam, dsp_seam_handle* out_handle);
int dsp_seam_connect(dsp_ctx* ctx, uint32_t seam, const dsp_seam_handle* peer_handle);
int dsp_seam_exchange_async(dsp_ctx* ctx); /* all seams of the context; asynchronous on the context's stream */
uint32_t dsp_seam_epoch(const dsp_ctx* ctx); /* number of exchanges started so far */
/* Fallback transport for a seam whose neighbour's memory cannot be mapped (no peer path between the GPUs, processes in separate
 * PID namespaces): NCCL send/recv, called from inside the library on the context's stream.  libnccl.so.2 is resolved with dlopen
 * when first needed (no link-time dependency).  Rank 0 makes an id with dsp_comm_unique_id and distributes the 128 bytes; every
 * rank calls dsp_comm_init (collective, like ncclCommInitRank); a seam is then connected with dsp_seam_connect_nccl(seam, rank of
 * the neighbour) instead of dsp_seam_connect.  dsp_seam_exchange_async and the halo-mode mesh calls are used exactly as above
 * (both routes may be mixed in one context).  The exchange is then not hidden behind the interior chunks to the same degree: the
 * receive completes in stream order before the mesh kernel starts. */
int dsp_comm_unique_id(uint8_t* out_id128);
int dsp_comm_init(dsp_ctx* ctx, int rank, int n_ranks, const uint8_t* id128);
int dsp_seam_connect_nccl(dsp_ctx* ctx, uint32_t seam, int peer_rank);

/* ---- scene: GameScene::update's visible set -> load -> mesh -> unload cycle (engine/scenes/scene_game.rs:39-82) -------- */
typedef struct dsp_scene_config {
    uint32_t horizontal_render_distance; /* "Horizonal Render Distance" of settings.json5 (engine/core/user_settings.rs:55-75) */
    uint32_t vertical_render_distance;
    int32_t gen_kind;                    /* dsp_gen_kind used by World::get_chunk; DSP_GEN_REFERENCE is the reference's rule */
    uint32_t seed;
    uint32_t mesh_mode;                  /* DSP_MESH_* (not ORDERED) */
    uint32_t reserved;
} dsp_scene_config;
typedef struct dsp_scene_stats {
    int32_t current_chunk[3]; /* (camera as i32).div_euclid(16), scene_game.rs:41-46 */
    uint32_t changed;         /* 0: the camera is still in last_chunk_pos and nothing was done (scene_game.rs:48) */
    uint64_t n_visible;       /* get_all_chunk_pos_in_render(...).len() */
    uint64_t n_loaded_now, n_unloaded_now;
    uint64_t n_loaded_total;  /* world.loaded_chunks.len() */
    uint64_t n_meshes;        /* amount_of_chunk_meshes(): chunk_meshes.values().flatten().count() */
    uint64_t vertex_total, index_total;
    uint64_t arena_bytes_used;
} dsp_scene_stats;
/* get_all_chunk_pos_in_render (scene_game.rs:170-223) for a camera position, in the reference's order (x, then y, then z
 * ascending).  Host only, no context.  out_pos (cap x 3, may be NULL to query the count). */
int dsp_visible_chunks(const float camera_pos[3], uint32_t horizontal, uint32_t vertical, uint64_t cap, int32_t* out_pos, uint64_t* out_n);
/* GameScene::update for one camera position: if the camera's chunk changed, every visible chunk that is not loaded is
 * generated and meshed -- one terrain launch and one mesh launch for the whole set -- into a free interval of the arena,
 * and every loaded chunk that left the visible set is unloaded (its voxel slot and, once its whole batch is gone, its
 * arena interval are recycled).  Blocking, like the reference's update.  DSP_ERR_ARENA_FULL / DSP_ERR_STORE_FULL leave
 * the scene as it was.  With DSP_MESH_HALO only the newly loaded chunks are meshed (against every chunk resident at that
 * moment); chunks meshed earlier keep their meshes, which at worst contain border faces that a later neighbour now hides, or
 * lack faces that only an unloaded neighbour's solid voxels used to cover -- neither is visible from inside the loaded set.
 * The scene owns every chunk of its context and the arena: a context with dsp_generate_region boxes or with chunks
 * loaded by hand is refused (DSP_ERR_BAD_ARG); do not choose arena offsets by hand on the same context.  Every error
 * leaves the scene as it was. */
int dsp_scene_update(dsp_ctx* ctx, const float camera_pos[3], const dsp_scene_config* cfg, dsp_scene_stats* out_stats);
/* GameScene::draw's iteration (scene_game.rs:101-125): position and arena window of every non-empty mesh. */
int dsp_scene_draw_list(dsp_ctx* ctx, uint64_t cap, int32_t* out_pos /* cap x 3 */, dsp_mesh_entry* out_entries, uint64_t* out_n);
int dsp_scene_reset(dsp_ctx* ctx); /* unloads everything the scene loaded */

/* ---- data access ------------------------------------------------------------------------------------ */
int dsp_download_arena(dsp_ctx* ctx, uint64_t offset_bytes, uint64_t bytes, void* host_dst);
int dsp_download_chunk(dsp_ctx* ctx, uint32_t slot, uint16_t* host_dst /* 4096 */, int32_t* out_pos /* 3, may be NULL */);
/* Raw device pointers for zero-copy consumers (CUDA/Vulkan external-memory interop, NCCL halo
 * exchange, torch tensors over the same memory).  Valid until dsp_destroy. */
void* dsp_arena_device_ptr(dsp_ctx* ctx);
uint64_t dsp_arena_bytes(const dsp_ctx* ctx);
void* dsp_voxel_device_ptr(dsp_ctx* ctx);   /* max_chunks x 4096 u16; after writing through it call dsp_invalidate_chunks */
void* dsp_entries_device_ptr(dsp_ctx* ctx); /* dsp_mesh_entry[n] of the last mesh call */
void* dsp_stream(dsp_ctx* ctx);             /* cudaStream_t all work of this context is enqueued on */
uint64_t dsp_max_chunks(const dsp_ctx* ctx);

/* ---- arena export: the vertex arena as an external-memory object (zero-copy consumers) ----------------------------------
 * The reference's mesh ends up in a Vulkan vertex buffer (chunk_mesh.rs:111-124).  A context created with
 * DSP_CREATE_EXPORTABLE_ARENA can export its arena as a POSIX file descriptor (cuMemExportToShareableHandle): Vulkan imports
 * it with VK_KHR_external_memory_fd (VkImportMemoryFdInfoKHR, VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT; the engine would
 * add the extension next to khr_swapchain, engine/rendering/vulkan.rs:34-37) and binds a VERTEX_BUFFER over it -- the
 * dsp_mesh_entry offsets index that buffer directly; another CUDA process imports it with dsp_arena_import_fd.  The fd is the
 * caller's to pass on (SCM_RIGHTS / inheritance) and to close; each call returns a new one.  *out_allocation_bytes (>=
 * dsp_arena_bytes, rounded to the allocation granularity) is the size an importer must map.  INTEGRATION.md has the
 * Vulkano-side snippet. */
int dsp_arena_export_fd(dsp_ctx* ctx, int* out_fd, uint64_t* out_allocation_bytes);
/* CUDA-side importer: maps the allocation behind `fd` on `device` of THIS process; *out_device_ptr then aliases the exporting
 * context's arena.  No context needed; errors through dsp_last_error(NULL). */
typedef struct dsp_arena_import dsp_arena_import;
int dsp_arena_import_fd(int device, int fd, uint64_t allocation_bytes, dsp_arena_import** out_import, void** out_device_ptr);
int dsp_arena_import_read(dsp_arena_import* imp, uint64_t offset_bytes, uint64_t bytes, void* host_dst); /* blocking copy to host */
int dsp_arena_import_close(dsp_arena_import* imp);

/* Number of kernel launches this context has enqueued since creation (bench accounting). */
uint64_t dsp_kernel_launches(const dsp_ctx* ctx);
/* Device time of the kernels of the mesh calls alone (the mesh kernel and, where a call has one, the classify pre-pass in
 * front of it), from CUDA events the library records around them on the context's stream: number of calls, their summed
 * duration, and the newest one.  Waits for outstanding launches. */
int dsp_mesh_kernel_stats(dsp_ctx* ctx, uint64_t* out_count, double* out_total_ms, float* out_last_ms);
/* The event pair around every mesh launch costs about 2 us of stream time per call; a caller that does not read
 * dsp_mesh_kernel_stats can switch it off (on by default). */
int dsp_set_kernel_timing(dsp_ctx* ctx, int on);

/* How the chunks of the last dsp_mesh_host_chunks / dsp_mesh_host_chunks_ex call crossed the bus (see "packed upload" at
 * dsp_mesh_host_chunks_ex): out[0..5] = chunks sent as air / uniform / 2-bit / 4-bit / 8-bit / raw, out[6] = host worker threads
 * that packed them, out[7] = bytes the device read from host memory (packed voxels + 24 bytes of position, slot and descriptor
 * per chunk).  All zero when the call took another path. */
int dsp_host_upload_stats(dsp_ctx* ctx, uint64_t out[8]);
/* Test hook, no context: the host-side packer on one chunk.  blocks: 4,096 u16; packed: 8,192 bytes of room; *desc: bits 0-2 the
 * form (0 air, 1 uniform, 2 / 3 / 4 = 2 / 4 / 8 bits per voxel, 5 raw), bits 16-31 the block id of a uniform chunk.  portable != 0
 * runs the plain C++ packer instead of the AVX2 one (they must agree). */
int dsp_debug_pack_chunk(const uint16_t* blocks, uint8_t* packed, uint32_t* desc, int portable);

/* Debug aid: per-phase SM-cycle counters of the mesh kernels (32 of them), summed over CTAs, read and cleared.  All zero
 * unless the library was built with -DDSP_PHASE_TIMERS (not the shipped build). */
int dsp_debug_phase_cycles(dsp_ctx* ctx, uint64_t* out32);

#ifdef __cplusplus
}
#endif
#endif /* DESPAWN_B200_H */
