/* Region seams through the C ABI ONLY (plain C, no Python, no torch, no NCCL): a world of `wx` x 4 x 3 chunks is cut into two
 * X-slabs held by two contexts -- on two GPUs when the box has them (peer access over NVLink), on one GPU otherwise -- the seam
 * is wired with dsp_seam_create / dsp_seam_connect, and every epoch is dsp_seam_exchange_async on both contexts followed by one
 * DSP_MESH_HALO call each.  Every chunk's vertex stream must equal, byte for byte, what a single context that holds the whole
 * world produces (border rule of chunk_mesh.rs:46-51 applied across chunk and region borders).
 * Epoch 2 edits voxels on both sides of the seam first, so stale layers would be noticed.
 *
 *   seam_two_contexts [n_epochs]     prints "ok <devices> <chunks compared> <seam faces culled>"; exit code 0 on success */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/despawn_b200.h"

#define CHECK(call)                                                                                             \
    do {                                                                                                        \
        int rc_ = (call);                                                                                       \
        if (rc_ != DSP_OK) {                                                                                    \
            fprintf(stderr, "%s -> %d (%s): %s\n", #call, rc_, dsp_status_string(rc_), dsp_last_error(last_ctx)); \
            return 1;                                                                                           \
        }                                                                                                       \
    } while (0)

enum { WX = 6, WY = 4, WZ = 3, SEED = (int)0xD5E50001u };
static dsp_ctx* last_ctx = NULL;

static int stream_of(dsp_ctx* ctx, const dsp_mesh_entry* e, uint8_t** buf, size_t* cap) {
    const size_t bytes = (size_t)e->vertex_count * DSP_VERTEX_BYTES;
    if (bytes > *cap) { *buf = (uint8_t*)realloc(*buf, bytes); *cap = bytes; }
    last_ctx = ctx;
    return bytes ? dsp_download_arena(ctx, e->offset_bytes, bytes, *buf) : DSP_OK;
}

int main(int argc, char** argv) {
    const int epochs = argc > 1 ? atoi(argv[1]) : 3;
    const float uv[3][4] = {{0, 0, 0, 0}, {0.0f, 0.0f, 0.5f, 1.0f}, {0.5f, 0.0f, 1.0f, 1.0f}};
    dsp_ctx *whole = NULL, *c[2] = {NULL, NULL};
    CHECK(dsp_create(0, 256, 256u << 20, &whole));
    CHECK(dsp_create(0, 128, 128u << 20, &c[0]));
    int devices = 2;
    if (dsp_create(1, 128, 128u << 20, &c[1]) != DSP_OK) { devices = 1; CHECK(dsp_create(0, 128, 128u << 20, &c[1])); }
    dsp_ctx* all[3] = {whole, c[0], c[1]};
    for (int k = 0; k < 3; ++k) { last_ctx = all[k]; CHECK(dsp_set_uv_table(all[k], 3, &uv[0][0])); }

    /* world origin (3, 5, -2): the noise surface (h in [16, 239]) crosses chunk Y 5..8 in places */
    const int32_t origin[3] = {3, 5, -2};
    const uint32_t wdims[3] = {WX, WY, WZ}, half[3] = {WX / 2, WY, WZ};
    const int32_t origin1[3] = {3 + WX / 2, 5, -2};
    uint32_t first_w, first[2];
    last_ctx = whole; CHECK(dsp_generate_region(whole, origin, wdims, DSP_GEN_NOISE, (uint32_t)SEED, &first_w));
    last_ctx = c[0]; CHECK(dsp_generate_region(c[0], origin, half, DSP_GEN_NOISE, (uint32_t)SEED, &first[0]));
    last_ctx = c[1]; CHECK(dsp_generate_region(c[1], origin1, half, DSP_GEN_NOISE, (uint32_t)SEED, &first[1]));

    /* the seam: context 0's x = WX/2 - 1 layer (its +X / LEFT face) against context 1's x = 0 layer (-X / RIGHT), z outer, y inner */
    uint32_t slots0[WY * WZ], slots1[WY * WZ], seam[2];
    for (int z = 0, k = 0; z < WZ; ++z)
        for (int y = 0; y < WY; ++y, ++k) {
            slots0[k] = first[0] + (WX / 2 - 1) + (WX / 2) * (y + WY * z);
            slots1[k] = first[1] + 0 + (WX / 2) * (y + WY * z);
        }
    dsp_seam_handle h[2];
    last_ctx = c[0]; CHECK(dsp_seam_create(c[0], DSP_FACE_LEFT, WY * WZ, slots0, &seam[0], &h[0]));
    last_ctx = c[1]; CHECK(dsp_seam_create(c[1], DSP_FACE_RIGHT, WY * WZ, slots1, &seam[1], &h[1]));
    last_ctx = c[0]; CHECK(dsp_seam_connect(c[0], seam[0], &h[1]));
    last_ctx = c[1]; CHECK(dsp_seam_connect(c[1], seam[1], &h[0]));

    dsp_mesh_entry ew[WX * WY * WZ], eh[2][WX / 2 * WY * WZ], ep[WX / 2 * WY * WZ];
    uint8_t *a = NULL, *b = NULL;
    size_t cap_a = 0, cap_b = 0;
    unsigned long long compared = 0, culled = 0;
    for (int epoch = 1; epoch <= epochs; ++epoch) {
        if (epoch >= 2) {  /* edits on both border layers of the seam (and one interior), applied to the whole world as well */
            const int32_t sx = (origin[0] + WX / 2) * 16;  /* first voxel column of context 1 */
            dsp_edit ed[4] = {{sx - 1, 6 * 16 + epoch, -2 * 16 + 3, 0}, {sx, 6 * 16 + epoch, -2 * 16 + 3, 2},
                              {sx - 1, 7 * 16 + 1, -1 * 16 + epoch, (uint32_t)(epoch & 1)}, {sx + 20, 6 * 16 + 2, 5, 1}};
            uint64_t nd;
            last_ctx = whole; CHECK(dsp_apply_edits(whole, 4, ed, NULL, 0, &nd));
            last_ctx = c[0]; CHECK(dsp_apply_edits(c[0], 4, ed, NULL, 0, &nd));  /* edits outside a context's region are ignored */
            last_ctx = c[1]; CHECK(dsp_apply_edits(c[1], 4, ed, NULL, 0, &nd));
        }
        /* both pushes first, then both meshes (two contexts may share one GPU here, see the header) */
        last_ctx = c[0]; CHECK(dsp_seam_exchange_async(c[0]));
        last_ctx = c[1]; CHECK(dsp_seam_exchange_async(c[1]));
        last_ctx = c[0]; CHECK(dsp_mesh_range_async(c[0], first[0], WX / 2 * WY * WZ, DSP_MESH_HALO, 0));
        last_ctx = c[1]; CHECK(dsp_mesh_range_async(c[1], first[1], WX / 2 * WY * WZ, DSP_MESH_HALO, 0));
        uint64_t total;
        last_ctx = c[0]; CHECK(dsp_sync(c[0])); CHECK(dsp_mesh_result(c[0], eh[0], &total));
        last_ctx = c[1]; CHECK(dsp_sync(c[1])); CHECK(dsp_mesh_result(c[1], eh[1], &total));
        last_ctx = whole; CHECK(dsp_mesh_range(whole, first_w, WX * WY * WZ, DSP_MESH_HALO, 0, ew, &total));
        for (int r = 0; r < 2; ++r)
            for (int z = 0; z < WZ; ++z)
                for (int y = 0; y < WY; ++y)
                    for (int x = 0; x < WX / 2; ++x) {
                        const dsp_mesh_entry* got = &eh[r][x + (WX / 2) * (y + WY * z)];
                        const dsp_mesh_entry* want = &ew[(x + r * (WX / 2)) + WX * (y + WY * z)];
                        if (got->vertex_count != want->vertex_count) {
                            fprintf(stderr, "epoch %d rank %d chunk (%d,%d,%d): %u vertices, single context has %u\n", epoch, r, x, y, z, got->vertex_count, want->vertex_count);
                            return 2;
                        }
                        CHECK(stream_of(c[r], got, &a, &cap_a));
                        CHECK(stream_of(whole, want, &b, &cap_b));
                        if (memcmp(a, b, (size_t)got->vertex_count * DSP_VERTEX_BYTES) != 0) {
                            fprintf(stderr, "epoch %d rank %d chunk (%d,%d,%d): streams differ\n", epoch, r, x, y, z);
                            return 2;
                        }
                        ++compared;
                    }
        /* the seam really culls: chunk-local meshing of context 0 emits more faces than halo mode with the neighbour's layers */
        last_ctx = c[0]; CHECK(dsp_mesh_range(c[0], first[0], WX / 2 * WY * WZ, DSP_MESH_PARITY, 64u << 20, ep, &total));
        for (int k = 0; k < WY * WZ; ++k) culled += (ep[slots0[k] - first[0]].vertex_count - eh[0][slots0[k] - first[0]].vertex_count) / 6;
    }
    if (dsp_seam_epoch(c[0]) != (uint32_t)epochs || culled == 0) { fprintf(stderr, "epoch counter / culling check failed\n"); return 3; }
    printf("ok %d %llu %llu\n", devices, compared, culled);
    free(a); free(b);
    dsp_destroy(c[0]); dsp_destroy(c[1]); dsp_destroy(whole);
    return 0;
}
