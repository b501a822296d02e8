"""Parity at the FULL sizes of BASELINE.json configs 3, 4 and 5 (the sizes bench.py reports): every chunk of the 4,096-chunk
worst-case regions, every dirty chunk of a 10,000-edit frame over the 65,536-chunk world, and a sample of the 8,388,608-chunk
world (halo mode, neighbours generated on the CPU).  Byte compares against the oracle throughout."""
import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from despawnengine_b200 import synthetic as syn
from oracle import loader as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def align128(n):
    return (n + 127) & ~127


@pytest.mark.parametrize("kind,seed,faces_per_chunk", [(capi.GEN_CHECKER, 0, 12288), (capi.GEN_RANDOM, syn.SEED_RANDOM, None)])
def test_config3_every_chunk_of_the_full_region(uv_default, kind, seed, faces_per_chunk):
    """One launch over all 4,096 chunks (the measured configuration), then EVERY chunk's stream against the oracle."""
    n, dims = 4096, (16, 16, 16)
    pos = syn.box_positions((0, 0, 0), dims)
    with capi.Context(0, n, 6 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        first = ctx.generate_region((0, 0, 0), dims, kind, seed)
        entries, total = ctx.mesh_range(first, n)
        sizes = np.array([align128(int(v) * 20) for v in entries["vertex_count"]], dtype=np.int64)
        assert total == int(sizes.sum())
        if faces_per_chunk:
            assert (entries["vertex_count"] == faces_per_chunk * 6).all()
        order = np.argsort(entries["offset_bytes"], kind="stable")  # slots are gap-free, in completion order
        offs = entries["offset_bytes"].astype(np.int64)
        assert offs[order[0]] == 0 and np.array_equal(offs[order][1:], (offs[order] + sizes[order])[:-1])
        group = 256
        for g0 in range(0, n, group):
            sel = order[g0:g0 + group]                      # chunks whose slots are adjacent in the arena
            lo, hi = int(offs[sel[0]]), int(offs[sel[-1]] + sizes[sel[-1]])
            blob = ctx.download_arena(lo, hi - lo)
            ids = orc.generate_batch(kind, seed, pos[sel])
            for j, i in enumerate(sel[::37]):               # the generator, spot-checked on the device store
                assert np.array_equal(ctx.download_chunk(first + int(i))[0], ids[j * 37])
            counts, want = orc.mesh_batch(pos[sel], ids, uv_default, flavour=orc.FAIR)
            assert np.array_equal(entries["vertex_count"][sel].astype(np.int64), counts)
            wb, w0 = want.view(np.uint8).reshape(-1), 0
            for i, c in zip(sel, counts):
                nb, o = int(c) * 20, int(offs[i]) - lo
                assert np.array_equal(blob[o:o + nb], wb[w0:w0 + nb]), ("chunk", int(i))
                w0 += nb


def test_config4_one_full_frame_every_dirty_chunk(uv_default):
    """64x64x16-chunk world (65,536 chunks, 512 MiB of voxels), one frame of 10,000 SplitMix64 edits (seed 0xD5E50004) through
    dsp_apply_edits_and_remesh: the dirty list, every dirty chunk's voxels and every dirty chunk's mesh."""
    dims = (64, 16, 64)
    n = dims[0] * dims[1] * dims[2]
    with capi.Context(0, n, 3 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, syn.SEED_NOISE)
        state, edits = syn.edit_frame(syn.SEED_EDITS, 10000, dims)
        dirty, entries, total = ctx.apply_edits_and_remesh(edits)
        cx, cy, cz = edits["wx"] >> 4, edits["wy"] >> 4, edits["wz"] >> 4   # world.rs:69-74: div_euclid(16) ...
        slot_of_edit = first + cx + dims[0] * (cy + dims[1] * cz)
        vox_of_edit = (edits["wx"] & 15) + 16 * (edits["wy"] & 15) + 256 * (edits["wz"] & 15)  # ... rem_euclid(16), chunk.rs:44-46
        want_dirty = np.unique(slot_of_edit)
        assert np.array_equal(dirty.astype(np.int64), want_dirty) and len(dirty) > 9000
        local = dirty.astype(np.int64) - first
        pos = np.stack([local % dims[0], (local // dims[0]) % dims[1], local // (dims[0] * dims[1])], axis=1).astype(np.int32)
        ids = orc.generate_batch(orc.GEN_NOISE, syn.SEED_NOISE, pos)
        row_of_slot = {int(s): k for k, s in enumerate(dirty)}
        rows = np.array([row_of_slot[int(s)] for s in slot_of_edit])
        ids[rows, vox_of_edit] = edits["block"].astype(np.uint16)  # numpy assigns repeated indices in order: the last edit of a voxel wins
        counts, want = orc.mesh_batch(pos, ids, uv_default, flavour=orc.FAIR)
        assert np.array_equal(entries["vertex_count"].astype(np.int64), counts)
        blob = ctx.download_arena(0, total)
        wb, w0 = want.view(np.uint8).reshape(-1), 0
        for k, (e, c) in enumerate(zip(entries, counts)):
            nb, o = int(c) * 20, int(e["offset_bytes"])
            assert np.array_equal(blob[o:o + nb], wb[w0:w0 + nb]), ("dirty chunk", tuple(pos[k]))
            w0 += nb
        for k in range(0, len(dirty), 97):  # voxels of the device store after the frame
            assert np.array_equal(ctx.download_chunk(int(dirty[k]))[0], ids[k]), tuple(pos[k])
        # a second frame on top (edits accumulate), sampled
        state, edits2 = syn.edit_frame(state, 10000, dims)
        dirty2, entries2, _ = ctx.apply_edits_and_remesh(edits2)
        both = np.concatenate([edits, edits2])
        for k in np.linspace(0, len(dirty2) - 1, 48).astype(np.int64):
            l = int(dirty2[k]) - first
            p = (l % dims[0], (l // dims[0]) % dims[1], l // (dims[0] * dims[1]))
            v = syn.replay_edits_on_chunk(orc.generate(orc.GEN_NOISE, syn.SEED_NOISE, p), p, both)
            assert ctx.download_vertices(entries2[k]).tobytes() == orc.mesh_chunk(p, v, uv_default).tobytes(), p


def test_config5_full_world_sample_in_halo_and_parity_mode(uv_default):
    """The 512x512x32-chunk world (8,388,608 chunks, 64 GiB of voxels) on one GPU: chunks spread over the world -- surface
    band, world borders, buried and air chunks -- in halo mode against the extended oracle with CPU-generated neighbours,
    and in the reference's chunk-local mode."""
    dims = (512, 32, 512)
    n = dims[0] * dims[1] * dims[2]
    dirs = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
    with capi.Context(0, n, 2 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, syn.SEED_NOISE)
        rng = np.random.default_rng(5)
        sample = [(int(x), int(y), int(z)) for x, y, z in zip(rng.integers(0, 512, 40), rng.integers(1, 15, 40), rng.integers(0, 512, 40))]
        sample += [(0, 7, 0), (511, 8, 511), (0, 6, 300), (300, 9, 511), (255, 0, 255), (256, 31, 256), (17, 15, 400)]
        slots = np.array([first + x + dims[0] * (y + dims[1] * z) for x, y, z in sample], dtype=np.uint32)
        e_h, _ = ctx.mesh_chunks(slots, capi.MESH_HALO)
        e_p, _ = ctx.mesh_chunks(slots, capi.MESH_PARITY, arena_offset=1 << 30)
        for p, eh, ep in zip(sample, e_h, e_p):
            ids = orc.generate(orc.GEN_NOISE, syn.SEED_NOISE, p)
            slabs, present = np.zeros((6, 256), dtype=np.uint16), np.zeros(6, dtype=np.uint8)
            for f, d in enumerate(dirs):
                q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                if all(0 <= q[i] < dims[i] for i in range(3)):
                    v = orc.generate(orc.GEN_NOISE, syn.SEED_NOISE, q).reshape(16, 16, 16)
                    slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
                    present[f] = 1
            assert ctx.download_vertices(eh).tobytes() == orc.mesh_chunk_halo(p, ids, slabs, present, uv_default).tobytes(), ("halo", p)
            assert ctx.download_vertices(ep).tobytes() == orc.mesh_chunk(p, ids, uv_default).tobytes(), ("parity", p)
        # the whole world in halo mode, in launches of 262,144 chunks: totals are consistent and most chunks need no mesh
        faces = 0
        for b0 in range(0, n, 262144):
            ent, tot = ctx.mesh_range(first + b0, 262144, capi.MESH_HALO)
            assert tot == int(sum(align128(int(v) * 20) for v in ent["vertex_count"][ent["vertex_count"] > 0]))
            faces += int(ent["vertex_count"].astype(np.int64).sum()) // 6
        assert 300_000_000 < faces < 600_000_000  # 50.2 GB of vertices in round 1's measurement = 418 M faces
