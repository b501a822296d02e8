"""Packed upload of host-resident chunks (GPU, through the C ABI): dsp_mesh_host_chunks / dsp_mesh_host_chunks_ex pack every chunk
on the host into the smallest lossless form (csrc/host_pack.hpp), unpack_host_chunks_kernel expands it into the voxel store, the
ordinary mesh kernel follows.  Reference: World::load_chunk + build_chunk_mesh over Chunk.blocks in host memory
(/root/reference/src/content/world/chunks/chunk.rs:17, engine/scenes/scene_game.rs:56-64).  Bar: the voxel store holds the caller's
voxels bit for bit whatever the form, and the streams are the oracle's."""
import os

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc
from test_gpu_host_pull import align128, env, oracle_streams, pinned_copy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def every_form_world(n=700, seed=11):
    """Chunks of every packed form in no particular order: air, one id (small and beyond 8 bits), ids < 4, < 16, < 256, anything."""
    rng = np.random.default_rng(seed)
    ids = np.zeros((n, 4096), dtype=np.uint16)
    kinds = rng.integers(0, 7, n)
    for i, k in enumerate(kinds):
        if k == 1:
            ids[i] = rng.choice([1, 2, 9, 300, 65535])
        elif k >= 2:
            top = (3, 15, 255, 2000, 3)[k - 2]
            dens = rng.choice([0.03, 0.5, 0.97])
            ids[i] = (rng.random(4096) < dens) * rng.integers(1, top + 1, 4096)
    pos = np.stack([rng.permutation(n) % 40 - 20, np.arange(n) // 40, rng.permutation(n) // 40], axis=1).astype(np.int32)
    pos[:, 0] = np.arange(n) % 40 - 20  # unique (x, y) pairs; z random
    return pos, ids, kinds


def uv_table(rows=12):
    rng = np.random.default_rng(1)
    t = rng.random((rows, 4)).astype(np.float32)
    t[0] = 0
    return t  # ids >= rows take the reference's fallback row (chunk_mesh.rs:53-56)


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("with_lens", [False, True])
def test_every_form_lands_bit_exact_and_meshes_like_the_oracle(pinned, with_lens):
    pos, ids, kinds = every_form_world()
    n = len(pos)
    uv = uv_table()
    counts, want = oracle_streams(pos, ids, uv)
    pal = np.where((ids == 0).all(axis=1), 1, 5).astype(np.uint32) if with_lens else None
    keep, host = pinned_copy(ids) if pinned else (None, ids.copy())
    with env(DSP_E2E_PACK="2"), capi.Context(0, n + 8, 2 << 30) as c:  # (2: also on a box with fewer than seven CPUs)
        c.set_uv_table(uv)
        entries = np.zeros(n, dtype=capi.MESH_ENTRY_DTYPE)
        slots = np.zeros(n, dtype=np.uint32)
        total = c.mesh_host_chunks_ex_ptr(n, pos, host.ctypes.data, pal, entries, slots_out=slots)
        st = c.host_upload_stats()
        assert sum(st[k] for k in capi.PACK_FORMS) == n and all(st[k] > 0 for k in capi.PACK_FORMS), st
        assert st["air"] == int((ids == 0).all(axis=1).sum()) and st["host_threads"] >= 1
        assert st["device_read_bytes"] == 24 * n + 1024 * st["bits2"] + 2048 * st["bits4"] + 4096 * st["bits8"] + 8192 * st["raw"]
        assert np.array_equal(entries["vertex_count"].astype(np.int64), counts)
        assert total == sum(align128(int(k) * 20) for k in counts)
        for i in range(n):
            got, p = c.download_chunk(int(slots[i]))
            assert np.array_equal(got, ids[i]) and tuple(p) == tuple(pos[i]), (i, int(kinds[i]))
        assert [c.download_vertices(e).tobytes() for e in entries] == want
        # the summaries the unpack kernel left behave like a generator's: a second, plain mesh call gives the same streams, in
        # every mode that consults them
        e2, _ = c.mesh_chunks(slots)
        assert [c.download_vertices(e).tobytes() for e in e2] == want
        e3, _ = c.mesh_chunks(slots, capi.MESH_ORDERED)
        assert [c.download_vertices(e).tobytes() for e in e3] == want
        # ... and an edit inside a chunk that arrived as "one id" un-uniforms it
        k = int(np.flatnonzero(kinds == 1)[0])
        ed = np.zeros(1, dtype=capi.EDIT_DTYPE)
        ed[0] = (int(pos[k][0]) * 16 + 5, int(pos[k][1]) * 16 + 6, int(pos[k][2]) * 16 + 7, 0)
        c.apply_edits(ed)
        ids2 = ids.copy()
        ids2[k, 5 + 16 * 6 + 256 * 7] = 0
        e4, _ = c.mesh_chunks(slots[k:k + 1])
        assert c.download_vertices(e4[0]).tobytes() == oracle_streams(pos[k:k + 1], ids2[k:k + 1], uv)[1][0]
    del keep


@pytest.mark.parametrize("n", [256, 257, 271, 1000])
@pytest.mark.parametrize("threads", ["1", "3", None])
def test_ragged_sizes_thread_counts_and_recycled_slots(uv_default, n, threads):
    pos, ids, _ = every_form_world(1000, seed=n)
    ids = np.minimum(ids, 2)  # (ids of the default uv table)
    pos, ids = pos[:n], ids[:n]
    with env(DSP_HOST_THREADS=threads, DSP_E2E_PACK="2", DSP_E2E_PACK_PIECES=str(1 + n % 4)), capi.Context(0, n + 300, 1 << 30) as c:
        c.set_uv_table(uv_default)
        # scatter the free slots: load something, unload every other chunk of it -- the call's slots are not consecutive
        pre = np.stack([np.arange(300) + 1000, np.zeros(300), np.zeros(300)], axis=1).astype(np.int32)
        pre_slots = c.load_chunks(pre, np.zeros((300, 4096), dtype=np.uint16))
        c.unload_chunks(pre_slots[::2])
        entries = np.zeros(n, dtype=capi.MESH_ENTRY_DTYPE)
        slots = np.zeros(n, dtype=np.uint32)
        c.mesh_host_chunks_ex_ptr(n, pos, ids.ctypes.data, None, entries, slots_out=slots)
        assert sum(c.host_upload_stats()[k] for k in capi.PACK_FORMS) == n
        assert len(set(slots.tolist())) == n and not set(slots.tolist()) & set(pre_slots[1::2].tolist())
        counts, want = oracle_streams(pos, ids, uv_default)
        assert np.array_equal(entries["vertex_count"].astype(np.int64), counts)
        assert [c.download_vertices(e).tobytes() for e in entries] == want
        for i in range(0, n, 7):
            assert np.array_equal(c.download_chunk(int(slots[i]))[0], ids[i])
        # the same positions again, other voxels: same slots, new contents
        ids_b = np.roll(ids, 1, axis=0)
        c.mesh_host_chunks_ex_ptr(n, pos, ids_b.ctypes.data, None, entries, slots_out=slots)
        assert [c.download_vertices(e).tobytes() for e in entries] == oracle_streams(pos, ids_b, uv_default)[1]


def test_small_calls_readback_calls_and_the_switch_take_the_other_paths(uv_default):
    import torch
    pos, ids, _ = every_form_world(400)
    ids = np.minimum(ids, 2)
    keep, host = pinned_copy(ids)
    with env(DSP_E2E_PACK="2"), capi.Context(0, 512, 1 << 30) as c:
        c.set_uv_table(uv_default)
        entries = np.zeros(400, dtype=capi.MESH_ENTRY_DTYPE)
        c.mesh_host_chunks_ex_ptr(255, pos, host.ctypes.data, None, entries[:255])
        assert sum(c.host_upload_stats().values()) == 0  # below 256 chunks the workers' wake-up costs more than the bus
        c.mesh_host_chunks_ex_ptr(400, pos, host.ctypes.data, None, entries)
        assert c.host_upload_stats()["host_threads"] >= 1
        # vertices to the host as well: packed upload + device-driven download when the destination is pinned ...
        counts, want = oracle_streams(pos, ids, uv_default)
        mirror = torch.empty(192 << 20, dtype=torch.uint8).pin_memory()
        total = c.mesh_host_chunks_ex_ptr(400, pos, host.ctypes.data, None, entries, arena_offset=1280, host_vertices_ptr=mirror.data_ptr(), host_capacity=192 << 20)
        assert sum(c.host_upload_stats()[k] for k in capi.PACK_FORMS) == 400 and total == sum(align128(int(k) * 20) for k in counts)
        m = mirror.numpy()
        for e, w in zip(entries, want):
            o = int(e["offset_bytes"]) - 1280
            assert m[o:o + len(w)].tobytes() == w
        with pytest.raises(capi.DspError) as ei:  # too small a destination is reported, as on the other paths
            c.mesh_host_chunks_ex_ptr(400, pos, host.ctypes.data, None, entries, host_vertices_ptr=mirror.data_ptr(), host_capacity=1 << 20)
        assert ei.value.status == capi.ERR_BAD_ARG
        # ... and the copy-engine pipeline when it is pageable
        pageable = np.zeros(192 << 20, dtype=np.uint8)
        c.mesh_host_chunks_ex_ptr(400, pos, host.ctypes.data, None, entries, host_vertices_ptr=pageable.ctypes.data, host_capacity=192 << 20)
        assert sum(c.host_upload_stats().values()) == 0
        for e, w in list(zip(entries, want))[::9]:
            o = int(e["offset_bytes"])
            assert pageable[o:o + len(w)].tobytes() == w
    with env(DSP_E2E_PACK="0"), capi.Context(0, 512, 1 << 30) as c:
        c.set_uv_table(uv_default)
        c.mesh_host_chunks_ex_ptr(400, pos, host.ctypes.data, None, entries)
        assert sum(c.host_upload_stats().values()) == 0
    # few worker threads and a device-addressable array: the chunks are shared out between the cores and the bus (half of them are
    # fetched as they are); same result
    counts, want = oracle_streams(pos, ids, uv_default)
    for threads, raw_min in (("3", 150), ("7", 60)):
        with env(DSP_E2E_PACK="1", DSP_HOST_THREADS=threads), capi.Context(0, 512, 1 << 30) as c:
            c.set_uv_table(uv_default)
            slots = np.zeros(400, dtype=np.uint32)
            c.mesh_host_chunks_ex_ptr(400, pos, host.ctypes.data, None, entries, slots_out=slots)
            st = c.host_upload_stats()
            assert sum(st[k] for k in capi.PACK_FORMS) == 400 and st["raw"] >= raw_min, st
            assert [c.download_vertices(e).tobytes() for e in entries] == want
            for i in range(400):
                assert np.array_equal(c.download_chunk(int(slots[i]))[0], ids[i])
    del keep
