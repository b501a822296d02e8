"""dsp_mesh_host_chunks_ex (GPU, through the C ABI): host-resident chunks whose tiles the mesh kernel pulls out of pinned memory
itself, and the reference's palette length as the all-air early-out.

Reference: World::load_chunk + build_chunk_mesh over Chunk.blocks held in host memory (/root/reference/src/content/world/
chunks/chunk.rs:17, engine/scenes/scene_game.rs:56-64); palette.len() == 1 => None without looking at a voxel
(chunks/chunk_mesh.rs:18-20).  Every path -- packed upload (the default for bulk calls; tests/test_gpu_packed_upload.py has its own
cases), kernel pull from pinned memory, copy-engine pipeline (DSP_E2E_PULL=0), pageable memory -- must give the oracle's streams byte
for byte and leave the chunks resident with the caller's voxels."""
import os

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc

pytestmark = pytest.mark.gpu

SEED_NOISE = 0xD5E50001


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def align128(n):
    return (n + 127) & ~127


def box_positions(origin, dims):
    return np.array([(origin[0] + a, origin[1] + b, origin[2] + c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)


def oracle_streams(pos, ids, uv):
    counts, blob = orc.mesh_batch(pos, ids, uv, flavour=orc.FAIR)
    blob = blob.view(np.uint8).reshape(-1)
    out, o = [], 0
    for c in counts:
        out.append(blob[o:o + int(c) * 20].tobytes())
        o += int(c) * 20
    return counts, out


def pinned_copy(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).pin_memory()
    return t, t.numpy().view(np.uint16)


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        for k, v in self.kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def world(n_extra_random=6):
    """Noise columns (air / one-solid-id / surface chunks) plus a few random-density chunks: 1,400+ chunks, i.e. several sub-batches."""
    pos = box_positions((2, 0, -3), (9, 16, 10))
    ids = orc.generate_batch(orc.GEN_NOISE, SEED_NOISE, pos)
    rng = np.random.default_rng(17)
    extra_pos = np.array([(100 + k, 3, 7) for k in range(n_extra_random)], dtype=np.int32)
    extra = np.stack([((rng.random(4096) < d) * rng.integers(1, 3, 4096)).astype(np.uint16) for d in np.linspace(0.02, 0.98, n_extra_random)])
    pos, ids = np.concatenate([pos, extra_pos]), np.concatenate([ids, extra])
    perm = rng.permutation(len(pos))  # no run structure in the slot order
    return pos[perm], ids[perm]


@pytest.mark.parametrize("with_palette_lens", [False, True])
@pytest.mark.parametrize("path", ["packed", "packed pageable", "pull", "copy engine", "pageable"])
def test_host_chunks_every_path_equals_the_oracle(uv_default, path, with_palette_lens):
    pos, ids = world()
    n = len(pos)
    counts, want = oracle_streams(pos, ids, uv_default)
    pal = np.where((ids == 0).all(axis=1), 1, 3).astype(np.uint32) if with_palette_lens else None
    keep, host = pinned_copy(ids) if "pageable" not in path else (None, ids.copy())
    # (DSP_E2E_PULL 2: pull whatever the share of all-air chunks -- the default wants a tenth; DSP_E2E_PACK 0: no packed upload)
    with env(DSP_E2E_PULL="0" if path == "copy engine" else "2", DSP_E2E_PACK="2" if path.startswith("packed") else "0"):
        with capi.Context(0, n + 64, 1 << 30) as c:
            c.set_uv_table(uv_default)
            entries = np.zeros(n, dtype=capi.MESH_ENTRY_DTYPE)
            slots = np.zeros(n, dtype=np.uint32)
            launches0 = c.kernel_launches
            total = c.mesh_host_chunks_ex_ptr(n, pos, host.ctypes.data, pal, entries, slots_out=slots)
            assert np.array_equal(entries["vertex_count"].astype(np.int64), counts)
            assert total == sum(align128(int(k) * 20) for k in counts)
            assert [c.download_vertices(e).tobytes() for e in entries] == want
            assert c.kernel_launches > launches0
            st = c.host_upload_stats()
            assert (sum(st[k] for k in capi.PACK_FORMS) == n) == path.startswith("packed"), st
            # the chunks are resident with the caller's voxels, whichever way they came
            assert c.resident_chunks == n
            for i in range(0, n, 37):
                got, p = c.download_chunk(int(slots[i]))
                assert np.array_equal(got, ids[i]) and tuple(p) == tuple(pos[i])
            # ... and behave like any other resident chunk: an edit, then a plain mesh call of everything
            k = int(np.flatnonzero((ids == 0).all(axis=1))[0])
            ed = np.zeros(1, dtype=capi.EDIT_DTYPE)
            ed[0] = (int(pos[k][0]) * 16 + 3, int(pos[k][1]) * 16 + 4, int(pos[k][2]) * 16 + 5, 2)
            c.apply_edits(ed)
            ids2 = ids.copy()
            ids2[k, 3 + 16 * 4 + 256 * 5] = 2
            e2, _ = c.mesh_chunks(slots)
            _, want2 = oracle_streams(pos, ids2, uv_default)
            assert [c.download_vertices(e).tobytes() for e in e2] == want2
    del keep


@pytest.mark.parametrize("packed", ["2", "0"])
def test_palette_length_one_means_the_voxels_are_never_fetched(uv_default, packed):
    """chunk_mesh.rs:18-20: palette.len() == 1 => None.  The slot is zero-filled on the device: whatever was resident at that
    position before is gone, and what the (unread) host array holds does not matter."""
    pos = box_positions((0, 0, 0), (40, 1, 40))  # 1,600 chunks: several sub-batches
    n = len(pos)
    rng = np.random.default_rng(3)
    ids = ((rng.random((n, 4096)) < 0.05) * 1).astype(np.uint16)
    pal = np.full(n, 2, dtype=np.uint32)
    air = rng.random(n) < 0.4
    pal[air] = 1
    keep, host = pinned_copy(ids)  # the air-flagged chunks still hold voxels here: a caller that broke the invariant
    with env(DSP_E2E_PACK=packed), capi.Context(0, n, 512 << 20) as c:
        c.set_uv_table(uv_default)
        first_slots = c.load_chunks(pos, np.full((n, 4096), 1, dtype=np.uint16))  # every position resident and solid beforehand
        entries = np.zeros(n, dtype=capi.MESH_ENTRY_DTYPE)
        slots = np.zeros(n, dtype=np.uint32)
        c.mesh_host_chunks_ex_ptr(n, pos, host.ctypes.data, pal, entries, slots_out=slots)
        assert np.array_equal(slots, first_slots)
        expect = ids.copy()
        expect[air] = 0
        counts, want = oracle_streams(pos, expect, uv_default)
        assert np.array_equal(entries["vertex_count"].astype(np.int64), counts) and (entries["vertex_count"][air] == 0).all()
        for i in list(np.flatnonzero(air)[:5]) + list(np.flatnonzero(~air)[:5]):
            assert np.array_equal(c.download_chunk(int(slots[i]))[0], expect[i])
            assert c.download_vertices(entries[i]).tobytes() == want[i]
        # cross-chunk culling over the same chunks sees the air where it was declared (border masks are rebuilt from the store)
        e_h, _ = c.mesh_chunks(slots, capi.MESH_HALO)
    with capi.Context(0, n, 512 << 20) as c2:
        c2.set_uv_table(uv_default)
        s2 = c2.load_chunks(pos, expect)
        e_ref, _ = c2.mesh_chunks(s2, capi.MESH_HALO)
        assert np.array_equal(e_h["vertex_count"], e_ref["vertex_count"])
    del keep


@pytest.mark.parametrize("n", [0, 1, 5, 129, 1300])
def test_pull_path_small_and_ragged_batches_with_vertices_to_host(uv_default, n):
    import torch
    pos_all, ids_all = world()
    pos, ids = pos_all[:n], ids_all[:n]
    keep, host = pinned_copy(ids) if n else (None, np.zeros((0, 4096), dtype=np.uint16))
    with env(DSP_E2E_PULL="2", DSP_E2E_PACK="0"), capi.Context(0, 2048, 1 << 30) as c:  # (2: the kernel pull also in front of a download, where the default prefers the copy engine)
        c.set_uv_table(uv_default)
        entries = np.zeros(max(n, 1), dtype=capi.MESH_ENTRY_DTYPE)[:n]
        mirror = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        pal = np.where((ids == 0).all(axis=1), 1, 2).astype(np.uint32)
        total = c.mesh_host_chunks_ex_ptr(n, pos, host.ctypes.data if n else 0, pal, entries, arena_offset=256, host_vertices_ptr=mirror.data_ptr(), host_capacity=256 << 20)
        counts, want = oracle_streams(pos, ids, uv_default) if n else (np.zeros(0, dtype=np.int64), [])
        assert total == sum(align128(int(k) * 20) for k in counts)
        m = mirror.numpy()
        for e, w in zip(entries, want):
            o = int(e["offset_bytes"]) - 256
            assert m[o:o + len(w)].tobytes() == w
    del keep


@pytest.mark.parametrize("packed", ["2", "0"])
def test_pull_path_rejects_duplicates_and_rolls_back(uv_default, packed):
    pos, ids = world()
    pos, ids = pos[:600].copy(), ids[:600]
    pos[577] = pos[12]
    keep, host = pinned_copy(ids)
    with env(DSP_E2E_PACK=packed), capi.Context(0, 1024, 64 << 20) as c:
        c.set_uv_table(uv_default)
        entries = np.zeros(600, dtype=capi.MESH_ENTRY_DTYPE)
        with pytest.raises(capi.DspError) as ei:
            c.mesh_host_chunks_ex_ptr(600, pos, host.ctypes.data, None, entries)
        assert ei.value.status == capi.ERR_BAD_ARG and c.resident_chunks == 0
        # arena too small: reported with the bytes needed, nothing stays resident
        pos[577] = (999, 0, 0)
        with capi.Context(0, 1024, 1 << 20) as small:
            small.set_uv_table(uv_default)
            with pytest.raises(capi.DspError) as e2:
                small.mesh_host_chunks_ex_ptr(600, pos, host.ctypes.data, None, entries)
            assert e2.value.status == capi.ERR_ARENA_FULL and small.resident_chunks == 0
        total = c.mesh_host_chunks_ex_ptr(600, pos, host.ctypes.data, None, entries)  # the context is usable afterwards
        assert total > 0 and c.resident_chunks == 600
    del keep


@pytest.mark.parametrize("mode", [capi.MESH_GREEDY | capi.MESH_INDEXED, capi.MESH_GREEDY, capi.MESH_INDEXED])
def test_quad_modes_pull_equals_copy_engine_and_the_extension_oracle(uv_default, mode):
    pos, ids = world()
    n = len(pos)
    pal = np.where((ids == 0).all(axis=1), 1, 3).astype(np.uint32)
    keep, host = pinned_copy(ids)
    got = {}
    for label, e, pk in (("pull", "2", "0"), ("copy engine", "0", "0"), ("packed", "2", "2")):
        with env(DSP_E2E_PULL=e, DSP_E2E_PACK=pk):
            with capi.Context(0, n + 64, 1 << 30) as c:
                c.set_uv_table(uv_default)
                entries = np.zeros(n, dtype=capi.MESH_ENTRY_DTYPE)
                slots = np.zeros(n, dtype=np.uint32)
                c.mesh_host_chunks_ex_ptr(n, pos, host.ctypes.data, pal, entries, mode, slots_out=slots)
                got[label] = (entries.copy(), [(c.download_vertices(x).tobytes(), c.download_indices(x).tobytes() if mode & capi.MESH_INDEXED else b"") for x in entries])
                for i in range(0, n, 53):
                    assert np.array_equal(c.download_chunk(int(slots[i]))[0], ids[i])
    assert np.array_equal(got["pull"][0]["vertex_count"], got["copy engine"][0]["vertex_count"])
    assert np.array_equal(got["pull"][0]["index_count"], got["copy engine"][0]["index_count"])
    assert got["pull"][1] == got["copy engine"][1]
    assert np.array_equal(got["pull"][0]["vertex_count"], got["packed"][0]["vertex_count"]) and got["pull"][1] == got["packed"][1]
    for i in range(0, n, 29):
        v, ix, _ = orc.mesh_chunk_ext(tuple(int(k) for k in pos[i]), ids[i], uv_default, greedy=bool(mode & capi.MESH_GREEDY), indexed=bool(mode & capi.MESH_INDEXED))
        assert got["pull"][1][i][0] == v.tobytes()
        if mode & capi.MESH_INDEXED:
            assert got["pull"][1][i][1] == ix.tobytes()
    del keep


def test_bench_region_through_the_e2e_call_at_full_size(uv_default):
    """BASELINE config 2 (the bench's workload) through the call bench.py times as `e2e`: all 4,096 host-resident chunks of the
    16x16x16-chunk noise region, pinned, with their palette lengths -- kernel pull, all-air chunks never fetched.  Checksum of
    per-chunk checksums against the oracle; every chunk resident afterwards with its voxels."""
    import hashlib
    pos = box_positions((0, 0, 0), (16, 16, 16))
    ids = orc.generate_batch(orc.GEN_NOISE, SEED_NOISE, pos)
    counts, want = orc.mesh_batch(pos, ids, uv_default, flavour=orc.FAIR)
    pal = np.array([1 + len(np.setdiff1d(np.unique(c), [0])) for c in ids], dtype=np.uint32)
    assert int((pal == 1).sum()) == 1540
    keep, host = pinned_copy(ids)
    with capi.Context(0, 4096, 1 << 30) as c:
        c.set_uv_table(uv_default)
        entries = np.zeros(4096, dtype=capi.MESH_ENTRY_DTYPE)
        slots = np.zeros(4096, dtype=np.uint32)
        for rep in range(2):  # the second call finds every position resident
            total = c.mesh_host_chunks_ex_ptr(4096, pos, host.ctypes.data, pal, entries, slots_out=slots)
            assert np.array_equal(entries["vertex_count"].astype(np.int64), counts) and total == 396_833_664
            blob = c.download_arena(0, total)
            hg, hw, w0 = hashlib.sha256(), hashlib.sha256(), 0
            wb = want.view(np.uint8).reshape(-1)
            for e, k in zip(entries, counts):
                nb, o = int(k) * 20, int(e["offset_bytes"])
                hg.update(hashlib.sha256(blob[o:o + nb].tobytes()).digest())
                hw.update(hashlib.sha256(wb[w0:w0 + nb].tobytes()).digest())
                w0 += nb
            assert hg.hexdigest() == hw.hexdigest(), rep
        for i in range(0, 4096, 97):
            assert np.array_equal(c.download_chunk(int(slots[i]))[0], ids[i])
    del keep


def test_host_mirror_builds_all_host_chunks_like_the_reference(uv_default):
    """despawnengine_b200.world: `Chunk`s filled with Chunk::set_block as the reference's generators do (chunk.rs:86-120: flat,
    full, empty) plus a few hand-edited ones, meshed in bulk through build_all_host_chunks -- against the oracle's faithful
    flavour (string palettes and all) and against chunk-at-a-time build_chunk_mesh over put_chunk."""
    import torch
    from despawnengine_b200 import world as w
    uvs = {"template:dirt": w.AtlasUV((0.0, 0.0), (0.5, 1.0)), "template:engine": w.AtlasUV((0.5, 0.0), (1.0, 1.0))}
    rng = np.random.default_rng(2)
    chunks = []
    for k in range(60):
        c = w.Chunk(position=(k % 7 - 3, k // 7 - 4, 5 - k % 3))
        kind = k % 5
        if kind == 0:      # generate_flat
            for x in range(16):
                for y in range(8):
                    for z in range(16):
                        c.set_block(x, y, z, "template:dirt")
        elif kind == 1:    # generate_full, in the other block
            for x in range(16):
                for y in range(16):
                    for z in range(16):
                        c.set_block(x, y, z, "template:engine")
        elif kind == 2:    # generate_empty: palette stays ["base:air"]
            pass
        elif kind == 3:    # sparse, engine first: the palette order differs from the registry's
            for _ in range(200):
                x, y, z = (int(v) for v in rng.integers(0, 16, 3))
                c.set_block(x, y, z, "template:engine" if rng.random() < 0.5 else "template:dirt")
        else:              # placed and removed again: all air, but the palette has grown -- no early-out, and no faces
            c.set_block(1, 2, 3, "template:dirt")
            c.set_block(1, 2, 3, w.AIR_BLOCK_ID)
        chunks.append(c)
    pinned = torch.empty((len(chunks), 4096), dtype=torch.int16).pin_memory()
    world = w.World(0, 256, 256 << 20)
    try:
        meshes = w.build_all_host_chunks(world, chunks, uvs, blocks_out=pinned.numpy().view(np.uint16))
        for c, m in zip(chunks, meshes):
            gid = np.array([world.registry.id_of(b) for b in c.palette], dtype=np.uint16)[c.blocks]
            want = orc.mesh_chunk(c.position, gid, world.registry.uv_table(uvs), flavour=orc.FAITHFUL)
            assert (m is None) == (want.shape[0] == 0), c.position
            if m is not None:
                assert m.to_host().tobytes() == want.tobytes(), c.position
        assert sum(m is None for m in meshes) == 24  # the 12 never-touched chunks and the 12 emptied ones
        world2 = w.World(0, 256, 256 << 20)
        try:
            for c, m in zip(chunks[:20], meshes[:20]):
                world2.put_chunk(c)
                one = w.build_chunk_mesh(world2, c.position, uvs)
                assert (one is None) == (m is None) and (one is None or one.to_host().tobytes() == m.to_host().tobytes())
        finally:
            world2.close()
    finally:
        world.close()
