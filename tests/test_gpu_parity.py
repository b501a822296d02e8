"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bit-exact (byte compare) everywhere -- the vertex stream is f32 but every operation is
a single IEEE rounding, so there is no tolerance to state.  Run with `-m gpu` on a B200."""
import hashlib
import os

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc

pytestmark = pytest.mark.gpu

SEED_NOISE, SEED_RANDOM = 0xD5E50001, 0xD5E50003


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


@pytest.fixture()
def ctx(uv_default):
    with capi.Context(0, 8192, 1 << 30) as c:
        c.set_uv_table(uv_default)
        yield c


def align128(n):
    return (n + 127) & ~127


def check_entries_layout(entries, total, base=0, ordered=False):
    """Slots are 128-byte aligned, gap-free and disjoint inside [base, base + total).  With DSP_MESH_ORDERED they
    also follow each other in call order; otherwise their order is the kernel's completion order."""
    sizes = np.array([align128(int(e["vertex_count"]) * 20) for e in entries], dtype=np.int64)
    offs = entries["offset_bytes"].astype(np.int64)
    assert total == int(sizes.sum())
    assert (offs % 128 == 0).all()
    if ordered:
        assert np.array_equal(offs, base + np.concatenate([[0], np.cumsum(sizes)[:-1]]))
        return
    nz = sizes > 0
    order = np.argsort(offs[nz], kind="stable")
    o, z = offs[nz][order], sizes[nz][order]
    if len(o):
        assert o[0] == base and np.array_equal(o[1:], o[:-1] + z[:-1]) and o[-1] + z[-1] == base + total
    assert ((offs[~nz] >= base) & (offs[~nz] <= base + total)).all()


def compare_batch(ctx, pos, slots, entries, uv, kind=None, seed=None, ids=None):
    """Byte-compares every chunk of a batch with the oracle."""
    pos = np.asarray(pos, dtype=np.int32).reshape(-1, 3)
    if ids is None:
        ids = orc.generate_batch(kind, seed, pos)
    counts, want = orc.mesh_batch(pos, ids, uv, flavour=orc.FAIR)
    assert np.array_equal(entries["vertex_count"].astype(np.int64), counts)
    lo = int(entries["offset_bytes"].min())
    hi = int((entries["offset_bytes"] + entries["vertex_count"].astype(np.uint64) * 20).max())
    blob = ctx.download_arena(lo, hi - lo)
    wb = want.view(np.uint8).reshape(-1)
    w0 = 0
    for e, c in zip(entries, counts):
        n = int(c) * 20
        o = int(e["offset_bytes"]) - lo
        assert np.array_equal(blob[o:o + n], wb[w0:w0 + n]), "vertex stream differs"
        w0 += n
    return ids


# ---- config 1: single default-size chunk of the reference terrain -------------------------------------
@pytest.mark.parametrize("pos,faces", [((0, 0, 0), 1024), ((0, -1, 0), 1536), ((0, 1, 0), 0), ((-3, 0, 11), 1024), ((2, -7, -5), 1536)])
def test_reference_terrain_single_chunk(ctx, uv_default, pos, faces):
    slots = ctx.generate_chunks([pos], capi.GEN_REFERENCE)
    ids, p = ctx.download_chunk(int(slots[0]))
    assert tuple(p) == pos and np.array_equal(ids, orc.generate(orc.GEN_REFERENCE, 0, pos))
    entries, total = ctx.mesh_chunks(slots)
    assert int(entries[0]["vertex_count"]) == faces * 6 and total == align128(faces * 120)
    got = ctx.download_vertices(entries[0])
    assert got.tobytes() == orc.mesh_chunk(pos, ids, uv_default).tobytes()


def test_flat_chunk_matches_committed_golden(ctx):
    slots = ctx.generate_chunks([(0, 0, 0)], capi.GEN_REFERENCE)
    entries, _ = ctx.mesh_chunks(slots)
    golden = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", "flat_chunk_origin.bin"), dtype=np.uint8)
    assert ctx.download_vertices(entries[0]).tobytes() == golden.tobytes()


def test_gpu_streams_equal_the_fixtures_made_from_the_reference_sources(ctx):
    """tests/golden/refsrc_*: 38 chunks x 4 uv tables whose streams were produced from the reference's parsed Rust sources
    (tests/golden/make_golden_from_reference.py) -- no oracle in between.  The CUDA path must hit every digest."""
    import json
    g = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(g, "refsrc_cases.npz"))
    d = json.load(open(os.path.join(g, "refsrc_digests.json")))
    slots = ctx.load_chunks(z["pos"], z["ids"])
    for k, tname in enumerate(d["uv_tables"]):
        ctx.set_uv_table(z["uv"][k, :int(z["uv_rows"][k])])
        for mode in (capi.MESH_PARITY, capi.MESH_ORDERED):
            entries, _ = ctx.mesh_chunks(slots, mode)
            for case, e in zip(d["cases"], entries):
                want = case["streams"][tname]
                got = ctx.download_vertices(e)
                assert int(e["vertex_count"]) == want["vertices"] and hashlib.sha256(got.tobytes()).hexdigest() == want["sha256"], (case["name"], tname, mode)


def test_uv_variants(ctx, uv_default):
    """Swapped atlas assignment, missing block -> (0,0)-(1,1) fallback, non-representable thirds."""
    pos = [(1, 2, 3), (0, 0, 0)]
    slots = ctx.generate_chunks(pos, capi.GEN_RANDOM, 77)
    third = np.float32(1) / np.float32(3)
    tables = [uv_default[[0, 2, 1]], uv_default[:2],
              np.array([[0, 0, 0, 0], [third, 0, third + third, 1], [np.float32(2) * third, third, 1, np.float32(0.7)]], dtype=np.float32)]
    for uv in tables:
        ctx.set_uv_table(uv)
        entries, _ = ctx.mesh_chunks(slots)
        compare_batch(ctx, pos, slots, entries, uv, orc.GEN_RANDOM, 77)


# ---- generators ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,seed", [(capi.GEN_REFERENCE, 0), (capi.GEN_NOISE, SEED_NOISE), (capi.GEN_RANDOM, SEED_RANDOM), (capi.GEN_CHECKER, 0)])
def test_terrain_fill_matches_oracle(ctx, kind, seed):
    rng = np.random.default_rng(5)
    pos = np.concatenate([rng.integers(-40, 40, size=(60, 3)), [[0, 0, 0], [0, 1, 0], [0, -1, 0], [-1, 7, -1]]]).astype(np.int32)
    pos = np.unique(pos, axis=0)
    slots = ctx.generate_chunks(pos, kind, seed)
    want = orc.generate_batch(kind, seed, pos)
    for i, s in enumerate(slots):
        ids, p = ctx.download_chunk(int(s))
        assert np.array_equal(p, pos[i]) and np.array_equal(ids, want[i]), pos[i]


def test_region_generation_addresses_chunks_like_the_header_says(ctx):
    first = ctx.generate_region((-2, 5, 3), (3, 4, 2), capi.GEN_NOISE, SEED_NOISE)
    for (a, b, c) in [(0, 0, 0), (2, 3, 1), (1, 2, 0)]:
        slot = first + a + 3 * (b + 4 * c)
        ids, p = ctx.download_chunk(slot)
        assert tuple(p) == (-2 + a, 5 + b, 3 + c)
        assert np.array_equal(ids, orc.generate(orc.GEN_NOISE, SEED_NOISE, tuple(p)))
        assert ctx.find_chunk(tuple(p)) == slot
    assert ctx.find_chunk((9, 9, 9)) is None and ctx.resident_chunks == 24


# ---- batches: configs 2 and 3 at sizes the oracle finishes in seconds ---------------------------------------
def test_noise_region_batch(ctx, uv_default):
    dims = (6, 16, 5)  # full column height so that empty, surface and buried chunks all occur
    first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, SEED_NOISE)
    n = dims[0] * dims[1] * dims[2]
    entries, total = ctx.mesh_range(first, n)
    check_entries_layout(entries, total)
    pos = np.array([(a, b, c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)
    compare_batch(ctx, pos, None, entries, uv_default, orc.GEN_NOISE, SEED_NOISE)
    vc = entries["vertex_count"]
    assert (vc == 0).any() and (vc == 1536 * 6).any() and ((vc > 0) & (vc != 1536 * 6)).any()


@pytest.mark.parametrize("kind,seed,n", [(capi.GEN_RANDOM, SEED_RANDOM, 96), (capi.GEN_CHECKER, 0, 40)])
def test_dense_worst_case_batches(ctx, uv_default, kind, seed, n):
    rng = np.random.default_rng(11)
    pos = np.unique(rng.integers(-8, 8, size=(n * 2, 3)), axis=0)[:n].astype(np.int32)
    slots = ctx.generate_chunks(pos, kind, seed)
    entries, total = ctx.mesh_chunks(slots)
    check_entries_layout(entries, total)
    compare_batch(ctx, pos, slots, entries, uv_default, kind, seed)
    if kind == capi.GEN_CHECKER:
        assert (entries["vertex_count"] == 12288 * 6).all()  # the per-chunk maximum


def test_mixed_order_and_repeats_do_not_change_streams(ctx, uv_default):
    pos = np.array([(x, y, 0) for x in range(4) for y in range(5, 11)], dtype=np.int32)
    slots = ctx.generate_chunks(pos, capi.GEN_NOISE, SEED_NOISE)
    e1, _ = ctx.mesh_chunks(slots)
    streams = [ctx.download_vertices(e).tobytes() for e in e1]
    perm = np.random.default_rng(3).permutation(len(slots))
    e2, t2 = ctx.mesh_chunks(slots[perm], arena_offset=1 << 20)
    check_entries_layout(e2, t2, base=1 << 20)
    for j, i in enumerate(perm):
        assert ctx.download_vertices(e2[j]).tobytes() == streams[i]
    e3, _ = ctx.mesh_chunks(slots)  # idempotent
    assert [ctx.download_vertices(e).tobytes() for e in e3] == streams


@pytest.mark.parametrize("mode", [capi.MESH_ORDERED, capi.MESH_ORDERED | capi.MESH_HALO])
def test_ordered_mode_packs_slots_in_call_order(ctx, uv_default, mode):
    """DSP_MESH_ORDERED: same streams, slots back to back in the order of the call, bit-reproducible arena."""
    dims = (5, 16, 4)
    first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, SEED_NOISE)
    n = dims[0] * dims[1] * dims[2]
    e_ord, t_ord = ctx.mesh_range(first, n, mode)
    check_entries_layout(e_ord, t_ord, ordered=True)
    blob1 = ctx.download_arena(0, t_ord).copy()
    e_def, t_def = ctx.mesh_range(first, n, mode & ~capi.MESH_ORDERED, arena_offset=align128(t_ord))
    check_entries_layout(e_def, t_def, base=align128(t_ord))
    assert t_def == t_ord and np.array_equal(e_def["vertex_count"], e_ord["vertex_count"])
    for a, b in zip(e_ord[::7], e_def[::7]):
        assert ctx.download_vertices(a).tobytes() == ctx.download_vertices(b).tobytes()
    e_again, _ = ctx.mesh_range(first, n, mode)
    assert np.array_equal(e_again, e_ord) and np.array_equal(ctx.download_arena(0, t_ord), blob1)
    if not (mode & capi.MESH_HALO):
        pos = np.array([(a, b, c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)
        compare_batch(ctx, pos, None, e_ord, uv_default, orc.GEN_NOISE, SEED_NOISE)


# ---- host chunk data with per-chunk palettes (chunk.rs:58-65) ---------------------------------------------
def test_load_chunks_with_palettes(ctx, uv_default):
    rng = np.random.default_rng(9)
    n = 20
    pos = np.array([(i, -i, 2 * i) for i in range(n)], dtype=np.int32)
    gids = (rng.integers(0, 3, size=(n, 4096)) * (rng.random((n, 4096)) < 0.4)).astype(np.uint16)
    blocks = np.zeros_like(gids)
    pal, off = [], [0]
    for i in range(n):
        order = [0] + list(rng.permutation([1, 2]))  # palette order differs per chunk (first-seen order in the reference)
        inv = {g: k for k, g in enumerate(order)}
        blocks[i] = np.vectorize(inv.get)(gids[i])
        pal += order
        off.append(len(pal))
    slots = ctx.load_chunks(pos, blocks, np.array(pal, dtype=np.uint16), np.array(off, dtype=np.uint32))
    for i, s in enumerate(slots):
        assert np.array_equal(ctx.download_chunk(int(s))[0], gids[i])
    entries, _ = ctx.mesh_chunks(slots)
    compare_batch(ctx, pos, slots, entries, uv_default, ids=gids)
    # global ids straight through, including ids unknown to the uv table (fallback row)
    gids2 = gids.copy()
    gids2[gids2 == 2] = 9
    slots2 = ctx.load_chunks(pos + 100, gids2)
    entries2, _ = ctx.mesh_chunks(slots2)
    compare_batch(ctx, pos + 100, slots2, entries2, uv_default, ids=gids2)


def test_mesh_host_chunks_pipelined(ctx, uv_default):
    """dsp_mesh_host_chunks: host voxels in, sub-batched copy/mesh pipeline, same streams as load + mesh."""
    dims = (16, 10, 10)  # 1,600 chunks -> 4 sub-batches of 512
    pos = np.array([(a, 3 + b, c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_NOISE, SEED_NOISE, pos)
    slots, entries, total = ctx.mesh_host_chunks(pos, ids)
    check_entries_layout(entries, total)
    compare_batch(ctx, pos, slots, entries, uv_default, ids=ids)
    assert ctx.resident_chunks == len(pos) and np.array_equal(ctx.download_chunk(int(slots[777]))[0], ids[777])
    # again into a different part of the arena, after fragmenting the slot space so that slot runs are short
    ctx.unload_chunks(slots[100:1500:3])
    ids2 = ids[::-1].copy()
    slots2, entries2, total2 = ctx.mesh_host_chunks(pos, ids2, arena_offset=align128(total))
    check_entries_layout(entries2, total2, base=align128(total))
    compare_batch(ctx, pos, slots2, entries2, uv_default, ids=ids2)
    # ordered mode goes through the plain load + mesh path
    slots3, entries3, total3 = ctx.mesh_host_chunks(pos[:700], ids[:700], capi.MESH_ORDERED)
    check_entries_layout(entries3, total3, ordered=True)
    compare_batch(ctx, pos[:700], slots3, entries3, uv_default, ids=ids[:700])


# ---- edge cases ------------------------------------------------------------------------------------------------
def test_empty_and_single_voxel_chunks(ctx, uv_default):
    ids = np.zeros((4, 4096), dtype=np.uint16)
    ids[1, 0] = 1                    # corner voxel: 6 faces
    ids[2, 4095] = 2                 # opposite corner
    ids[3, 8 + 16 * 8 + 256 * 8] = 1  # interior voxel
    pos = np.array([(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0)], dtype=np.int32)
    slots = ctx.load_chunks(pos, ids)
    entries, total = ctx.mesh_chunks(slots)
    assert list(entries["vertex_count"]) == [0, 36, 36, 36] and total == 3 * align128(720)
    compare_batch(ctx, pos, slots, entries, uv_default, ids=ids)
    e0, t0 = ctx.mesh_chunks(np.array([], dtype=np.uint32))
    assert len(e0) == 0 and t0 == 0


def test_large_chunk_coordinates_round_like_f32(ctx, uv_default):
    pos = np.array([(2 ** 21 + 3, -(2 ** 22) - 1, 2 ** 27 - 5), (-(2 ** 27) + 1, 2 ** 20, 7)], dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_CHECKER, 0, np.zeros((2, 3), dtype=np.int32))
    slots = ctx.load_chunks(pos, ids)
    entries, _ = ctx.mesh_chunks(slots)
    compare_batch(ctx, pos, slots, entries, uv_default, ids=ids)


def test_arena_overflow_is_reported_with_the_needed_size(uv_default):
    with capi.Context(0, 64, 256 * 1024) as c:  # 256 KiB arena, one full chunk needs 180 KiB
        c.set_uv_table(uv_default)
        slots = c.generate_chunks([(0, -1, 0), (1, -1, 0), (2, -1, 0)], capi.GEN_REFERENCE)
        with pytest.raises(capi.DspError) as ei:
            c.mesh_chunks(slots)
        assert ei.value.status == capi.ERR_ARENA_FULL and str(3 * 184320) in str(ei.value)
        entries, total = c.mesh_chunks(slots[:1])  # still usable afterwards
        assert total == 184320
        with pytest.raises(capi.DspError):
            c.mesh_chunks(slots[:1], arena_offset=100)  # not a multiple of 128


def test_unload_frees_slots_and_store_full_is_reported(uv_default):
    with capi.Context(0, 4, 1 << 20) as c:
        c.set_uv_table(uv_default)
        slots = c.generate_chunks([(i, 0, 0) for i in range(4)], capi.GEN_REFERENCE)
        with pytest.raises(capi.DspError) as ei:
            c.generate_chunks([(9, 0, 0)], capi.GEN_REFERENCE)
        assert ei.value.status == capi.ERR_STORE_FULL
        c.unload_chunks(slots[1:3])
        assert c.resident_chunks == 2 and c.find_chunk((1, 0, 0)) is None
        with pytest.raises(capi.DspError) as ei:
            c.mesh_chunks(slots)  # contains unloaded slots
        assert ei.value.status == capi.ERR_NOT_RESIDENT
        s2 = c.generate_chunks([(9, 0, 0), (8, -1, 0)], capi.GEN_REFERENCE)
        assert set(map(int, s2)) == {int(slots[1]), int(slots[2])}
        assert c.find_chunk((0, 0, 0)) == int(slots[0])  # regenerating a resident position reuses its slot
        assert int(c.generate_chunks([(0, 0, 0)], capi.GEN_REFERENCE)[0]) == int(slots[0])


def test_async_api(ctx, uv_default):
    first = ctx.generate_region((0, 6, 0), (4, 2, 4), capi.GEN_NOISE, SEED_NOISE)
    ctx.mesh_range_async(first, 32)
    ctx.sync()
    entries, total = ctx.mesh_result(32)
    check_entries_layout(entries, total)
    pos = np.array([(a, 6 + b, c) for c in range(4) for b in range(2) for a in range(4)], dtype=np.int32)
    compare_batch(ctx, pos, None, entries, uv_default, orc.GEN_NOISE, SEED_NOISE)


# ---- edits (config 4 at a size the oracle finishes in seconds) -----------------------------------------------
def test_edits_then_remesh_dirty_chunks(ctx, uv_default):
    dims = (4, 3, 4)
    origin = (0, 6, 0)
    first = ctx.generate_region(origin, dims, capi.GEN_NOISE, SEED_NOISE)
    n = dims[0] * dims[1] * dims[2]
    pos = np.array([(origin[0] + a, origin[1] + b, origin[2] + c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_NOISE, SEED_NOISE, pos)
    rng = np.random.default_rng(4)
    k = 3000
    edits = np.zeros(k, dtype=capi.EDIT_DTYPE)
    edits["wx"] = rng.integers(origin[0] * 16 - 8, (origin[0] + dims[0]) * 16 + 8, k)  # some fall outside: ignored
    edits["wy"] = rng.integers(origin[1] * 16, (origin[1] + dims[1]) * 16, k)
    edits["wz"] = rng.integers(origin[2] * 16, (origin[2] + dims[2]) * 16, k)
    edits["block"] = np.where(rng.random(k) < 0.5, 0, rng.integers(1, 3, k))
    edits[10:20] = edits[0:10]      # repeated voxels: the later edit must win
    edits["block"][10:20] = 2
    dirty = ctx.apply_edits(edits)
    want_dirty = set()
    for e in edits:  # world.rs:69-74 addressing
        c = (int(e["wx"]) >> 4, int(e["wy"]) >> 4, int(e["wz"]) >> 4)
        a, b, cc = c[0] - origin[0], c[1] - origin[1], c[2] - origin[2]
        if 0 <= a < dims[0] and 0 <= b < dims[1] and 0 <= cc < dims[2]:
            i = a + dims[0] * (b + dims[1] * cc)
            ids[i, (int(e["wx"]) & 15) + 16 * (int(e["wy"]) & 15) + 256 * (int(e["wz"]) & 15)] = e["block"]
            want_dirty.add(first + i)
    assert sorted(want_dirty) == list(map(int, dirty))
    for s in dirty[:8]:
        assert np.array_equal(ctx.download_chunk(int(s))[0], ids[int(s) - first])
    entries, total = ctx.mesh_chunks(dirty)
    check_entries_layout(entries, total)
    sel = np.array([int(s) - first for s in dirty])
    compare_batch(ctx, pos[sel], dirty, entries, uv_default, ids=ids[sel])


# ---- halo-culling extension (not reference behaviour; checked against the extended oracle) ---------------------
def test_halo_mode_matches_extended_oracle(ctx, uv_default):
    dims = (3, 3, 3)
    first = ctx.generate_region((0, 0, 0), dims, capi.GEN_RANDOM, SEED_RANDOM)
    extra = ctx.generate_chunks([(3, 1, 1), (-1, 1, 1), (1, 1, 4)], capi.GEN_NOISE, SEED_NOISE)  # one detached at z=4
    world = {}
    for c in range(3):
        for b in range(3):
            for a in range(3):
                world[(a, b, c)] = (first + a + 3 * (b + 3 * c), orc.generate(orc.GEN_RANDOM, SEED_RANDOM, (a, b, c)))
    for p, s in zip([(3, 1, 1), (-1, 1, 1), (1, 1, 4)], extra):
        world[p] = (int(s), orc.generate(orc.GEN_NOISE, SEED_NOISE, p))
    positions = list(world)
    slots = np.array([world[p][0] for p in positions], dtype=np.uint32)
    entries, total = ctx.mesh_chunks(slots, capi.MESH_HALO)
    check_entries_layout(entries, total)
    dirs = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
    fewer = 0
    for p, e in zip(positions, entries):
        slabs = np.zeros((6, 256), dtype=np.uint16)
        present = np.zeros(6, dtype=np.uint8)
        for f, d in enumerate(dirs):
            q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
            if q not in world:
                continue
            v = world[q][1].reshape(16, 16, 16)  # [z][y][x]
            present[f] = 1
            slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
        want = orc.mesh_chunk_halo(p, world[p][1], slabs, present, uv_default)
        assert ctx.download_vertices(e).tobytes() == want.tobytes(), p
        fewer += orc.mesh_chunk(p, world[p][1], uv_default).shape[0] - want.shape[0]
    assert fewer > 0
    # the unit faces of halo mode are a subset of the reference's faces for the same chunk
    e_par, _ = ctx.mesh_chunks(slots[:5], capi.MESH_PARITY, arena_offset=align128(total))
    for ep, eh in zip(e_par, entries[:5]):
        fp = {bytes(r) for r in ctx.download_vertices(ep).view(np.uint8).reshape(-1, 120)}
        fh = {bytes(r) for r in ctx.download_vertices(eh).view(np.uint8).reshape(-1, 120)}
        assert fh <= fp


def test_two_contexts_with_seam_exchange_equal_one_context(uv_default):
    """Config-5 mechanics at a small size: the world is cut into two X-slabs held by two contexts (two 'ranks' on one
    GPU); border slabs are packed on one, handed over in device memory, registered on the other; halo-mode meshes of
    every chunk must equal those of a single context that holds the whole world."""
    import torch
    from despawnengine_b200 import partition as part
    from despawnengine_b200.halo import RegionOnDevice
    world = part.Region((3, 5, -2), (6, 4, 3))
    sp = part.SlabPartition(world, 2)
    with capi.Context(0, 256, 256 << 20) as whole, capi.Context(0, 128, 128 << 20) as c0, capi.Context(0, 128, 128 << 20) as c1:
        for c in (whole, c0, c1):
            c.set_uv_table(uv_default)
        ref = RegionOnDevice(whole, world, capi.GEN_NOISE, SEED_NOISE)
        devs = [RegionOnDevice(c0, sp.region(0), capi.GEN_NOISE, SEED_NOISE), RegionOnDevice(c1, sp.region(1), capi.GEN_NOISE, SEED_NOISE)]
        # the exchange, done by hand instead of through torch.distributed (one process here)
        for rank, dev in enumerate(devs):
            for t in sp.seam_transfers(rank):
                peer = devs[t.peer]
                peer_positions = sp.region(t.peer).face_positions(part.OPPOSITE_FACE[t.send_face])
                buf = torch.empty((len(peer_positions), 512), dtype=torch.uint8, device="cuda")
                peer.ctx.pack_border_slabs(peer.slots_of(peer_positions), part.OPPOSITE_FACE[t.send_face], buf.data_ptr())
                peer.ctx.sync()
                dev.ctx.set_halo_slabs(dev.slots_of(t.positions), t.send_face, buf.data_ptr())
                dev.ctx.sync()
        e_ref, _ = whole.mesh_range(ref.first_slot, world.n_chunks, capi.MESH_HALO)
        fewer = 0
        for rank, dev in enumerate(devs):
            reg = sp.region(rank)
            e, total = dev.ctx.mesh_range(dev.first_slot, reg.n_chunks, capi.MESH_HALO)
            check_entries_layout(e, total)
            e_par, _ = dev.ctx.mesh_range(dev.first_slot, reg.n_chunks, capi.MESH_PARITY, arena_offset=align128(total))
            for i, p in enumerate(reg.positions()):
                want = whole.download_vertices(e_ref[world.index_of(p)])
                assert dev.ctx.download_vertices(e[i]).tobytes() == want.tobytes(), (rank, tuple(p))
            fewer += int(e_par["vertex_count"].sum()) - int(e["vertex_count"].sum())
        assert fewer > 0
        # without the exchange the seam faces come back (reference behaviour at a non-resident neighbour)
        c0.clear_halo_slabs()
        e_no, _ = c0.mesh_range(devs[0].first_slot, sp.region(0).n_chunks, capi.MESH_HALO)
        seam = [sp.region(0).index_of(p) for p in sp.region(0).face_positions(part.FACE_LEFT)]
        assert int(e_no["vertex_count"][seam].sum()) > int(np.array([e_ref[world.index_of(p)]["vertex_count"] for p in sp.region(0).face_positions(part.FACE_LEFT)]).sum())


# ---- full BASELINE config 2 size: 16x16x16 chunks ------------------------------------------------------------
def test_config2_full_region_against_oracle(ctx, uv_default):
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, SEED_NOISE)
    entries, total = ctx.mesh_range(first, 4096)
    check_entries_layout(entries, total)
    pos = np.array([(a, b, c) for c in range(16) for b in range(16) for a in range(16)], dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_NOISE, SEED_NOISE, pos)
    counts, want = orc.mesh_batch(pos, ids, uv_default, flavour=orc.FAIR)
    assert np.array_equal(entries["vertex_count"].astype(np.int64), counts)
    blob = ctx.download_arena(0, total)
    # checksum of checksums: per-chunk sha256 over the exact stream, then one digest over all
    hg, hw, w0 = hashlib.sha256(), hashlib.sha256(), 0
    wb = want.view(np.uint8).reshape(-1)
    for e, c in zip(entries, counts):
        nb, o = int(c) * 20, int(e["offset_bytes"])
        hg.update(hashlib.sha256(blob[o:o + nb].tobytes()).digest())
        hw.update(hashlib.sha256(wb[w0:w0 + nb].tobytes()).digest())
        w0 += nb
    assert hg.hexdigest() == hw.hexdigest()
