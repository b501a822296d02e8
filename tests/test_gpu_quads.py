"""GPU parity for the quad modes (DSP_MESH_GREEDY / DSP_MESH_INDEXED) through the C ABI, byte
for byte against the extension oracle (oracle/mesher_oracle.cpp: merge_quads / emit_quads)."""
import hashlib

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc

pytestmark = pytest.mark.gpu

SEED_NOISE, SEED_RANDOM = 0xD5E50001, 0xD5E50003
MODES = [(capi.MESH_GREEDY, True, False), (capi.MESH_INDEXED, False, True), (capi.MESH_GREEDY | capi.MESH_INDEXED, True, True)]


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


def slot_bytes(e):
    return align128(int(e["vertex_count"]) * 20) + align128(int(e["index_count"]) * 4)


def check_layout(entries, total, base=0):
    sizes = np.array([slot_bytes(e) for e in entries], dtype=np.int64)
    offs = entries["offset_bytes"].astype(np.int64)
    assert total == int(sizes.sum()) and (offs % 128 == 0).all()
    nz = sizes > 0
    order = np.argsort(offs[nz], kind="stable")
    o, z = offs[nz][order], sizes[nz][order]
    if len(o):
        assert o[0] == base and np.array_equal(o[1:], o[:-1] + z[:-1]) and o[-1] + z[-1] == base + total


def compare(ctx, positions, ids_list, entries, uv, greedy, indexed, halo=None):
    for k, (p, ids, e) in enumerate(zip(positions, ids_list, entries)):
        kw = {} if halo is None else {"nbr_slabs": halo[k][0], "nbr_present": halo[k][1]}
        v, i, q = orc.mesh_chunk_ext(p, ids, uv, greedy=greedy, indexed=indexed, **kw)
        assert int(e["vertex_count"]) == v.shape[0] and int(e["index_count"]) == i.shape[0], (p, e, v.shape, i.shape)
        assert ctx.download_vertices(e).tobytes() == v.tobytes(), p
        if indexed:
            assert np.array_equal(ctx.download_indices(e), i), p


@pytest.mark.parametrize("mode,greedy,indexed", MODES)
def test_reference_terrain(ctx, uv_default, mode, greedy, indexed):
    positions = [(0, 0, 0), (0, -1, 0), (0, 1, 0), (-3, 0, 11), (2, -7, -5)]
    slots = ctx.generate_chunks(positions, capi.GEN_REFERENCE)
    entries, total = ctx.mesh_chunks(slots, mode)
    check_layout(entries, total)
    ids = [orc.generate(orc.GEN_REFERENCE, 0, p) for p in positions]
    compare(ctx, positions, ids, entries, uv_default, greedy, indexed)
    if greedy:
        assert [int(e["vertex_count"]) for e in entries] == [q * (4 if indexed else 6) for q in (6, 6, 0, 6, 6)]


@pytest.mark.parametrize("mode,greedy,indexed", MODES)
@pytest.mark.parametrize("kind,seed,origin,dims", [(capi.GEN_NOISE, SEED_NOISE, (2, 5, 3), (4, 4, 4)), (capi.GEN_RANDOM, SEED_RANDOM, (-2, -1, 0), (3, 2, 2)),
                                                   (capi.GEN_CHECKER, 0, (0, 0, 0), (2, 1, 2))])
def test_generated_regions(ctx, uv_default, mode, greedy, indexed, kind, seed, origin, dims):
    first = ctx.generate_region(origin, dims, kind, seed)
    n = dims[0] * dims[1] * dims[2]
    entries, total = ctx.mesh_range(first, n, mode)
    check_layout(entries, total)
    positions = [(origin[0] + a, origin[1] + b, origin[2] + c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])]
    ids = [orc.generate(kind, seed, p) for p in positions]
    compare(ctx, positions, ids, entries, uv_default, greedy, indexed)


@pytest.mark.parametrize("mode,greedy,indexed", MODES)
def test_many_block_ids_and_ragged_chunks(ctx, uv_default, mode, greedy, indexed):
    """Random chunks with several block ids (one beyond the uv table -> fallback uv), very sparse and very dense, loaded
    from the host; more than one record window (> 4,096 quads) in the dense ones."""
    rng = np.random.default_rng(7)
    blocks, positions = [], []
    for k, dens in enumerate((0.0, 0.002, 0.05, 0.3, 0.5, 0.7, 0.95, 1.0)):
        blocks.append(((rng.random(4096) < dens) * rng.integers(1, 5, 4096)).astype(np.uint16))
        positions.append((k - 3, 2 * k - 5, 7 - k))
    one = np.zeros(4096, dtype=np.uint16); one[15 + 16 * 15 + 256 * 15] = 2
    blocks.append(one); positions.append((100, -100, 100))
    slots = ctx.load_chunks(positions, np.stack(blocks))
    entries, total = ctx.mesh_chunks(slots, mode)
    check_layout(entries, total)
    compare(ctx, positions, blocks, entries, uv_default, greedy, indexed)
    assert int(entries[0]["vertex_count"]) == 0 and int(entries[-1]["vertex_count"]) == (4 if indexed else 6) * 6


def test_greedy_with_halo_culling(ctx, uv_default):
    dims = (3, 3, 3)
    first = ctx.generate_region((0, 0, 0), dims, capi.GEN_NOISE, SEED_NOISE)
    # noise terrain near the surface: pick the y range the heightmap crosses
    first2 = ctx.generate_region((0, 6, 0), dims, capi.GEN_NOISE, SEED_NOISE)
    for base, oy in ((first, 0), (first2, 6)):
        world = {(a, oy + b, c): orc.generate(orc.GEN_NOISE, SEED_NOISE, (a, oy + b, c)) for c in range(3) for b in range(3) for a in range(3)}
        positions = [(a, oy + b, c) for c in range(3) for b in range(3) for a in range(3)]
        entries, total = ctx.mesh_range(base, 27, capi.MESH_HALO | capi.MESH_GREEDY | capi.MESH_INDEXED)
        check_layout(entries, total)
        dirs = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
        halo = []
        for p in positions:
            slabs = np.zeros((6, 256), dtype=np.uint16)
            present = np.zeros(6, dtype=np.uint8)
            for f, d in enumerate(dirs):
                q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                # chunks of the other region are resident too (same context), so they count as neighbours
                if q in world or (0 <= q[0] < 3 and 0 <= q[2] < 3 and (0 <= q[1] < 3 or 6 <= q[1] < 9)):
                    v = (world[q] if q in world else orc.generate(orc.GEN_NOISE, SEED_NOISE, q)).reshape(16, 16, 16)
                    present[f] = 1
                    slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
            halo.append((slabs, present))
        compare(ctx, positions, [world[p] for p in positions], entries, uv_default, True, True, halo=halo)


def test_ordered_is_refused_with_quad_modes(ctx):
    slots = ctx.generate_chunks([(0, 0, 0)], capi.GEN_REFERENCE)
    with pytest.raises(capi.DspError) as ei:
        ctx.mesh_chunks(slots, capi.MESH_ORDERED | capi.MESH_GREEDY)
    assert ei.value.status == capi.ERR_BAD_ARG


def test_arena_overflow_reports_needed_bytes_in_indexed_mode(uv_default):
    with capi.Context(0, 64, 4096) as c:
        c.set_uv_table(uv_default)
        slots = c.generate_chunks([(0, 0, 0), (1, 0, 0)], capi.GEN_CHECKER)
        with pytest.raises(capi.DspError) as ei:
            c.mesh_chunks(slots, capi.MESH_INDEXED)
        assert ei.value.status == capi.ERR_ARENA_FULL


def test_config2_full_region_greedy_indexed_checksum(ctx, uv_default):
    """BASELINE configs[1] size (4,096 chunks): checksum of per-chunk digests against the oracle, plus the size-independent
    property that the merged quads of every chunk cover exactly as many unit faces as the reference emits."""
    first = ctx.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_NOISE, SEED_NOISE)
    entries, total = ctx.mesh_range(first, 4096, capi.MESH_GREEDY | capi.MESH_INDEXED)
    check_layout(entries, total)
    e_par, _ = ctx.mesh_range(first, 4096, capi.MESH_PARITY, arena_offset=align128(total))
    blob = ctx.download_arena(0, total)
    got, want = hashlib.sha256(), hashlib.sha256()
    quads_total = 0
    for k in range(4096):
        p = (k % 16, (k // 16) % 16, k // 256)
        ids = orc.generate(orc.GEN_NOISE, SEED_NOISE, p)
        v, i, q = orc.mesh_chunk_ext(p, ids, uv_default, greedy=True, indexed=True)
        want.update(hashlib.sha256(v.tobytes() + i.tobytes()).digest())
        e = entries[k]
        o = int(e["offset_bytes"]); nv = int(e["vertex_count"]) * 20; ni = int(e["index_count"]) * 4
        oi = o + align128(nv)
        got.update(hashlib.sha256(blob[o:o + nv].tobytes() + blob[oi:oi + ni].tobytes()).digest())
        quads_total += q.shape[0]
        area = int((((q >> 15) & 15) + 1).astype(np.int64) @ (((q >> 19) & 15) + 1).astype(np.int64)) if q.shape[0] else 0
        assert area * 6 == int(e_par[k]["vertex_count"])
    assert got.hexdigest() == want.hexdigest()
    assert quads_total * 4 == int(entries["vertex_count"].astype(np.int64).sum())


def test_host_chunk_pipeline_in_quad_mode(ctx, uv_default):
    """dsp_mesh_host_chunks (H2D sub-batches overlapped with meshing) with DSP_MESH_GREEDY | DSP_MESH_INDEXED: the sub-batches
    share one arena cursor, so slots stay gap-free and disjoint across them."""
    dims = (12, 10, 10)  # 1,200 chunks -> 3 sub-batches
    pos = np.array([(a, 3 + b, c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_NOISE, SEED_NOISE, pos)
    slots, entries, total = ctx.mesh_host_chunks(pos, ids, capi.MESH_GREEDY | capi.MESH_INDEXED)
    check_layout(entries, total)
    for k in range(0, len(pos), 7):
        v, i, _ = orc.mesh_chunk_ext(pos[k], ids[k], uv_default, greedy=True, indexed=True)
        assert ctx.download_vertices(entries[k]).tobytes() == v.tobytes() and np.array_equal(ctx.download_indices(entries[k]), i), pos[k]


def test_arena_offsets_beyond_4_gib(uv_default):
    """Config 3a size: 4,096 checkerboard chunks = 6.04 GB of vertices; offsets are 64-bit end to end."""
    with capi.Context(0, 4096, 6 << 30) as c:
        c.set_uv_table(uv_default)
        first = c.generate_region((0, 0, 0), (16, 16, 16), capi.GEN_CHECKER, 0)
        entries, total = c.mesh_range(first, 4096, capi.MESH_PARITY)
        assert total == 4096 * 12288 * 120 and int(entries["offset_bytes"].max()) > (5 << 30)
        hi = int(np.argmax(entries["offset_bytes"]))
        for k in (0, hi, 2049):
            p = (k % 16, (k // 16) % 16, k // 256)
            want = orc.mesh_chunk(p, orc.generate(orc.GEN_CHECKER, 0, p), uv_default)
            assert c.download_vertices(entries[k]).tobytes() == want.tobytes(), p
        offs = np.sort(entries["offset_bytes"].astype(np.int64))
        assert np.array_equal(np.diff(offs), np.full(4095, 12288 * 120))


def _chunk_with_face_count(n_faces, mixed_ids):
    """A chunk with exactly n_faces visible faces (= quads when mixed_ids: the two voxels of a pair differ, so nothing merges).
    Lines along x at (y, z): even lines (y + z even) offer slots x0 in {0,3,6,9,12} for a pair (10 faces) or a single voxel
    (6 faces); odd lines offer single-voxel slots at x in {2,5,8,11,14}, which are only diagonal to the even lines' cells."""
    even = [(x0, y, z) for z in range(16) for y in range(16) if (y + z) % 2 == 0 for x0 in (0, 3, 6, 9, 12)]
    odd = [(x, y, z) for z in range(16) for y in range(16) if (y + z) % 2 == 1 for x in (2, 5, 8, 11, 14)]
    b = next(b for b in range(min(len(even), n_faces // 10), -1, -1) if (n_faces - 10 * b) % 6 == 0 and (n_faces - 10 * b) // 6 <= len(even) - b + len(odd))
    a = (n_faces - 10 * b) // 6
    ids = np.zeros(4096, dtype=np.uint16)
    for k, (x0, y, z) in enumerate(even[:b]):
        ids[x0 + 16 * y + 256 * z] = 1
        ids[x0 + 1 + 16 * y + 256 * z] = 2 if mixed_ids else 1
    for (x, y, z) in (even[b:] + odd)[:a]:
        ids[x + 16 * y + 256 * z] = 1
    return ids


@pytest.mark.parametrize("mode,greedy,indexed,cap", [(capi.MESH_INDEXED, False, True, 4096), (capi.MESH_GREEDY, True, False, 2048),
                                                      (capi.MESH_GREEDY | capi.MESH_INDEXED, True, True, 2048)])
def test_quad_counts_around_the_record_window(ctx, uv_default, mode, greedy, indexed, cap):
    """Chunks whose quad count is exactly the kernel's record window, one less, one more (window +- 2, counts are even), twice
    the window, and a tile boundary: the windowed record / emission loops must not drop or duplicate a quad."""
    counts = [cap - 2, cap, cap + 2, 2 * cap, 2 * cap + 2, 128, 126, 130, 256, 258]
    blocks = [_chunk_with_face_count(c, mixed_ids=True) for c in counts]
    positions = [(k, -k, 2 * k) for k in range(len(counts))]
    slots = ctx.load_chunks(positions, np.stack(blocks))
    entries, total = ctx.mesh_chunks(slots, mode)
    check_layout(entries, total)
    assert [int(e["vertex_count"]) // (4 if indexed else 6) for e in entries] == counts
    compare(ctx, positions, blocks, entries, uv_default, greedy, indexed)


@pytest.mark.parametrize("mode,greedy,indexed", MODES)
def test_structured_patterns(ctx, uv_default, mode, greedy, indexed):
    """Runs that change block id, shifted runs, cavities, plates in every orientation (tests/patterns.py)."""
    from patterns import structured_chunks
    pats = structured_chunks()
    positions = [(k - 5, 3 - k, 2 * k) for k in range(len(pats))]
    blocks = list(pats.values())
    slots = ctx.load_chunks(positions, np.stack(blocks))
    entries, total = ctx.mesh_chunks(slots, mode)
    check_layout(entries, total)
    compare(ctx, positions, blocks, entries, uv_default, greedy, indexed)
