"""Config 4 semantics (incremental remesh): block edits in world coordinates (world.rs:50-74 addressing, Chunk::set_block
chunk.rs:49-68), edits of one call apply in order (the last edit of a voxel wins), the dirty chunks are re-meshed whole
(build_chunk_mesh is the reference's only granularity).  Both addressing paths -- on the device for box-region worlds, on
the host for chunks held in the position map -- must give the same voxels, dirty list and byte-exact meshes."""
import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc

pytestmark = pytest.mark.gpu
SEED = 0xD5E50001


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def make_edits(rng, k, lo, hi, dup_frac=0.3):
    e = np.zeros(k, dtype=capi.EDIT_DTYPE)
    for a, name in enumerate(("wx", "wy", "wz")):
        e[name] = rng.integers(lo[a], hi[a], k)
    e["block"] = np.where(rng.random(k) < 0.5, 0, rng.integers(1, 4, k))
    nd = int(k * dup_frac)  # many repeated voxels with different blocks: order matters
    src = rng.integers(0, k, nd)
    dst = rng.integers(0, k, nd)
    for name in ("wx", "wy", "wz"):
        e[name][dst] = e[name][src]
    return e


def replay(edits, chunks):
    """chunks: {chunk pos: ids (modified in place)}.  Returns the set of touched chunk positions."""
    touched = set()
    for e in edits:
        c = (int(e["wx"]) >> 4, int(e["wy"]) >> 4, int(e["wz"]) >> 4)  # div_euclid(16)
        if c in chunks:
            chunks[c][(int(e["wx"]) & 15) + 16 * (int(e["wy"]) & 15) + 256 * (int(e["wz"]) & 15)] = e["block"]  # rem_euclid(16)
            touched.add(c)
    return touched


@pytest.mark.parametrize("layout", ["regions", "map", "mixed"])
@pytest.mark.parametrize("fused", [False, True])
def test_edits_apply_in_order_and_dirty_chunks_remesh_byte_exact(uv_default, layout, fused):
    rng = np.random.default_rng(11)
    with capi.Context(0, 1024, 256 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        slot_of = {}
        boxes = [((-2, 5, -3), (3, 3, 4)), ((4, 6, 1), (2, 2, 2))]
        if layout in ("regions", "mixed"):
            for origin, dims in boxes:
                first = ctx.generate_region(origin, dims, capi.GEN_NOISE, SEED)
                for c in range(dims[2]):
                    for b in range(dims[1]):
                        for a in range(dims[0]):
                            slot_of[(origin[0] + a, origin[1] + b, origin[2] + c)] = first + a + dims[0] * (b + dims[1] * c)
        if layout in ("map", "mixed"):
            extra = [(-3, 6, 0), (-3, 7, 0), (10, 6, 10)] if layout == "mixed" else [(origin[0] + a, origin[1] + b, origin[2] + c) for origin, dims in boxes
                                                                                      for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])]
            for p, s in zip(extra, ctx.generate_chunks(extra, capi.GEN_NOISE, SEED)):
                slot_of[p] = int(s)
        chunks = {p: orc.generate(orc.GEN_NOISE, SEED, p) for p in slot_of}
        pos_of = {s: p for p, s in slot_of.items()}
        for frame in range(3):
            edits = make_edits(rng, 4000, (-4 * 16, 4 * 16, -4 * 16), (12 * 16, 9 * 16, 12 * 16))
            if fused:
                dirty, entries, total = ctx.apply_edits_and_remesh(edits)
            else:
                dirty = ctx.apply_edits(edits)
                entries, total = ctx.mesh_chunks(dirty)
            touched = replay(edits, chunks)
            assert list(map(int, dirty)) == sorted(slot_of[p] for p in touched)
            assert total == int(sum(((int(e["vertex_count"]) * 20 + 127) & ~127) for e in entries))
            for s, e in zip(dirty, entries):
                p = pos_of[int(s)]
                assert np.array_equal(ctx.download_chunk(int(s))[0], chunks[p]), p
                assert ctx.download_vertices(e).tobytes() == orc.mesh_chunk(p, chunks[p], uv_default).tobytes(), p


def test_many_edits_to_one_voxel_and_no_resident_target(uv_default):
    with capi.Context(0, 64, 16 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        first = ctx.generate_region((0, -1, 0), (1, 1, 1), capi.GEN_REFERENCE, 0)  # one full chunk
        e = np.zeros(5000, dtype=capi.EDIT_DTYPE)
        e["wx"], e["wy"], e["wz"] = 3, -5, 9
        e["block"] = np.arange(5000) % 3
        e["block"][-1] = 0  # the last one destroys the block
        dirty, entries, _ = ctx.apply_edits_and_remesh(e)
        ids = orc.generate(orc.GEN_REFERENCE, 0, (0, -1, 0))
        ids[3 + 16 * 11 + 256 * 9] = 0
        assert list(dirty) == [first] and np.array_equal(ctx.download_chunk(first)[0], ids)
        assert ctx.download_vertices(entries[0]).tobytes() == orc.mesh_chunk((0, -1, 0), ids, uv_default).tobytes()
        far = np.zeros(10, dtype=capi.EDIT_DTYPE)
        far["wx"] = 1000
        dirty, entries, total = ctx.apply_edits_and_remesh(far)
        assert len(dirty) == 0 and total == 0
        bad = np.zeros(1, dtype=capi.EDIT_DTYPE)
        bad["block"] = 70000
        with pytest.raises(capi.DspError) as ei:
            ctx.apply_edits(bad)
        assert ei.value.status == capi.ERR_BAD_ARG


@pytest.mark.parametrize("layout", ["regions", "map"])
def test_halo_mode_edits_also_remesh_the_neighbour_across_an_edited_border(uv_default, layout):
    """SURVEY.md section 8e, incremental case: with cross-chunk culling an edit on a border voxel changes the neighbour's mesh."""
    rng = np.random.default_rng(5)
    origin, dims = (0, 0, 0), (3, 3, 3)
    positions = [(origin[0] + a, origin[1] + b, origin[2] + c) for c in range(3) for b in range(3) for a in range(3)]
    with capi.Context(0, 256, 256 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        if layout == "regions":
            first = ctx.generate_region(origin, dims, capi.GEN_RANDOM, 21)
            slot_of = {p: first + i for i, p in enumerate(positions)}
        else:
            slot_of = {p: int(s) for p, s in zip(positions, ctx.generate_chunks(positions, capi.GEN_RANDOM, 21))}
        chunks = {p: orc.generate(orc.GEN_RANDOM, 21, p) for p in positions}
        pos_of = {s: p for p, s in slot_of.items()}
        dirs = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]

        def halo_mesh(p):
            slabs = np.zeros((6, 256), dtype=np.uint16)
            present = np.zeros(6, dtype=np.uint8)
            for f, d in enumerate(dirs):
                q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                if q in chunks:
                    v = chunks[q].reshape(16, 16, 16)
                    present[f] = 1
                    slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
            return orc.mesh_chunk_halo(p, chunks[p], slabs, present, uv_default)

        for frame in range(3):
            k = 300
            e = np.zeros(k, dtype=capi.EDIT_DTYPE)
            for name in ("wx", "wy", "wz"):
                coarse = rng.integers(0, 3, k) * 16
                fine = np.where(rng.random(k) < 0.6, rng.choice([0, 15], k), rng.integers(0, 16, k))  # mostly border layers
                e[name] = coarse + fine
            e["block"] = rng.integers(0, 3, k)
            dirty, entries, total = ctx.apply_edits_and_remesh(e, capi.MESH_HALO)
            want = set()
            for ed in e:
                w = (int(ed["wx"]), int(ed["wy"]), int(ed["wz"]))
                c = tuple(v >> 4 for v in w)
                chunks[c][(w[0] & 15) + 16 * (w[1] & 15) + 256 * (w[2] & 15)] = ed["block"]
                want.add(c)
                for a in range(3):
                    if (w[a] & 15) in (0, 15):
                        q = list(c); q[a] += -1 if (w[a] & 15) == 0 else 1
                        if tuple(q) in chunks:
                            want.add(tuple(q))
            assert list(map(int, dirty)) == sorted(slot_of[p] for p in want)
            for s, en in zip(dirty, entries):
                p = pos_of[int(s)]
                assert ctx.download_vertices(en).tobytes() == halo_mesh(p).tobytes(), (frame, p)
            # and the whole world is consistent: nothing outside the dirty set needed a remesh
            all_slots = np.array([slot_of[p] for p in positions], dtype=np.uint32)
            entries_all, _ = ctx.mesh_chunks(all_slots, capi.MESH_HALO, arena_offset=(total + 127) & ~127)
            for p, en in zip(positions, entries_all):
                assert ctx.download_vertices(en).tobytes() == halo_mesh(p).tobytes(), ("full", frame, p)


def test_get_block_world_reads_what_edits_wrote(uv_default):
    """World::get_block_world (world.rs:50-65) on the device store, negative coordinates included."""
    with capi.Context(0, 64, 16 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        ctx.generate_chunks([(-1, 0, -1), (0, 0, 0)], capi.GEN_REFERENCE)
        assert ctx.get_block_world(-1, 7, -16) == 1 and ctx.get_block_world(-1, 8, -16) == 0   # flat chunk: y < 8 dirt
        assert ctx.get_block_world(16, 0, 0) is None                                           # chunk (1,0,0) is not resident
        e = np.zeros(2, dtype=capi.EDIT_DTYPE)
        e["wx"], e["wy"], e["wz"], e["block"] = [-16, 15], [15, 0], [-1, 15], [2, 0]
        ctx.apply_edits(e)
        assert ctx.get_block_world(-16, 15, -1) == 2 and ctx.get_block_world(15, 0, 15) == 0


def test_buried_chunks_are_skipped_by_summary_until_an_edit_touches_them(uv_default):
    """Halo mode drops all-air chunks and chunks buried among uniformly solid neighbours before the mesh kernel (chunk
    summaries written by the generators).  An edit must reset the summary: carving a cavity INSIDE a buried chunk gives it
    faces again, and opening a hole in its border gives faces to the neighbour too."""
    origin, dims = (0, -4, 0), (3, 6, 3)  # reference terrain: y < 0 solid, y == 0 flat, y > 0 air
    positions = [(origin[0] + a, origin[1] + b, origin[2] + c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])]
    dirs = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
    with capi.Context(0, 256, 256 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        first = ctx.generate_region(origin, dims, capi.GEN_REFERENCE, 0)
        chunks = {p: orc.generate(orc.GEN_REFERENCE, 0, p) for p in positions}
        slot_of = {p: first + i for i, p in enumerate(positions)}

        def halo_mesh(p):
            slabs, present = np.zeros((6, 256), dtype=np.uint16), np.zeros(6, dtype=np.uint8)
            for f, d in enumerate(dirs):
                q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                if q in chunks:
                    v = chunks[q].reshape(16, 16, 16)
                    present[f] = 1
                    slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
            return orc.mesh_chunk_halo(p, chunks[p], slabs, present, uv_default)

        def check_world(mode, greedy):
            entries, total = ctx.mesh_range(first, len(positions), mode)
            for p, e in zip(positions, entries):
                if greedy:
                    continue
                assert ctx.download_vertices(e).tobytes() == halo_mesh(p).tobytes(), p
            return entries

        e0 = check_world(capi.MESH_HALO, False)
        buried = (1, -2, 1)
        assert int(e0[positions.index(buried)]["vertex_count"]) == 0 and int(e0[positions.index((1, 1, 1))]["vertex_count"]) == 0
        # a cavity in the middle of the buried chunk, and a hole through its +x border layer
        ed = np.zeros(2, dtype=capi.EDIT_DTYPE)
        ed["wx"], ed["wy"], ed["wz"], ed["block"] = [16 + 8, 16 + 15], [-32 + 8, -32 + 3], [16 + 8, 16 + 4], [0, 0]
        dirty, entries, _ = ctx.apply_edits_and_remesh(ed, capi.MESH_HALO)
        for w in ((24, -24, 24), (31, -29, 20)):
            chunks[buried][(w[0] & 15) + 16 * (w[1] & 15) + 256 * (w[2] & 15)] = 0
        assert sorted(map(int, dirty)) == sorted([slot_of[buried], slot_of[(2, -2, 1)]])
        for s, e in zip(dirty, entries):
            p = positions[int(s) - first]
            assert int(e["vertex_count"]) > 0 and ctx.download_vertices(e).tobytes() == halo_mesh(p).tobytes(), p
        e1 = check_world(capi.MESH_HALO, False)
        assert int(e1[positions.index(buried)]["vertex_count"]) == 6 * (6 + 5) and int(e1[positions.index((2, -2, 1))]["vertex_count"]) == 6 * (256 + 1)  # world-edge +x faces + the one behind the hole
        # quad mode goes through the same pre-pass
        eg, _ = ctx.mesh_range(first, len(positions), capi.MESH_HALO | capi.MESH_GREEDY | capi.MESH_INDEXED)
        assert int(eg[positions.index((0, -2, 0))]["vertex_count"]) > 0 and int(eg[positions.index((1, -3, 1))]["vertex_count"]) == 0
