"""World / Chunk semantics of the C ABI that mirror the reference's (world.rs:28-47 get_chunk caching, load_chunk idempotence) and
the library's own housekeeping rules (error paths leave no trace, one chunk per position, summaries stay honest)."""
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


@pytest.fixture()
def ctx(uv_default):
    with capi.Context(0, 4096, 512 << 20) as c:
        c.set_uv_table(uv_default)
        yield c


def edit(wx, wy, wz, block):
    e = np.zeros(1, dtype=capi.EDIT_DTYPE)
    e["wx"], e["wy"], e["wz"], e["block"] = wx, wy, wz, block
    return e


def test_get_chunk_returns_the_cached_chunk_edits_included(ctx, uv_default):
    """world.rs:29-31: a second get_chunk / load_chunk of a position returns the Arc<Chunk> that is already there."""
    pos = [(0, 0, 0), (1, 0, 0)]
    slots = ctx.generate_chunks(pos, capi.GEN_REFERENCE)
    ctx.apply_edits(edit(3, 9, 4, 2))  # place a block above the flat surface of chunk (0,0,0)
    assert ctx.get_block_world(3, 9, 4) == 2
    again = ctx.generate_chunks(pos + [(2, 0, 0)], capi.GEN_REFERENCE)
    assert list(again[:2]) == list(slots) and ctx.resident_chunks == 3
    assert ctx.get_block_world(3, 9, 4) == 2, "the second load discarded an edit"
    want = orc.generate(orc.GEN_REFERENCE, 0, (0, 0, 0))
    want[3 + 16 * 9 + 256 * 4] = 2
    ids, _ = ctx.download_chunk(int(slots[0]))
    assert np.array_equal(ids, want)
    e, _ = ctx.mesh_chunks(slots[:1])
    assert ctx.download_vertices(e[0]).tobytes() == orc.mesh_chunk((0, 0, 0), want, uv_default).tobytes()
    # a different generator for a resident position changes nothing either (the reference has one generator per Y anyway)
    ctx.generate_chunks(pos, capi.GEN_CHECKER)
    assert np.array_equal(ctx.download_chunk(int(slots[0]))[0], want)
    # the explicit escape hatch regenerates
    ctx.regenerate_chunks(pos[:1], capi.GEN_REFERENCE)
    assert ctx.get_block_world(3, 9, 4) == 0
    # a position listed twice is one chunk
    s = ctx.generate_chunks([(7, 7, 7), (7, 7, 7)], capi.GEN_NOISE, SEED)
    assert s[0] == s[1] and ctx.resident_chunks == 4


def test_store_full_in_the_middle_of_a_batch_loads_nothing(uv_default):
    with capi.Context(0, 4, 1 << 20) as c:
        c.set_uv_table(uv_default)
        c.generate_chunks([(0, 0, 0), (1, 0, 0)], capi.GEN_REFERENCE)
        with pytest.raises(capi.DspError) as ei:
            c.generate_chunks([(2, 0, 0), (3, 0, 0), (4, 0, 0)], capi.GEN_REFERENCE)
        assert ei.value.status == capi.ERR_STORE_FULL
        assert c.resident_chunks == 2 and c.find_chunk((2, 0, 0)) is None
        assert len(c.generate_chunks([(2, 0, 0), (3, 0, 0)], capi.GEN_REFERENCE)) == 2


def test_duplicate_positions_in_upload_batches_are_rejected_and_leave_no_trace(ctx, uv_default):
    pos = np.array([(0, 0, 0), (1, 0, 0), (0, 0, 0)], dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_RANDOM, 5, pos)
    with pytest.raises(capi.DspError) as ei:
        ctx.load_chunks(pos, ids)
    assert ei.value.status == capi.ERR_BAD_ARG and ctx.resident_chunks == 0
    # pipelined host path, duplicate far into the batch (a later sub-batch): everything is rolled back
    n = 1500
    big = np.array([(i % 40, i // 40, 3) for i in range(n)], dtype=np.int32)
    big[1400] = big[7]
    ids = orc.generate_batch(orc.GEN_NOISE, SEED, big)
    with pytest.raises(capi.DspError) as ei:
        ctx.mesh_host_chunks(big, ids)
    assert ei.value.status == capi.ERR_BAD_ARG and ctx.resident_chunks == 0 and ctx.find_chunk((0, 0, 3)) is None
    # and the context is fine afterwards
    big[1400] = (99, 99, 99)
    ids[1400] = orc.generate(orc.GEN_NOISE, SEED, (99, 99, 99))
    slots, entries, total = ctx.mesh_host_chunks(big, ids)
    assert ctx.resident_chunks == n
    for i in (0, 7, 1400, 1499):
        assert ctx.download_vertices(entries[i]).tobytes() == orc.mesh_chunk(big[i], ids[i], uv_default).tobytes()


def test_regions_may_not_overlap_and_clones_are_identical_copies(ctx, uv_default):
    first = ctx.generate_region((0, 6, 0), (4, 3, 2), capi.GEN_NOISE, SEED)
    ctx.generate_chunks([(20, 0, 0)], capi.GEN_REFERENCE)
    for origin, dims in [((3, 8, 1), (2, 2, 2)), ((0, 6, 0), (4, 3, 2)), ((19, 0, 0), (3, 1, 1))]:
        with pytest.raises(capi.DspError) as ei:
            ctx.generate_region(origin, dims, capi.GEN_NOISE, SEED)
        assert ei.value.status == capi.ERR_BAD_ARG
    assert ctx.resident_chunks == 25
    ctx.generate_region((4, 6, 0), (1, 1, 1), capi.GEN_NOISE, SEED)  # abutting is fine
    clone = ctx.clone_region(first, (100, -3, 50))
    with pytest.raises(capi.DspError):
        ctx.clone_region(first, (2, 6, 0))
    e0, _ = ctx.mesh_range(first, 24)
    e1, t1 = ctx.mesh_range(clone, 24, arena_offset=64 << 20)
    assert np.array_equal(e0["vertex_count"], e1["vertex_count"])
    for i in range(24):
        a, b, c = i % 4, (i // 4) % 3, i // 12
        ids, p = ctx.download_chunk(clone + i)
        assert tuple(p) == (100 + a, -3 + b, 50 + c) and ctx.find_chunk(tuple(p)) == clone + i
        assert np.array_equal(ids, orc.generate(orc.GEN_NOISE, SEED, (a, 6 + b, c)))
        assert ctx.download_vertices(e1[i]).tobytes() == orc.mesh_chunk(tuple(p), ids, uv_default).tobytes()
    # edits address the clone by its own coordinates (device pipeline: all chunks live in boxes -> unload the map chunk first)
    ctx.unload_chunks([ctx.find_chunk((20, 0, 0))])
    dirty = ctx.apply_edits(edit(100 * 16 + 5, -3 * 16 + 2, 50 * 16 + 9, 7))
    assert list(dirty) == [clone] and ctx.get_block_world(100 * 16 + 5, -3 * 16 + 2, 50 * 16 + 9) == 7
    assert ctx.get_block_world(5, 6 * 16 + 2, 9) != 7


def test_invalidate_chunks_after_an_external_write(ctx, uv_default):
    """Summaries let halo-mode calls skip all-air chunks; an interop writer must be able to reset them."""
    try:
        from cuda.bindings import runtime as cudart
    except ImportError:  # older cuda-python layout
        from cuda import cudart
    first = ctx.generate_region((0, 12, 0), (2, 2, 2), capi.GEN_NOISE, SEED)  # far above the surface (h <= 239): all air
    e, _ = ctx.mesh_range(first, 8, capi.MESH_HALO)
    assert not e["vertex_count"].any()
    block = np.full(4096, 3, dtype=np.uint16)
    ctx.sync()
    (err,) = cudart.cudaMemcpy(ctx.voxel_device_ptr + (first + 5) * 8192, block.ctypes.data, 8192, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice)
    assert int(err) == 0
    stale, _ = ctx.mesh_range(first, 8, capi.MESH_HALO)
    assert not stale["vertex_count"].any(), "expected the stale summary to hide the write (that is what the call below is for)"
    ctx.invalidate_chunks([first + 5])
    e, _ = ctx.mesh_range(first, 8, capi.MESH_HALO)
    assert int(e["vertex_count"][5]) == 1536 * 6 and not np.delete(e["vertex_count"], 5).any()


def test_re_registering_a_seam_replaces_its_slabs(ctx, uv_default):
    import torch
    first = ctx.generate_region((0, 0, 0), (2, 2, 2), capi.GEN_RANDOM, 3)
    slots = np.array([first + 1, first + 3, first + 5, first + 7], dtype=np.uint32)  # the x = 1 layer; +X neighbours live elsewhere
    solid = torch.full((4, 256), 1, dtype=torch.int16, device="cuda")
    air = torch.zeros((4, 256), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.set_halo_slabs(slots, capi.FACE_LEFT, solid.data_ptr())
    e_solid, _ = ctx.mesh_range(first, 8, capi.MESH_HALO)
    l0 = ctx.kernel_launches
    ctx.mesh_range(first, 8, capi.MESH_HALO)
    per_call = ctx.kernel_launches - l0
    for _ in range(5):  # a re-exchange after every seam edit
        ctx.set_halo_slabs(slots, capi.FACE_LEFT, air.data_ptr())
    ctx.sync()
    l0 = ctx.kernel_launches
    e_air, _ = ctx.mesh_range(first, 8, capi.MESH_HALO)
    assert ctx.kernel_launches - l0 == per_call, "re-registered slabs piled up as extra batches"
    assert int(e_air["vertex_count"].sum()) > int(e_solid["vertex_count"].sum())  # the newer (air) slabs are the ones in force
    e_none = None
    ctx.clear_halo_slabs()
    e_none, _ = ctx.mesh_range(first, 8, capi.MESH_HALO)
    assert np.array_equal(e_none["vertex_count"], e_air["vertex_count"])  # no neighbour == air neighbour


def test_scene_refuses_a_context_with_hand_loaded_chunks(ctx):
    ctx.generate_chunks([(0, 0, 0)], capi.GEN_REFERENCE)
    with pytest.raises(capi.DspError) as ei:
        ctx.scene_update((2.5, 22.5, -2.5), 2, 1)
    assert ei.value.status == capi.ERR_BAD_ARG
    ctx.unload_chunks([ctx.find_chunk((0, 0, 0))])
    st = ctx.scene_update((2.5, 22.5, -2.5), 2, 1)
    assert int(st["n_loaded_total"]) > 0
