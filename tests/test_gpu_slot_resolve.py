"""Position -> slot resolution of the batch calls (World::get_chunk's lookup, world.rs:29-31).  The batch calls try "the slot
after the previous answer" before the hash map / box list (acquire_slot in csrc/dsp_api.cu); dsp_find_chunk never does.  Whatever
order a caller submits positions in, both must name the same slot, and a position must keep its slot until it is unloaded."""
import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def check(ctx, pos):
    """one batch call over `pos` (resident or not); every answer against the independent lookup"""
    pos = np.asarray(pos, dtype=np.int32).reshape(-1, 3)
    slots = ctx.generate_chunks(pos, capi.GEN_REFERENCE)
    for p, s in zip(pos, slots):
        assert ctx.find_chunk(tuple(int(v) for v in p)) == int(s), (tuple(p), int(s))
    return slots


def test_map_slots_in_every_order():
    rng = np.random.default_rng(7)
    with capi.Context(0, 2048, 64 << 20) as ctx:
        pos = np.array([(x, y, z) for z in range(-3, 3) for y in range(-2, 2) for x in range(-4, 4)], dtype=np.int32)
        first = check(ctx, pos)                      # fresh slots, ascending
        assert list(first) == list(range(len(pos)))
        assert (check(ctx, pos) == first).all()      # the same order again: every answer is "the next slot"
        assert (check(ctx, pos[::-1]) == first[::-1]).all()
        perm = rng.permutation(len(pos))
        assert (check(ctx, pos[perm]) == first[perm]).all()
        assert (check(ctx, pos[::3]) == first[::3]).all()    # ascending with gaps: prediction misses every time
        runs = np.concatenate([np.arange(40, 80), np.arange(0, 40), np.arange(100, 120), np.arange(80, 100)])
        assert (check(ctx, pos[runs]) == first[runs]).all()  # runs of consecutive slots with jumps between them
        # unload a few: their slots come back for OTHER positions; the old neighbours must not be mistaken for them
        gone = first[[5, 6, 7, 50]]
        ctx.unload_chunks(gone)
        for i in (5, 6, 7, 50):
            assert ctx.find_chunk(tuple(int(v) for v in pos[i])) is None
        far = np.array([(100 + k, 0, 0) for k in range(6)], dtype=np.int32)
        far_slots = check(ctx, far)
        assert set(int(s) for s in gone) <= set(int(s) for s in far_slots)  # recycled
        again = check(ctx, pos)                       # positions 5, 6, 7, 50 are loaded afresh, everything else keeps its slot
        keep = np.ones(len(pos), dtype=bool)
        keep[[5, 6, 7, 50]] = False
        assert (again[keep] == first[keep]).all()
        assert len(set(int(s) for s in again) | set(int(s) for s in far_slots)) == len(pos) + len(far)
        assert (check(ctx, far) == far_slots).all()
        assert ctx.resident_chunks == len(pos) + len(far)


def test_box_region_and_map_slots_side_by_side():
    rng = np.random.default_rng(11)
    with capi.Context(0, 4096, 64 << 20) as ctx:
        before = check(ctx, [(50, 0, 0), (51, 0, 0), (0, 0, 50)])               # map slots 0..2
        first = ctx.generate_region((0, 0, 0), (5, 3, 4), capi.GEN_NOISE, 1)     # box slots 3..62, position (0,0,0) is slot 3
        first2 = ctx.generate_region((-7, 2, 9), (2, 2, 2), capi.GEN_NOISE, 1)   # a second box right behind it
        after = check(ctx, [(52, 0, 0), (0, 0, 0), (53, 0, 0)])                  # (0,0,0) must resolve to the box, not to a fresh slot
        assert int(after[1]) == first and ctx.resident_chunks == 3 + 60 + 8 + 2
        box = np.array([(x, y, z) for z in range(4) for y in range(3) for x in range(5)], dtype=np.int32)
        want = first + np.arange(60)
        assert (check(ctx, box) == want).all()                                    # box order: row and plane wraps included
        assert (check(ctx, box[::-1]) == want[::-1]).all()
        perm = rng.permutation(60)
        assert (check(ctx, box[perm]) == want[perm]).all()
        zyx = np.array([(x, y, z) for x in range(5) for y in range(3) for z in range(4)], dtype=np.int32)  # another axis order
        assert (check(ctx, zyx) == first + zyx[:, 0] + 5 * (zyx[:, 1] + 3 * zyx[:, 2])).all()
        box2 = np.array([(-7 + x, 2 + y, 9 + z) for z in range(2) for y in range(2) for x in range(2)], dtype=np.int32)
        # the end of one box runs straight into the next box's first slot, and from there into map slots
        mixed = np.concatenate([box[-7:], box2, np.array([(52, 0, 0), (53, 0, 0), (50, 0, 0), (51, 0, 0)], dtype=np.int32), box[:3]])
        got = check(ctx, mixed)
        assert (got[:7] == want[-7:]).all() and (got[7:15] == first2 + np.arange(8)).all()
        assert list(got[15:19]) == [int(after[0]), int(after[2]), int(before[0]), int(before[1])]
        # a hole in the box: the unloaded position is loaded again through the map and keeps that slot; its box neighbours are untouched
        hole = int(want[17])
        ctx.unload_chunks(np.array([hole], dtype=np.uint32))
        assert ctx.find_chunk(tuple(int(v) for v in box[17])) is None
        got = check(ctx, box)
        assert int(got[17]) != hole and (np.delete(got, 17) == np.delete(want, 17)).all()
        assert (check(ctx, box) == got).all()
        assert (check(ctx, box[perm]) == got[perm]).all()


def test_host_chunk_calls_report_the_slots_find_chunk_reports():
    """the packed upload resolves positions piece by piece on the calling thread: out_slots against dsp_find_chunk, same order and shuffled"""
    rng = np.random.default_rng(3)
    n = 600
    pos = np.array([(i % 10, (i // 10) % 6, i // 60) for i in range(n)], dtype=np.int32)
    blocks = np.zeros((n, 4096), dtype=np.uint16)
    blocks[:, :256] = rng.integers(0, 3, size=(n, 256), dtype=np.uint16)
    with capi.Context(0, 1024, 256 << 20) as ctx:
        ctx.set_uv_table(np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32))
        slots0, ent0, _ = ctx.mesh_host_chunks(pos, blocks)
        for order in (np.arange(n), rng.permutation(n), np.arange(n)[::-1].copy()):
            slots, ent, _ = ctx.mesh_host_chunks(pos[order], blocks[order])
            assert (slots == slots0[order]).all()
            assert (ent["vertex_count"] == ent0["vertex_count"][order]).all()
            for k in range(0, n, 37):
                assert ctx.find_chunk(tuple(int(v) for v in pos[order][k])) == int(slots[k])
