"""Chunk summaries in the parity mesh kernel and the two-pass DSP_MESH_ORDERED mode (GPU, through the C ABI).

The parity kernel consults the generators' chunk summaries: an all-air chunk ends at the ticket (the reference's
palette.len() == 1 early-out, /root/reference/src/content/world/chunks/chunk_mesh.rs:18-20), a chunk of one known solid block
id is meshed from that id without its tile being read.  Neither may change a byte of the result: every case below is compared
with the CPU oracle, and with the same call on a context created under DSP_SUMMARY_SKIP=0 (every tile read)."""
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


def streams_of(ctx, entries):
    return [ctx.download_vertices(e).tobytes() for e in entries]


def oracle_streams(pos, ids, uv):
    counts, blob = orc.mesh_batch(pos, ids, uv, flavour=orc.FAIR)
    blob = blob.view(np.uint8).reshape(-1)
    out, o = [], 0
    for c in counts:
        out.append(blob[o:o + int(c) * 20].tobytes())
        o += int(c) * 20
    return counts, out


@pytest.mark.parametrize("kind,seed,origin,dims", [(capi.GEN_NOISE, SEED_NOISE, (3, 0, -2), (6, 16, 5)), (capi.GEN_REFERENCE, 0, (-2, -2, 1), (4, 4, 3))])
def test_summary_paths_change_no_byte(uv_default, kind, seed, origin, dims):
    """Air / one-solid-id / surface chunks of a generated box: summaries used (default), not used (DSP_SUMMARY_SKIP=0) and
    invalidated give the same entries' counts and the same streams, all equal to the oracle's."""
    pos = box_positions(origin, dims)
    n = len(pos)
    ids = orc.generate_batch(kind, seed, pos)
    counts, want = oracle_streams(pos, ids, uv_default)
    kinds = {"air": int((ids == 0).all(axis=1).sum()), "uniform": int(((ids == ids[:, :1]).all(axis=1) & (ids[:, 0] != 0)).sum())}
    assert kinds["air"] > 0 and kinds["uniform"] > 0 and kinds["air"] + kinds["uniform"] < n, kinds  # the box exercises all three paths
    results = {}
    for label, env in (("summaries", None), ("no skip", "0"), ("air only", "1")):
        old = os.environ.get("DSP_SUMMARY_SKIP")
        if env is None:
            os.environ.pop("DSP_SUMMARY_SKIP", None)
        else:
            os.environ["DSP_SUMMARY_SKIP"] = env
        try:
            with capi.Context(0, n + 8, 1 << 30) as c:
                c.set_uv_table(uv_default)
                first = c.generate_region(origin, dims, kind, seed)
                for mode in (capi.MESH_PARITY, capi.MESH_ORDERED):
                    e, total = c.mesh_range(first, n, mode)
                    assert np.array_equal(e["vertex_count"].astype(np.int64), counts), (label, mode)
                    assert total == sum(align128(int(k) * 20) for k in counts)
                    assert streams_of(c, e) == want, (label, mode)
                if env is None:
                    c.invalidate_chunks(np.arange(first, first + n, dtype=np.uint32))
                    e, total = c.mesh_range(first, n)
                    assert streams_of(c, e) == want, "invalidated"
                results[label] = total
        finally:
            if old is None:
                os.environ.pop("DSP_SUMMARY_SKIP", None)
            else:
                os.environ["DSP_SUMMARY_SKIP"] = old
    assert len(set(results.values())) == 1


def test_edit_in_air_and_uniform_chunks_resets_their_summaries(uv_default):
    """Chunk::set_block (chunk.rs:49-68) on an all-air chunk and on a one-solid-id chunk: the summaries may not survive the edit."""
    origin, dims = (0, 0, 0), (2, 16, 2)
    pos = box_positions(origin, dims)
    n = len(pos)
    ids = orc.generate_batch(capi.GEN_NOISE, SEED_NOISE, pos)
    air = int(np.flatnonzero((ids == 0).all(axis=1))[0])
    uni = int(np.flatnonzero((ids == ids[:, :1]).all(axis=1) & (ids[:, 0] != 0))[0])
    with capi.Context(0, n, 256 << 20) as c:
        c.set_uv_table(uv_default)
        first = c.generate_region(origin, dims, capi.GEN_NOISE, SEED_NOISE)
        c.mesh_range(first, n)  # summaries in use
        edits = []
        for chunk, local, block in ((air, (5, 6, 7), 2), (uni, (0, 15, 3), 0), (uni, (8, 8, 8), 2)):
            p = pos[chunk]
            edits.append((int(p[0]) * 16 + local[0], int(p[1]) * 16 + local[1], int(p[2]) * 16 + local[2], block))
            ids[chunk, local[0] + 16 * local[1] + 256 * local[2]] = block
        ed = np.zeros(len(edits), dtype=capi.EDIT_DTYPE)
        for k, (wx, wy, wz, block) in enumerate(edits):
            ed[k] = (wx, wy, wz, block)
        c.apply_edits(ed)
        _, want = oracle_streams(pos, ids, uv_default)
        for mode in (capi.MESH_PARITY, capi.MESH_ORDERED):
            e, _ = c.mesh_range(first, n, mode)
            assert streams_of(c, e) == want, mode
        assert len(want[air]) == 36 * 20  # one voxel in the former air chunk: six faces


@pytest.mark.parametrize("n", [1, 15, 16, 17, 33, 700])
def test_ordered_mode_ragged_counts_and_slot_lists(uv_default, n):
    """The count pass takes 16 chunks per CTA and the scan 2,048 per sweep: counts that are no multiple of either, through an
    explicit (shuffled) slot list, packed in the order of the list."""
    rng = np.random.default_rng(n)
    dims = (4, 16, 11)
    pos_all = box_positions((0, 0, 0), dims)
    ids_all = orc.generate_batch(capi.GEN_NOISE, SEED_NOISE, pos_all)
    with capi.Context(0, len(pos_all), 1 << 30) as c:
        c.set_uv_table(uv_default)
        first = c.generate_region((0, 0, 0), dims, capi.GEN_NOISE, SEED_NOISE)
        pick = rng.permutation(len(pos_all))[:n]
        slots = (first + pick).astype(np.uint32)
        e, total = c.mesh_chunks(slots, capi.MESH_ORDERED)
        counts, want = oracle_streams(pos_all[pick], ids_all[pick], uv_default)
        sizes = np.array([align128(int(k) * 20) for k in counts], dtype=np.int64)
        assert np.array_equal(e["vertex_count"].astype(np.int64), counts) and total == int(sizes.sum())
        assert np.array_equal(e["offset_bytes"].astype(np.int64), np.concatenate([[0], np.cumsum(sizes)[:-1]]))
        assert streams_of(c, e) == want
        # the same list at an arena offset, twice: reproducible layout
        e2, t2 = c.mesh_chunks(slots, capi.MESH_ORDERED, arena_offset=align128(total) + 128)
        assert t2 == total and np.array_equal(e2["offset_bytes"] - e["offset_bytes"], np.full(n, align128(total) + 128, dtype=np.uint64))


def test_ordered_mode_reports_arena_overflow_with_the_size_needed(uv_default):
    with capi.Context(0, 64, 256 * 1024) as c:  # one full chunk needs 180 KiB
        c.set_uv_table(uv_default)
        slots = c.generate_chunks([(0, -1, 0), (1, -1, 0), (2, -1, 0)], capi.GEN_REFERENCE)
        with pytest.raises(capi.DspError) as ei:
            c.mesh_chunks(slots, capi.MESH_ORDERED)
        assert ei.value.status == capi.ERR_ARENA_FULL and str(3 * 184320) in str(ei.value)
        e, total = c.mesh_chunks(slots[:1], capi.MESH_ORDERED)  # still usable afterwards, in both modes
        assert total == 184320
        e, total = c.mesh_chunks(slots[:1])
        assert total == 184320
