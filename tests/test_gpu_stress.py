"""Race hunting without compute-sanitizer (closed on this GPU pool): the same chunks are meshed over and over in shuffled
orders and batch sizes, in every mode, and every result must hash to what the oracle produced once.  A missing barrier in
the persistent kernels (tile reuse, record windows, staging tiles, ticket hand-over) shows up as an intermittent mismatch."""
import hashlib
import os

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def test_repeated_shuffled_batches_are_stable_in_every_mode(uv_default):
    rng = np.random.default_rng(99)
    chunks = []  # (pos, ids): a mixture that alternates cheap and expensive chunks so CTAs run out of step
    for k, (kind, seed) in enumerate([(orc.GEN_REFERENCE, 0)] * 6 + [(orc.GEN_NOISE, 0xD5E50001)] * 14 + [(orc.GEN_RANDOM, 5)] * 6 + [(orc.GEN_CHECKER, 0)] * 3):
        pos = (int(rng.integers(-4, 4)), int(rng.integers(-1, 10)) if kind != orc.GEN_NOISE else int(rng.integers(4, 10)), int(rng.integers(-4, 4)) + 9 * k)
        chunks.append((pos, orc.generate(kind, seed, pos)))
    for dens in (0.001, 0.02, 0.2, 0.8, 0.999):
        chunks.append(((50, 0, len(chunks)), ((rng.random(4096) < dens) * rng.integers(1, 4, 4096)).astype(np.uint16)))
    modes = [capi.MESH_PARITY, capi.MESH_ORDERED, capi.MESH_INDEXED, capi.MESH_GREEDY, capi.MESH_GREEDY | capi.MESH_INDEXED]
    want = {}
    for m in modes:
        for i, (p, ids) in enumerate(chunks):
            if m in (capi.MESH_PARITY, capi.MESH_ORDERED):
                v, ix = orc.mesh_chunk(p, ids, uv_default), np.empty(0, np.uint32)
            else:
                v, ix, _ = orc.mesh_chunk_ext(p, ids, uv_default, greedy=bool(m & capi.MESH_GREEDY), indexed=bool(m & capi.MESH_INDEXED))
            want[(m, i)] = (v.shape[0], ix.shape[0], hashlib.sha256(v.tobytes() + ix.tobytes()).digest())
    with capi.Context(0, 4096, 1 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        slots = ctx.load_chunks([p for p, _ in chunks], np.stack([ids for _, ids in chunks]))
        for rep in range(int(os.environ.get("DSP_STRESS_REPS", "12"))):
            for m in modes:
                # many copies of the list in random order: enough chunks that every resident CTA takes several tickets
                order = rng.permutation(np.tile(np.arange(len(chunks)), int(rng.integers(8, 40))))
                entries, total = ctx.mesh_chunks(slots[order], m)
                lo = 0
                blob = ctx.download_arena(0, total)
                for j, e in zip(order, entries):
                    nv, ni, digest = want[(m, int(j))]
                    assert (int(e["vertex_count"]), int(e["index_count"])) == (nv, ni), (rep, m, j)
                    o = int(e["offset_bytes"])
                    oi = o + ((nv * 20 + 127) & ~127)
                    got = hashlib.sha256(blob[o:o + nv * 20].tobytes() + blob[oi:oi + ni * 4].tobytes()).digest()
                    assert got == digest, (rep, m, int(j), chunks[int(j)][0])


def test_repeated_shuffled_batches_are_stable_with_cross_chunk_culling(uv_default):
    rng = np.random.default_rng(7)
    dirs = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]
    with capi.Context(0, 4096, 1 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        worlds = []
        for origin, dims, kind, seed in (((0, 6, 0), (3, 3, 3), capi.GEN_NOISE, 0xD5E50001), ((20, 0, 0), (3, 2, 2), capi.GEN_RANDOM, 9),
                                         ((40, -2, 0), (2, 3, 2), capi.GEN_REFERENCE, 0)):
            first = ctx.generate_region(origin, dims, kind, seed)
            for c in range(dims[2]):
                for b in range(dims[1]):
                    for a in range(dims[0]):
                        p = (origin[0] + a, origin[1] + b, origin[2] + c)
                        worlds.append((p, first + a + dims[0] * (b + dims[1] * c), orc.generate(kind, seed, p)))
        ids_of = {p: ids for p, _, ids in worlds}
        slots = np.array([s for _, s, _ in worlds], dtype=np.uint32)
        want = {}
        for m in (capi.MESH_HALO, capi.MESH_HALO | capi.MESH_GREEDY | capi.MESH_INDEXED):
            for i, (p, _, ids) in enumerate(worlds):
                slabs, present = np.zeros((6, 256), dtype=np.uint16), np.zeros(6, dtype=np.uint8)
                for f, d in enumerate(dirs):
                    q = (p[0] + d[0], p[1] + d[1], p[2] + d[2])
                    if q in ids_of:
                        v = ids_of[q].reshape(16, 16, 16)
                        present[f] = 1
                        slabs[f] = [v[:, 0, :], v[:, 15, :], v[15, :, :], v[0, :, :], v[:, :, 15], v[:, :, 0]][f].reshape(256)
                g = bool(m & capi.MESH_GREEDY)
                v, ix, _ = orc.mesh_chunk_ext(p, ids, uv_default, greedy=g, indexed=g, nbr_slabs=slabs, nbr_present=present)
                want[(m, i)] = (v.shape[0], ix.shape[0], hashlib.sha256(v.tobytes() + ix.tobytes()).digest())
        for rep in range(int(os.environ.get("DSP_STRESS_REPS", "12"))):
            for m in (capi.MESH_HALO, capi.MESH_HALO | capi.MESH_GREEDY | capi.MESH_INDEXED):
                order = rng.permutation(np.tile(np.arange(len(worlds)), int(rng.integers(8, 30))))
                entries, total = ctx.mesh_chunks(slots[order], m)
                blob = ctx.download_arena(0, total)
                for j, e in zip(order, entries):
                    nv, ni, digest = want[(m, int(j))]
                    assert (int(e["vertex_count"]), int(e["index_count"])) == (nv, ni), (rep, m, j)
                    o = int(e["offset_bytes"])
                    oi = o + ((nv * 20 + 127) & ~127)
                    assert hashlib.sha256(blob[o:o + nv * 20].tobytes() + blob[oi:oi + ni * 4].tobytes()).digest() == digest, (rep, m, worlds[int(j)][0])


def test_repeated_shuffled_batches_are_stable_on_generated_regions(uv_default):
    """The chunk-summary paths of the parity kernel under the same treatment: generated boxes (so summaries exist) that mix
    all-air chunks (finished at the ticket), chunks of one solid id (tile rebuilt in shared memory behind a byte-less barrier
    phase) and surface chunks (tile loaded), in shuffled orders with repeats -- a tile load landing in a buffer that was just
    written by hand, or a ticket skipped twice, would show up as an intermittent mismatch.  Default and ordered mode."""
    rng = np.random.default_rng(31)
    with capi.Context(0, 4096, 2 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        worlds = []
        for origin, dims, kind, seed in (((0, 2, 0), (2, 12, 2), capi.GEN_NOISE, 0xD5E50001), ((40, -2, 0), (2, 4, 2), capi.GEN_REFERENCE, 0)):
            first = ctx.generate_region(origin, dims, kind, seed)
            for c in range(dims[2]):
                for b in range(dims[1]):
                    for a in range(dims[0]):
                        p = (origin[0] + a, origin[1] + b, origin[2] + c)
                        worlds.append((p, first + a + dims[0] * (b + dims[1] * c), orc.generate(kind, seed, p)))
        slots = np.array([s for _, s, _ in worlds], dtype=np.uint32)
        want = []
        kinds = {"air": 0, "uniform": 0, "mixed": 0}
        for p, _, ids in worlds:
            v = orc.mesh_chunk(p, ids, uv_default)
            want.append((v.shape[0], hashlib.sha256(v.tobytes()).digest()))
            kinds["air" if not ids.any() else ("uniform" if (ids == ids[0]).all() else "mixed")] += 1
        assert min(kinds.values()) >= 4, kinds
        for rep in range(int(os.environ.get("DSP_STRESS_REPS", "12"))):
            for m in (capi.MESH_PARITY, capi.MESH_ORDERED):
                order = rng.permutation(np.tile(np.arange(len(worlds)), int(rng.integers(8, 40))))
                entries, total = ctx.mesh_chunks(slots[order], m)
                blob = ctx.download_arena(0, total)
                for j, e in zip(order, entries):
                    nv, digest = want[int(j)]
                    assert int(e["vertex_count"]) == nv, (rep, m, j)
                    o = int(e["offset_bytes"])
                    assert hashlib.sha256(blob[o:o + nv * 20].tobytes()).digest() == digest, (rep, m, worlds[int(j)][0])
