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
