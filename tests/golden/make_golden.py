"""Writes tests/golden/oracle_digests.json: a freeze of the BUILDER-DEFINED terrain generators (noise, random,
checker: the reference has none of them) and of the oracle's meshes of those chunks, so that a later edit to the
oracle or to the terrain spec cannot silently drift.

These are NOT the reference pins.  The fixtures that say what the reference does -- refsrc_cases.npz,
refsrc_digests.json, flat_chunk_origin.bin -- are written by make_golden_from_reference.py from the reference's
parsed Rust sources, with no oracle involved.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import loader as orc  # noqa: E402

UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
CASES = [
    (orc.GEN_REFERENCE, 0, (0, 0, 0)), (orc.GEN_REFERENCE, 0, (0, -1, 0)), (orc.GEN_REFERENCE, 0, (0, 1, 0)),
    (orc.GEN_REFERENCE, 0, (-7, 0, 12)),
    (orc.GEN_NOISE, 0xD5E50001, (0, 7, 0)), (orc.GEN_NOISE, 0xD5E50001, (5, 8, 9)), (orc.GEN_NOISE, 0xD5E50001, (15, 6, 3)),
    (orc.GEN_NOISE, 0xD5E50001, (-2, 9, -5)),
    (orc.GEN_RANDOM, 0xD5E50003, (0, 0, 0)), (orc.GEN_RANDOM, 0xD5E50003, (3, -4, 5)),
    (orc.GEN_CHECKER, 0, (0, 0, 0)), (orc.GEN_CHECKER, 0, (1, 1, 1)),
]

here = os.path.dirname(os.path.abspath(__file__))
out = {"uv_table": UV.tolist(), "cases": []}
for kind, seed, pos in CASES:
    ids = orc.generate(kind, seed, pos)
    v = orc.mesh_chunk(pos, ids, UV)
    out["cases"].append({"kind": kind, "seed": seed, "pos": list(pos), "vertices": int(v.shape[0]),
                         "ids_sha256": hashlib.sha256(ids.tobytes()).hexdigest(),
                         "mesh_sha256": hashlib.sha256(v.tobytes()).hexdigest()})
json.dump(out, open(os.path.join(here, "oracle_digests.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases")
