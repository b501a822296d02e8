"""Writes tests/golden/oracle_digests.json and tests/golden/flat_chunk_origin.bin.

The reference is Rust and cannot be built or imported here (no rustc), so these goldens are
produced by the C++ restatement in oracle/ AFTER it has passed the hand-derived known answers in
tests/test_oracle_known_answers.py and agreed with the numpy restatement.  They freeze that state so
a later edit to the oracle (or to the terrain spec) cannot silently drift.

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
# one full small fixture: the reference's default chunk (config 1), 122,880 bytes
ids = orc.generate(orc.GEN_REFERENCE, 0, (0, 0, 0))
orc.mesh_chunk((0, 0, 0), ids, UV).tofile(os.path.join(here, "flat_chunk_origin.bin"))
print("wrote", len(out["cases"]), "cases")
