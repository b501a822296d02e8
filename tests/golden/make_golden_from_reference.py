"""Writes the committed fixtures of tests/golden/ FROM THE REFERENCE'S SOURCES.

    python tests/golden/make_golden_from_reference.py        (run where /root/reference exists)

oracle/reference_source.py parses the Rust sources under /root/reference/src (face tables of cube.rs, the emit
conditions / emission order / map_uv / pos_offset of chunk_mesh.rs, index() and the generators of chunk.rs, the
terrain rule of world.rs) and meshes chunks from the parsed text.  Neither the C++ oracle nor the CUDA library is
involved in producing these files: inputs are made here with numpy (or by the parsed terrain rule), outputs by the
parser's emitter.  The GPU box has no /root/reference; there these fixtures are what "the reference says".

  refsrc_cases.npz      pos (n,3) i32, ids (n,4096) u16 global block ids, uv (k,r,4) f32 tables (rows padded with NaN), uv_rows (k,)
  refsrc_digests.json   per case and uv table: vertex count and SHA-256 of the exact vertex stream; SHA-256 of every
                        parsed source file; what was parsed (templates, conditions, order)
  flat_chunk_origin.bin the full stream of the reference's default chunk (config 1), 122,880 bytes
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import reference_source as rs  # noqa: E402
from patterns import structured_chunks  # noqa: E402


def uv_tables():
    third = np.float32(1) / np.float32(3)
    default = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
    return {
        "default": default,
        "swapped": default[[0, 2, 1]],
        "short_table_fallback": default[:2],  # id 2 and above -> the source's fallback (0,0)-(1,1)
        "thirds": np.array([[0, 0, 0, 0], [third, 0, third + third, 1], [np.float32(2) * third, third, 1, np.float32(0.7)]], dtype=np.float32),
    }


def cases(ref):
    rng = np.random.default_rng(0xD5E5)
    z, y, x = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
    out = []

    def add(name, pos, ids):
        out.append((name, tuple(int(p) for p in pos), np.ascontiguousarray(ids, dtype=np.uint16).reshape(4096)))

    for pos in [(0, 0, 0), (0, -1, 0), (0, 1, 0), (-7, 0, 12), (2 ** 21 + 3, -(2 ** 22) - 1, 2 ** 27 - 5), (-(2 ** 27) + 1, 0, 7)]:
        add(f"reference_terrain{pos}", pos, ref.reference_terrain(pos))
    for p, hi in [(0.5, 3), (0.1, 3), (0.9, 3), (0.5, 10)]:
        ids = rng.integers(1, hi, size=4096) * (rng.random(4096) < p)
        add(f"bernoulli_p{p}_ids<{hi}", rng.integers(-50, 50, size=3), ids)
    add("checkerboard", (3, 3, 3), ((x + y + z) % 2 == 0).astype(np.uint16))
    add("checkerboard_inverse", (1, 1, 1), ((x + y + z) % 2 == 1).astype(np.uint16) * 2)
    for k, idx in enumerate([0, 4095, 8 + 16 * 8 + 256 * 8, 15, 15 * 16, 15 * 256]):
        ids = np.zeros(4096, dtype=np.uint16)
        ids[idx] = 1 + k % 2
        add(f"single_voxel_{idx}", (k + 1, -k, 2 * k), ids)
    for name, ids in structured_chunks().items():
        add("pattern_" + name, rng.integers(-9, 9, size=3), ids)
    for k in range(4):  # heightfield terrain with a surface block: solid below h(x,z), id 2 on top
        h = (6 + 5 * np.sin((x + 16 * k) * 0.37) + 4 * np.cos(z * 0.23 + k)).astype(np.int64)
        ids = np.where(y < h - 1, 1, np.where(y == h - 1, 2, 0))
        add(f"heightfield_{k}", (k + 30, 0, -k), ids)
    assert len({c[1] for c in out}) == len(out), "positions must be unique: the fixtures are loaded as one batch (one chunk per position)"
    return out


def main():
    ref = rs.parse()
    tables = uv_tables()
    cs = cases(ref)
    pos = np.array([c[1] for c in cs], dtype=np.int32)
    ids = np.stack([c[2] for c in cs])
    rows = max(t.shape[0] for t in tables.values())
    uv = np.full((len(tables), rows, 4), np.nan, dtype=np.float32)
    for k, t in enumerate(tables.values()):
        uv[k, :t.shape[0]] = t
    np.savez_compressed(os.path.join(HERE, "refsrc_cases.npz"), pos=pos, ids=ids, uv=uv,
                        uv_rows=np.array([t.shape[0] for t in tables.values()], dtype=np.int32))
    digests = {"generated_by": "tests/golden/make_golden_from_reference.py (oracle/reference_source.py: parsed Rust sources, numpy emitter)",
               "source_sha256": {p: hashlib.sha256(open(os.path.join(ref.root, p), "rb").read()).hexdigest() for p in rs.FILES.values()},
               "parsed": {"emission_order": [f for _, f in ref.emission], "air_conditions": ref.air_conditions, "index": ref.index_expr,
                          "coords": ref.coord_exprs, "map_uv": ref.map_uv_exprs, "pos_offset": ref.pos_offset_expr,
                          "terrain_rule": ref.terrain_rule, "templates": {k: v.tolist() for k, v in ref.templates.items()}},
               "uv_tables": list(tables), "cases": []}
    for name, p, v in cs:
        entry = {"name": name, "pos": list(p), "ids_sha256": hashlib.sha256(v.tobytes()).hexdigest(), "streams": {}}
        for tname, t in tables.items():
            s = ref.mesh_chunk(p, v, t)
            entry["streams"][tname] = {"vertices": int(s.shape[0]), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}
        digests["cases"].append(entry)
    json.dump(digests, open(os.path.join(HERE, "refsrc_digests.json"), "w"), indent=1)
    ref.mesh_chunk((0, 0, 0), ref.reference_terrain((0, 0, 0)), tables["default"]).tofile(os.path.join(HERE, "flat_chunk_origin.bin"))
    print("wrote", len(cs), "cases x", len(tables), "uv tables")


if __name__ == "__main__":
    main()
