"""The mechanical pin of the oracle (and of the CUDA constant tables) to the reference's own SOURCES.

The reference is Rust (no rustc here) and ships no tests or vectors for this path, so the strongest pin available is
the source text: oracle/reference_source.py parses cube.rs / chunk_mesh.rs / chunk.rs / world.rs / math.rs under
/root/reference at test time and meshes chunks from the parsed tables, conditions and order -- it shares nothing with
oracle/mesher_oracle.cpp.  Here the oracle must reproduce those streams byte for byte, the committed fixtures must be
what the parser produces today, and the hand-typed corner codes of the CUDA kernels must equal the parsed templates.

The tests that need /root/reference skip where it does not exist (the GPU box); the fixture-based tests at the bottom
run everywhere.
"""
import hashlib
import json
import os
import re

import numpy as np
import pytest

from oracle import loader as orc
from oracle import numpy_oracle as npo
from oracle import reference_source as rs

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
needs_reference = pytest.mark.skipif(not rs.available(), reason="/root/reference is not present on this machine")


@pytest.fixture(scope="module")
def ref():
    return rs.parse()


@pytest.fixture(scope="module")
def fixtures():
    z = np.load(os.path.join(GOLDEN, "refsrc_cases.npz"))
    d = json.load(open(os.path.join(GOLDEN, "refsrc_digests.json")))
    tables = {name: z["uv"][k, :int(z["uv_rows"][k])].copy() for k, name in enumerate(d["uv_tables"])}
    return z["pos"], z["ids"], tables, d


# ---- what the parser read ------------------------------------------------------------------------------------------
@needs_reference
def test_parsed_structure(ref):
    assert ref.chunk_size == 16 and ref.max_chunk_index == ref.chunk_size ** 3 - 1
    assert len(ref.emission) == 6 and len(ref.templates) >= 6
    # every template is [c0, c1, c2, c2, c3, c0] over four distinct corners (what the indexed layout relies on)
    for _, name in ref.emission:
        t = ref.templates[name]
        assert np.array_equal(t[2], t[3]) and np.array_equal(t[0], t[5])
        assert len({tuple(r) for r in t.tolist()}) == 4
        assert set(np.abs(t[:, :3]).ravel().tolist()) == {0.5} and set(t[:, 3:].ravel().tolist()) <= {0.0, 1.0}


@needs_reference
def test_each_condition_looks_where_its_template_points(ref):
    """`right_air` tests idx - 1 and RIGHT_FACE lies in the plane x = -0.5: direction names and geometry are tied together by the
    source, not by anybody's reading of it."""
    for cond, face in ref.emission:
        expr = ref.air_conditions[cond]
        arg = re.search(r"is_air\((.*)\)\s*$", expr).group(1)
        delta = int(rs.rust_eval(arg, {"idx": 0, "CHUNK_SIZE": ref.chunk_size}))
        axis = {1: 0, ref.chunk_size: 1, ref.chunk_size ** 2: 2}[abs(delta)]
        assert abs(delta) == int(ref.index(*[1 if a == axis else 0 for a in range(3)]))
        t = ref.templates[face]
        assert (t[:, axis] == (0.5 if delta > 0 else -0.5)).all(), (cond, face)
        # and the border test of the same condition is on the same axis, at the end the neighbour lies beyond
        border = re.match(r"\s*(\w)\s*==\s*(.+?)\s*\|\|", expr)
        assert "xyz".index(border.group(1)) == axis
        assert int(rs.rust_eval(border.group(2), {"CHUNK_SIZE": ref.chunk_size})) == (ref.chunk_size - 1 if delta > 0 else 0)


# ---- the oracle against the parsed reference ---------------------------------------------------------------------
@needs_reference
def test_oracle_streams_equal_the_parsed_reference(ref, fixtures):
    pos, ids, tables, _ = fixtures
    for tname, uv in tables.items():
        for i in range(len(pos)):
            want = ref.mesh_chunk(pos[i], ids[i], uv).tobytes()
            assert orc.mesh_chunk(pos[i], ids[i], uv, flavour=orc.FAITHFUL).tobytes() == want, (i, tname, "faithful flavour")
            assert orc.mesh_chunk(pos[i], ids[i], uv, flavour=orc.FAIR).tobytes() == want, (i, tname, "fair flavour")
        assert npo.mesh_chunk(pos[0], ids[0], uv).tobytes() == ref.mesh_chunk(pos[0], ids[0], uv).tobytes()


@needs_reference
def test_oracle_streams_equal_the_parsed_reference_on_random_chunks(ref):
    rng = np.random.default_rng(20261020)
    uv = rng.random((6, 4)).astype(np.float32)
    for k in range(40):
        p = rng.integers(-2 ** 20, 2 ** 20, size=3).astype(np.int32)
        ids = (rng.integers(1, 9, size=4096) * (rng.random(4096) < rng.random())).astype(np.uint16)
        assert orc.mesh_chunk(p, ids, uv).tobytes() == ref.mesh_chunk(p, ids, uv).tobytes(), k


@needs_reference
def test_oracle_terrain_equals_the_parsed_world_rule(ref):
    for y in range(-3, 4):
        for x, z in [(0, 0), (-5, 9), (2 ** 20, -7)]:
            assert np.array_equal(orc.generate(orc.GEN_REFERENCE, 0, (x, y, z)), ref.reference_terrain((x, y, z))), (x, y, z)


@needs_reference
def test_committed_fixtures_are_what_the_parser_produces_today(ref, fixtures):
    pos, ids, tables, d = fixtures
    for p, want in d["source_sha256"].items():
        assert hashlib.sha256(open(os.path.join(ref.root, p), "rb").read()).hexdigest() == want, f"{p} changed: rerun make_golden_from_reference.py"
    assert d["parsed"]["emission_order"] == [f for _, f in ref.emission]
    for i, case in enumerate(d["cases"]):
        assert hashlib.sha256(ids[i].tobytes()).hexdigest() == case["ids_sha256"]
        for tname, uv in tables.items():
            s = ref.mesh_chunk(pos[i], ids[i], uv)
            assert case["streams"][tname] == {"vertices": int(s.shape[0]), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}, (case["name"], tname)
    flat = np.fromfile(os.path.join(GOLDEN, "flat_chunk_origin.bin"), dtype=np.uint8)
    assert flat.tobytes() == ref.mesh_chunk((0, 0, 0), ref.reference_terrain((0, 0, 0)), tables["default"]).tobytes()


# ---- the CUDA constant table against the parsed reference ----------------------------------------------------------
@needs_reference
def test_cuda_corner_codes_equal_the_parsed_templates(ref):
    """mesh_kernels.cuh encodes each face as four corners `corner(sx, sy, sz, u, v)`; they must be the four distinct
    corners c0, c1, c2, c3 of the parsed template, and the packing order of kCodes0/1 must be the parsed emission order."""
    src = open(os.path.join(os.path.dirname(HERE), "despawnengine_b200", "csrc", "mesh_kernels.cuh")).read()
    names = {"TOP_FACE": "kCodeTop", "BOTTOM_FACE": "kCodeBottom", "REAR_FACE": "kCodeRear", "FRONT_FACE": "kCodeFront", "RIGHT_FACE": "kCodeRight", "LEFT_FACE": "kCodeLeft"}
    for face, cname in names.items():
        m = re.search(r"constexpr uint32_t " + cname + r" = face_code\((.*?)\);", src)
        corners = re.findall(r"corner\(\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(\d)\s*,\s*(\d)\s*\)", m.group(1))
        assert len(corners) == 4
        t = ref.templates[face]
        for k, row in zip((0, 1, 2, 4), corners):
            sx, sy, sz, u, v = (int(c) for c in row)
            assert [0.5 * sx, 0.5 * sy, 0.5 * sz, float(u), float(v)] == t[k].tolist(), (face, k)
    order = [names[f] for _, f in ref.emission]
    m0 = re.search(r"kCodes0 = static_cast<uint64_t>\((\w+)\) \| \(static_cast<uint64_t>\((\w+)\) << 20\) \| \(static_cast<uint64_t>\((\w+)\) << 40\)", src)
    m1 = re.search(r"kCodes1 = static_cast<uint64_t>\((\w+)\) \| \(static_cast<uint64_t>\((\w+)\) << 20\) \| \(static_cast<uint64_t>\((\w+)\) << 40\)", src)
    assert list(m0.groups()) + list(m1.groups()) == order


# ---- fixtures only: runs without /root/reference (e.g. on the GPU box) ---------------------------------------------
def test_oracle_reproduces_the_committed_reference_fixtures(fixtures):
    pos, ids, tables, d = fixtures
    assert "parsed Rust sources" in d["generated_by"]
    for i, case in enumerate(d["cases"]):
        for tname, uv in tables.items():
            s = orc.mesh_chunk(pos[i], ids[i], uv)
            assert case["streams"][tname] == {"vertices": int(s.shape[0]), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}, (case["name"], tname)
    flat = np.fromfile(os.path.join(GOLDEN, "flat_chunk_origin.bin"), dtype=np.uint8)
    assert flat.tobytes() == orc.mesh_chunk((0, 0, 0), orc.generate(orc.GEN_REFERENCE, 0, (0, 0, 0)), tables["default"]).tobytes()
