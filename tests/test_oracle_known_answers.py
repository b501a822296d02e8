"""Pins for the CPU oracle.  The reference has no tests or golden vectors for this path (SURVEY.md
section 4), so the pins are (a) known answers derived by hand from the reference source (section 8c items
1-7) and (b) agreement between two independently written restatements (C++ walk vs numpy masks).
"""
import hashlib

import numpy as np
import pytest

from oracle import loader as orc
from oracle import numpy_oracle as npo

# Hand-typed from /root/reference/src/engine/rendering/cube.rs (NOT taken from the oracle tables):
# 4 corners per face as (px,py,pz,u,v); a face is corners (0,1,2,2,3,0).
BOTTOM = [(-.5, -.5, .5, 1, 1), (.5, -.5, .5, 0, 1), (.5, -.5, -.5, 0, 0), (-.5, -.5, -.5, 1, 0)]  # cube.rs:261-286
REAR = [(-.5, -.5, -.5, 0, 1), (.5, -.5, -.5, 1, 1), (.5, .5, -.5, 1, 0), (-.5, .5, -.5, 0, 0)]    # cube.rs:183-208
RIGHT = [(-.5, -.5, -.5, 0, 1), (-.5, .5, -.5, 0, 0), (-.5, .5, .5, 1, 0), (-.5, -.5, .5, 1, 1)]   # cube.rs:209-234
TOP = [(-.5, .5, -.5, 1, 1), (.5, .5, -.5, 0, 1), (.5, .5, .5, 0, 0), (-.5, .5, .5, 1, 0)]         # cube.rs:287-312


def expand(corners, off, uvrow):
    out = []
    for i in (0, 1, 2, 2, 3, 0):
        px, py, pz, u, v = corners[i]
        out.append((px + off[0], py + off[1], pz + off[2],
                    uvrow[0] + u * (uvrow[2] - uvrow[0]), uvrow[1] + v * (uvrow[3] - uvrow[1])))
    return out


def as_rows(verts):
    return np.concatenate([verts["position"], verts["tex_coords"]], axis=1)


def test_flat_chunk_counts(uv_default):
    ids = orc.generate(orc.GEN_REFERENCE, 0, (0, 0, 0))
    assert ids.reshape(16, 16, 16)[:, :8, :].all() and not ids.reshape(16, 16, 16)[:, 8:, :].any()
    v = orc.mesh_chunk((0, 0, 0), ids, uv_default)
    assert v.shape[0] == 1024 * 6 and v.nbytes == 122880  # 256 top + 256 bottom + 4*128 sides


def test_full_chunk_counts(uv_default):
    ids = orc.generate(orc.GEN_REFERENCE, 0, (3, -2, 5))
    assert (ids == 1).all()
    v = orc.mesh_chunk((3, -2, 5), ids, uv_default)
    assert v.shape[0] == 1536 * 6 and v.nbytes == 184320


def test_empty_chunk_is_none(uv_default):
    ids = orc.generate(orc.GEN_REFERENCE, 0, (0, 1, 0))
    assert not ids.any()
    assert orc.mesh_chunk((0, 1, 0), ids, uv_default).shape[0] == 0


def test_first_18_vertices_of_flat_chunk(uv_default):
    """idx 0 of the flat chunk at the origin emits exactly BOTTOM, REAR, RIGHT, in that order."""
    ids = orc.generate(orc.GEN_REFERENCE, 0, (0, 0, 0))
    v = as_rows(orc.mesh_chunk((0, 0, 0), ids, uv_default))
    want = np.array(expand(BOTTOM, (0, 0, 0), uv_default[1]) + expand(REAR, (0, 0, 0), uv_default[1])
                    + expand(RIGHT, (0, 0, 0), uv_default[1]), dtype=np.float32)
    assert np.array_equal(v[:18], want)
    assert tuple(v[0]) == (-0.5, -0.5, 0.5, 0.5, 1.0)  # first vertex: BOTTOM c0, uv = map_uv([1,1])
    # idx 1 (x=1,y=0,z=0) then emits BOTTOM, REAR only
    want2 = np.array(expand(BOTTOM, (1, 0, 0), uv_default[1]) + expand(REAR, (1, 0, 0), uv_default[1]), dtype=np.float32)
    assert np.array_equal(v[18:30], want2)


def test_first_top_face_of_flat_chunk(uv_default):
    """The first voxel with a TOP face is idx 7*16 = 112 (x=0,y=7,z=0): TOP, REAR, RIGHT."""
    ids = orc.generate(orc.GEN_REFERENCE, 0, (0, 0, 0))
    v = as_rows(orc.mesh_chunk((0, 0, 0), ids, uv_default))
    # faces emitted before idx 112: y=0 row: 16 bottom + 16 rear + right + left = 34; y=1..6 rows: 16 rear + 2 = 18 each
    before = 34 + 6 * 18
    want = np.array(expand(TOP, (0, 7, 0), uv_default[1]) + expand(REAR, (0, 7, 0), uv_default[1])
                    + expand(RIGHT, (0, 7, 0), uv_default[1]), dtype=np.float32)
    assert np.array_equal(v[before * 6: before * 6 + 18], want)


def test_checkerboard_is_the_maximum(uv_default):
    ids = orc.generate(orc.GEN_CHECKER, 0, (0, 0, 0))
    assert ids.sum() == 2048
    v = orc.mesh_chunk((0, 0, 0), ids, uv_default)
    assert v.shape[0] == 12288 * 6


def test_chunk_position_is_a_pure_translation(uv_default):
    ids = orc.generate(orc.GEN_RANDOM, 0xD5E50003, (0, 0, 0))
    a = orc.mesh_chunk((0, 0, 0), ids, uv_default)
    b = orc.mesh_chunk((5, -3, 9), ids, uv_default)
    assert np.array_equal(a["tex_coords"], b["tex_coords"])
    assert np.array_equal(a["position"] + np.float32(16) * np.array([5, -3, 9], dtype=np.float32), b["position"])


def visible_chunks(cam_chunk, hori, vert):
    """scene_game.rs:170-223 restated (cylinder test; loop bounds min-1 .. max+1 exclusive)."""
    out = []
    cx, cy, cz = cam_chunk
    for x in range(cx - hori - 1, cx + hori + 1):
        for y in range(cy - vert - 1, cy + vert + 1):
            for z in range(cz - hori - 1, cz + hori + 1):
                if abs(y - cy) <= vert and (x - cx) ** 2 + (z - cz) ** 2 <= hori ** 2:
                    out.append((x, y, z))
    return out


def test_default_settings_world_totals(uv_default):
    """settings.json5:4-5 (10 / 4) and the start camera (2.5, 22.5, -2.5) -> chunk (0,1,-1):
    2,853 chunks, 1,268 meshes, 10,712,064 vertices."""
    vis = visible_chunks((0, 1, -1), 10, 4)
    assert len(vis) == 2853
    pos = np.array(vis, dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_REFERENCE, 0, pos)
    counts, _ = orc.mesh_batch(pos, ids, uv_default, flavour=orc.FAIR, threads=4)
    assert int((counts > 0).sum()) == 1268
    assert int(counts.sum()) == 10712064


@pytest.mark.parametrize("kind,seed", [(orc.GEN_REFERENCE, 0), (orc.GEN_NOISE, 0xD5E50001), (orc.GEN_RANDOM, 0xD5E50003),
                                       (orc.GEN_CHECKER, 0)])
@pytest.mark.parametrize("pos", [(0, 0, 0), (1, 7, -2), (-3, -1, 4), (15, 9, 15)])
def test_cpp_and_numpy_restatements_agree(kind, seed, pos, uv_default):
    ids_c = orc.generate(kind, seed, pos)
    ids_n = npo.generate(kind, seed, pos)
    assert np.array_equal(ids_c, ids_n)
    vc = orc.mesh_chunk(pos, ids_c, uv_default, orc.FAITHFUL)
    vf = orc.mesh_chunk(pos, ids_c, uv_default, orc.FAIR)
    vn = npo.mesh_chunk(pos, ids_n, uv_default)
    assert vc.tobytes() == vf.tobytes() == vn.tobytes()


def test_uv_variants(uv_default):
    """Swapped atlas assignment (texture_atlas.rs:39 iterates a HashMap, so either can happen), the
    (0,0)-(1,1) fallback for a block missing from block_uvs (chunk_mesh.rs:53-56), and a 3-tile atlas
    whose thirds are not exactly representable (uv_min + 1*(uv_max-uv_min) != uv_max in general)."""
    ids = orc.generate(orc.GEN_RANDOM, 7, (2, 2, 2))
    swapped = uv_default[[0, 2, 1]]
    third = np.float32(1) / np.float32(3)
    odd = np.array([[0, 0, 0, 0], [third, 0, third + third, 1], [np.float32(2) * third, third, 1, np.float32(0.7)]], dtype=np.float32)
    only_dirt = uv_default[:2]  # block id 2 missing -> fallback
    for uv in (swapped, odd, only_dirt):
        a = orc.mesh_chunk((2, 2, 2), ids, uv, orc.FAITHFUL)
        b = npo.mesh_chunk((2, 2, 2), ids, uv)
        assert a.tobytes() == b.tobytes()
    a = orc.mesh_chunk((2, 2, 2), ids, only_dirt)
    sel = np.repeat(ids[ids != 0], 1)  # sanity: fallback block really gets u in {0,1}
    assert set(np.unique(a["tex_coords"][:, 0]).tolist()) <= {0.0, 0.5, 1.0} and sel.size > 0


def test_large_coordinates_round_like_f32(uv_default):
    """Positions go through f32: (chunk as f32)*16 then + local then + template (chunk_mesh.rs:61-62,67)."""
    ids = orc.generate(orc.GEN_CHECKER, 0, (0, 0, 0))
    pos = (2 ** 21 + 3, -(2 ** 22) - 1, 2 ** 27 + 5)
    a = orc.mesh_chunk(pos, ids, uv_default)
    b = npo.mesh_chunk(pos, ids, uv_default)
    assert a.tobytes() == b.tobytes()


def test_noise_terrain_shape():
    seed = 0xD5E50001
    hs = np.array([[orc.terrain_height(seed, x, z) for x in range(0, 256, 5)] for z in range(0, 256, 5)])
    assert hs.min() >= 16 and hs.max() <= 239 and hs.std() > 4
    assert np.array_equal(hs, npo.terrain_height(seed, *np.meshgrid(np.arange(0, 256, 5), np.arange(0, 256, 5))))
    # negative world coordinates use floor division for the lattice cell
    assert orc.terrain_height(seed, -1, -33) == int(npo.terrain_height(seed, np.array(-1), np.array(-33)))


def test_golden_digests(uv_default):
    """Committed digests (tests/golden/oracle_digests.json, written by tests/golden/make_golden.py)."""
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_digests.json")
    golden = json.load(open(path))
    for case in golden["cases"]:
        ids = orc.generate(case["kind"], case["seed"], case["pos"])
        v = orc.mesh_chunk(case["pos"], ids, uv_default)
        assert hashlib.sha256(ids.tobytes()).hexdigest() == case["ids_sha256"], case
        assert v.shape[0] == case["vertices"], case
        assert hashlib.sha256(v.tobytes()).hexdigest() == case["mesh_sha256"], case
