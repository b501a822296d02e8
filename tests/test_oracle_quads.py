"""Pins for the greedy / indexed extension of the oracle (builder-defined spec, oracle/mesher_oracle.cpp).  The
reference has neither mode; what ties them to it: re-expanded quads == the reference's visible-face set, unmerged
quads == the reference's stream, indexed expansion == the stream.  Two independent formulations (C++ cell walk, Python
bit masks) must give the same quad records."""
import numpy as np
import pytest

from oracle import loader as orc
from oracle import numpy_oracle as npo

SEED_NOISE, SEED_RANDOM = 0xD5E50001, 0xD5E50003


def _cases():
    rng = np.random.default_rng(20261019)
    out = [((0, 0, 0), orc.generate(orc.GEN_REFERENCE, 0, (0, 0, 0))), ((0, -1, 0), orc.generate(orc.GEN_REFERENCE, 0, (0, -1, 0))),
           ((0, 1, 0), orc.generate(orc.GEN_REFERENCE, 0, (0, 1, 0))), ((5, 8, 9), orc.generate(orc.GEN_NOISE, SEED_NOISE, (5, 8, 9))),
           ((3, 6, 1), orc.generate(orc.GEN_NOISE, SEED_NOISE, (3, 6, 1))), ((1, 2, 3), orc.generate(orc.GEN_RANDOM, SEED_RANDOM, (1, 2, 3))),
           ((0, 0, 0), orc.generate(orc.GEN_CHECKER, 0, (0, 0, 0)))]
    for dens in (0.1, 0.6, 0.95):
        out.append(((1, -2, 3), ((rng.random(4096) < dens) * rng.integers(1, 4, 4096)).astype(np.uint16)))  # three ids + fallback id 3
        out.append(((-4, 0, 7), (rng.random(4096) < dens).astype(np.uint16)))
    return out


@pytest.mark.parametrize("case", range(len(_cases())))
def test_two_formulations_agree_and_quads_cover_the_reference_faces(case, uv_default):
    pos, ids = _cases()[case]
    stream, _, quads = orc.mesh_chunk_ext(pos, ids, uv_default, greedy=True)
    assert np.array_equal(quads, npo.greedy_records(ids))
    assert np.array_equal(npo.expand_records_to_unit_faces(quads), npo.face_masks(ids).reshape(6, 4096))
    assert stream.shape[0] == 6 * quads.shape[0]
    # emission order: anchor voxel ascending, then direction
    key = ((quads & 0xFFF).astype(np.int64) << 3) | ((quads >> 12) & 7)
    assert (np.diff(key) > 0).all()
    # unmerged + non-indexed == the reference restatement, byte for byte
    unit, _, unit_q = orc.mesh_chunk_ext(pos, ids, uv_default)
    assert unit.tobytes() == orc.mesh_chunk(pos, ids, uv_default).tobytes()
    assert ((unit_q >> 15) == 0).all()
    # indexed layouts expand to the streams
    for greedy, want in ((False, unit), (True, stream)):
        v, i, q = orc.mesh_chunk_ext(pos, ids, uv_default, greedy=greedy, indexed=True)
        assert v.shape[0] == 4 * q.shape[0] and i.shape[0] == 6 * q.shape[0]
        assert np.array_equal(i.reshape(-1, 6) - 4 * np.arange(q.shape[0], dtype=np.uint32)[:, None], np.tile([0, 1, 2, 2, 3, 0], (q.shape[0], 1)))
        assert v[i].tobytes() == want.tobytes()


def test_known_quad_counts(uv_default):
    # full chunk: one 16x16 quad per side; flat chunk (y < 8 solid): top/bottom 16x16, sides 16x8
    for pos in ((0, -1, 0), (0, 0, 0)):
        ids = orc.generate(orc.GEN_REFERENCE, 0, pos)
        _, _, q = orc.mesh_chunk_ext(pos, ids, uv_default, greedy=True)
        assert q.shape[0] == 6
        dims = {int((r >> 12) & 7): (int((r >> 15) & 15) + 1, int((r >> 19) & 15) + 1) for r in q}
        side_h = 16 if pos[1] < 0 else 8
        assert dims == {0: (16, 16), 1: (16, 16), 2: (16, side_h), 3: (16, side_h), 4: (side_h, 16), 5: (side_h, 16)}
    # checkerboard: nothing merges
    ids = orc.generate(orc.GEN_CHECKER, 0, (0, 0, 0))
    s, _, q = orc.mesh_chunk_ext((0, 0, 0), ids, uv_default, greedy=True)
    assert q.shape[0] == 12288 and s.tobytes() == orc.mesh_chunk((0, 0, 0), ids, uv_default).tobytes()


def test_merged_quad_corners_are_the_reference_vertices_of_the_corner_voxels(uv_default):
    """A 16x16 TOP quad of the flat chunk at (2,0,-3): its corners must equal the reference's TOP-face vertices of the
    four corner voxels (chunk_mesh.rs:61-62,67 arithmetic), uv = the template's mapped through map_uv."""
    pos = (2, 0, -3)
    ids = orc.generate(orc.GEN_REFERENCE, 0, pos)
    stream, _, q = orc.mesh_chunk_ext(pos, ids, uv_default, greedy=True)
    ref = orc.mesh_chunk(pos, ids, uv_default).reshape(-1, 6)  # one row per unit face, reference order
    k = int(np.nonzero(((q >> 12) & 7) == 0)[0][0])
    quad = stream.reshape(-1, 6)[k]
    # reference TOP faces: find by position of vertex 0 (c0 = -x, -z corner), vertex 1 (+x,-z), vertex 2 (+x,+z), vertex 4 (-x,+z)
    want_c0 = (2 * 16 - 0.5, 7.5, -3 * 16 - 0.5)
    want_c2 = (2 * 16 + 15.5, 7.5, -3 * 16 + 15.5)
    assert tuple(quad[0]["position"]) == want_c0 and tuple(quad[2]["position"]) == want_c2
    assert tuple(quad[1]["position"]) == (want_c2[0], 7.5, want_c0[2]) and tuple(quad[4]["position"]) == (want_c0[0], 7.5, want_c2[2])
    top_faces = [f for f in ref if f[0]["position"][1] == 7.5 and f[1]["position"][1] == 7.5 and f[2]["position"][1] == 7.5 and f[4]["position"][1] == 7.5]
    assert any(tuple(f[0]["position"]) == want_c0 and tuple(f[0]["tex_coords"]) == tuple(quad[0]["tex_coords"]) for f in top_faces)
    assert any(tuple(f[2]["position"]) == want_c2 and tuple(f[2]["tex_coords"]) == tuple(quad[2]["tex_coords"]) for f in top_faces)


def test_halo_masks_feed_the_merge(uv_default):
    pos = (0, 0, 0)
    ids = orc.generate(orc.GEN_REFERENCE, 0, (0, -1, 0))  # full chunk
    slabs = np.ones((6, 256), dtype=np.uint16)
    present = np.array([1, 0, 1, 0, 1, 0], dtype=np.uint8)  # solid neighbours above, behind (-z) and at -x
    _, _, q = orc.mesh_chunk_ext(pos, ids, uv_default, greedy=True, nbr_slabs=slabs, nbr_present=present)
    assert sorted(int((r >> 12) & 7) for r in q) == [1, 3, 5]


def test_structured_patterns_agree_between_formulations(uv_default):
    from patterns import structured_chunks
    for name, ids in structured_chunks().items():
        stream, _, quads = orc.mesh_chunk_ext((3, -1, 2), ids, uv_default, greedy=True)
        assert np.array_equal(quads, npo.greedy_records(ids)), name
        assert np.array_equal(npo.expand_records_to_unit_faces(quads), npo.face_masks(ids).reshape(6, 4096)), name
        unit = orc.mesh_chunk((3, -1, 2), ids, uv_default)
        assert quads.shape[0] <= unit.shape[0] // 6, name
    # a merge must never cross a block-id boundary: two ids side by side give two quads per long side
    _, _, q = orc.mesh_chunk_ext((0, 0, 0), structured_chunks()["two_halves"], uv_default, greedy=True)
    assert q.shape[0] == 2 + 4 * 2
