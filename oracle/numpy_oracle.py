"""TEST INFRASTRUCTURE ONLY -- second, independent CPU restatement (numpy, vectorised).

A deliberately different formulation of the same algorithm as ``mesher_oracle.cpp``: where the
C++ oracle walks voxels one at a time the way ``build_chunk_mesh`` does
(/root/reference/src/content/world/chunks/chunk_mesh.rs:32-107), this file computes the six
visibility masks for the whole chunk at once and gathers the vertex stream with array indexing.
Tests require the two to agree byte for byte, which is the strongest pin available while the
reference itself cannot be compiled here (no rustc).  PARITY UNPINNED by reference tests.

Only ``tests/`` may import this module.
"""
from __future__ import annotations

import numpy as np

CHUNK_SIZE = 16  # chunk.rs:7

# Face templates in emission order TOP, BOTTOM, REAR, FRONT, RIGHT(-X), LEFT(+X)
# (chunk_mesh.rs:65-106).  Each is 4 unique corners c0..c3 expanded as (c0,c1,c2,c2,c3,c0)
# (cube.rs:156-312).  Corner = (px, py, pz, u, v).
_CORNERS = {
    "TOP": [(-.5, .5, -.5, 1, 1), (.5, .5, -.5, 0, 1), (.5, .5, .5, 0, 0), (-.5, .5, .5, 1, 0)],          # cube.rs:287
    "BOTTOM": [(-.5, -.5, .5, 1, 1), (.5, -.5, .5, 0, 1), (.5, -.5, -.5, 0, 0), (-.5, -.5, -.5, 1, 0)],   # cube.rs:261
    "REAR": [(-.5, -.5, -.5, 0, 1), (.5, -.5, -.5, 1, 1), (.5, .5, -.5, 1, 0), (-.5, .5, -.5, 0, 0)],     # cube.rs:183
    "FRONT": [(-.5, -.5, .5, 1, 1), (-.5, .5, .5, 1, 0), (.5, .5, .5, 0, 0), (.5, -.5, .5, 0, 1)],        # cube.rs:156
    "RIGHT": [(-.5, -.5, -.5, 0, 1), (-.5, .5, -.5, 0, 0), (-.5, .5, .5, 1, 0), (-.5, -.5, .5, 1, 1)],    # cube.rs:209
    "LEFT": [(.5, -.5, -.5, 1, 1), (.5, -.5, .5, 0, 1), (.5, .5, .5, 0, 0), (.5, .5, -.5, 1, 0)],         # cube.rs:235
}
FACE_ORDER = ["TOP", "BOTTOM", "REAR", "FRONT", "RIGHT", "LEFT"]
_EXPAND = [0, 1, 2, 2, 3, 0]
# TEMPLATES[f, k] = (px, py, pz, u, v) of vertex k of face f
TEMPLATES = np.array([[_CORNERS[n][i] for i in _EXPAND] for n in FACE_ORDER], dtype=np.float32)  # (6,6,5)

VERTEX_DTYPE = np.dtype([("position", np.float32, 3), ("tex_coords", np.float32, 2)])  # vertex.rs:5-12
assert VERTEX_DTYPE.itemsize == 20


def face_masks(ids: np.ndarray) -> np.ndarray:
    """ids: (4096,) u16, index x + 16y + 256z.  Returns bool (6, 16z, 16y, 16x): face f of voxel emitted.

    chunk_mesh.rs:46-51 -- the neighbour across a chunk border always counts as air.
    """
    v = ids.reshape(16, 16, 16)  # [z][y][x]
    solid = v != 0
    air = np.pad(~solid, 1, constant_values=True)  # border = air
    c = slice(1, 17)
    m = np.empty((6, 16, 16, 16), dtype=bool)
    m[0] = solid & air[c, 2:18, c]   # top    (y+1)
    m[1] = solid & air[c, 0:16, c]   # bottom (y-1)
    m[2] = solid & air[0:16, c, c]   # rear   (z-1)
    m[3] = solid & air[2:18, c, c]   # front  (z+1)
    m[4] = solid & air[c, c, 0:16]   # right  (x-1)
    m[5] = solid & air[c, c, 2:18]   # left   (x+1)
    return m


def map_uv_table(uv: np.ndarray) -> np.ndarray:
    """uv: (n,4) f32 rows (min.x, min.y, max.x, max.y) -> (n, 2 orig, 2 comp) of
    uv_min + orig * (uv_max - uv_min) evaluated in f32 (chunk_mesh.rs:23-28)."""
    uv = np.asarray(uv, dtype=np.float32)
    lo, hi = uv[:, 0:2], uv[:, 2:4]
    d = (hi - lo).astype(np.float32)
    out = np.empty((uv.shape[0], 2, 2), dtype=np.float32)
    with np.errstate(invalid="ignore"):
        out[:, 0, :] = lo + (np.float32(0.0) * d).astype(np.float32)
        out[:, 1, :] = lo + (np.float32(1.0) * d).astype(np.float32)
    return out


def mesh_chunk(pos, ids: np.ndarray, uv: np.ndarray) -> np.ndarray:
    """Returns the vertex stream (structured array of VERTEX_DTYPE) the reference would upload,
    empty when it would return None."""
    ids = np.asarray(ids, dtype=np.uint16).reshape(4096)
    uv = np.asarray(uv, dtype=np.float32).reshape(-1, 4)
    m = face_masks(ids)                                   # (6, z, y, x)
    # stream order: idx ascending (z, y, x), then face order -> move face axis last and flatten
    em = np.moveaxis(m, 0, -1).reshape(4096, 6)
    vox, face = np.nonzero(em)                            # row-major => idx-major, face-minor
    n = vox.shape[0]
    out = np.zeros(n * 6, dtype=VERTEX_DTYPE)
    if n == 0:
        return out
    x = (vox % 16).astype(np.float32)
    y = ((vox // 16) % 16).astype(np.float32)
    z = (vox // 256).astype(np.float32)
    cpos = np.asarray(pos, dtype=np.int32).astype(np.float32) * np.float32(16.0)  # math.rs:135-139,166-170
    off = np.stack([x + cpos[0], y + cpos[1], z + cpos[2]], axis=1).astype(np.float32)  # chunk_mesh.rs:61-62
    t = TEMPLATES[face]                                   # (n, 6, 5)
    posn = (t[:, :, 0:3] + off[:, None, :]).astype(np.float32)   # v.position += pos_offset
    bid = ids[vox].astype(np.int64)
    table = np.concatenate([uv, np.array([[0, 0, 1, 1]], dtype=np.float32)], axis=0)  # fallback row
    bid = np.where(bid < uv.shape[0], bid, uv.shape[0])   # unknown block id -> (0,0)-(1,1) (chunk_mesh.rs:53-56)
    mapped = map_uv_table(table)[bid]                     # (n, orig, comp)
    orig = t[:, :, 3:5].astype(np.int64)                  # (n, 6, 2) in {0,1}
    u = np.take_along_axis(mapped[:, :, 0], orig[:, :, 0], axis=1)
    v = np.take_along_axis(mapped[:, :, 1], orig[:, :, 1], axis=1)
    out["position"] = posn.reshape(-1, 3)
    out["tex_coords"][:, 0] = u.reshape(-1)
    out["tex_coords"][:, 1] = v.reshape(-1)
    return out


# ---- builder-defined terrain spec (DESIGN.md "Terrain spec"), independent restatement ------------
_M32 = np.uint64(0xFFFFFFFF)


def _mix32(h):
    h = np.asarray(h, dtype=np.uint64) & _M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7FEB352D)) & _M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846CA68B)) & _M32
    h ^= h >> np.uint64(16)
    return h


def _u32(a):
    return np.asarray(a, dtype=np.int64).astype(np.uint64) & _M32


def hash2(seed, a, b):
    inner = _mix32((_u32(b) * np.uint64(0x85EBCA77) + np.uint64(0xC2B2AE3D)) & _M32)
    mid = _mix32(((_u32(a) * np.uint64(0x9E3779B1)) & _M32) ^ inner)
    return _mix32((np.uint64(seed) & _M32) ^ mid)


def hash3(seed, a, b, c):
    return _mix32((hash2(seed, a, b) + _u32(c) * np.uint64(0x27D4EB2F)) & _M32)


def _octave(seed, wx, wz, o):
    s = 5 - o
    wx = np.asarray(wx, dtype=np.int64)
    wz = np.asarray(wz, dtype=np.int64)
    ix, iz = wx >> s, wz >> s
    tx = ((wx & ((1 << s) - 1)) << (8 - s)).astype(np.uint64)
    tz = ((wz & ((1 << s) - 1)) << (8 - s)).astype(np.uint64)
    sx = (tx * tx * (np.uint64(768) - np.uint64(2) * tx)) >> np.uint64(16)
    sz = (tz * tz * (np.uint64(768) - np.uint64(2) * tz)) >> np.uint64(16)
    sd = (int(seed) + o) & 0xFFFFFFFF
    v00 = hash2(sd, ix, iz) & np.uint64(0xFF)
    v10 = hash2(sd, ix + 1, iz) & np.uint64(0xFF)
    v01 = hash2(sd, ix, iz + 1) & np.uint64(0xFF)
    v11 = hash2(sd, ix + 1, iz + 1) & np.uint64(0xFF)
    a = v00 * (np.uint64(256) - sx) + v10 * sx
    b = v01 * (np.uint64(256) - sx) + v11 * sx
    return (a * (np.uint64(256) - sz) + b * sz) >> np.uint64(16)


def terrain_height(seed, wx, wz):
    n = np.uint64(4) * _octave(seed, wx, wz, 0) + np.uint64(2) * _octave(seed, wx, wz, 1) + _octave(seed, wx, wz, 2)
    return (np.int64(16) + (n >> np.uint64(3)).astype(np.int64))


GEN_REFERENCE, GEN_NOISE, GEN_RANDOM, GEN_CHECKER = 0, 1, 2, 3


def generate(kind: int, seed: int, pos) -> np.ndarray:
    """(4096,) u16 global block ids for chunk ``pos``; index x + 16y + 256z."""
    z, y, x = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
    wx, wy, wz = pos[0] * 16 + x, pos[1] * 16 + y, pos[2] * 16 + z
    if kind == GEN_REFERENCE:  # world.rs:35-42 + chunk.rs:86-120
        if pos[1] > 0:
            ids = np.zeros((16, 16, 16), dtype=np.uint16)
        elif pos[1] == 0:
            ids = (y < 8).astype(np.uint16)
        else:
            ids = np.ones((16, 16, 16), dtype=np.uint16)
    elif kind == GEN_NOISE:
        h = terrain_height(seed, wx, wz)
        ids = np.where(wy < h - 1, 1, np.where(wy == h - 1, 2, 0)).astype(np.uint16)
    elif kind == GEN_RANDOM:
        r = hash3(seed, wx, wy, wz)
        ids = np.where((r & np.uint64(1)) != 0, np.uint64(1) + ((r >> np.uint64(1)) & np.uint64(1)), 0).astype(np.uint16)
    elif kind == GEN_CHECKER:
        ids = (((wx + wy + wz) & 1) == 0).astype(np.uint16)
    else:
        raise ValueError(kind)
    return ids.reshape(4096)


# ---- extension: greedy run-stack merge, bit-mask formulation -------------------------------------------------
# Independent of merge_quads() in mesher_oracle.cpp (which walks cells and compares runs cell by cell): here every
# row is a 16-bit mask and a run is a (start bit, end bit) pair, the formulation the CUDA kernel uses.  Tests
# require both to give the same quad records.  Spec: header comment of the extension block in mesher_oracle.cpp.
def _starts_ends(f: int, eq: int):
    """f: visible-face mask of a row along the run axis, eq: bit u = id[u] == id[u-1].  Returns (S, E) masks."""
    s = f & ~((f << 1) & eq) & 0xFFFF
    e = f & ~((f >> 1) & (eq >> 1)) & 0xFFFF
    return s, e


def _continued(f_a: int, eq_a: int, f_b: int, eq_b: int, eqv_b: int) -> int:
    """Start bits of row b whose run is identical (start, end, id) to a run of the previous row a."""
    s_a, e_a = _starts_ends(f_a, eq_a)
    s_b, e_b = _starts_ends(f_b, eq_b)
    cand = s_a & s_b & eqv_b
    diff = e_a ^ e_b
    out = 0
    while cand:
        u0 = (cand & -cand).bit_length() - 1
        cand &= cand - 1
        eb = ((e_b >> u0) & -(e_b >> u0)).bit_length() - 1  # run length - 1
        if ((diff >> u0) & ((2 << eb) - 1)) == 0:
            out |= 1 << u0
    return out


def greedy_records(ids: np.ndarray, masks: np.ndarray | None = None) -> np.ndarray:
    """Quad records anchor | dir << 12 | (w-1) << 15 | (h-1) << 19 in emission order (anchor voxel ascending, then
    direction).  masks: bool (6,16z,16y,16x) visible faces (default: the reference cull, face_masks)."""
    ids = np.asarray(ids, dtype=np.uint16).reshape(4096)
    if masks is None:
        masks = face_masks(ids)
    v = ids.reshape(256, 16).astype(np.int64)  # [row = y + 16 z][x]
    rows = range(256)

    def bits(b):  # bool[16] -> int, bit i = b[i]
        return int(sum(1 << i for i in range(16) if b[i]))

    f = [[bits(masks[d].reshape(256, 16)[r]) for r in rows] for d in range(6)]
    eqx = [bits(np.concatenate([[False], v[r, 1:] == v[r, :-1]])) for r in rows]
    eqy = [bits(v[r] == v[r - 1]) if (r & 15) > 0 else 0 for r in rows]
    eqz = [bits(v[r] == v[r - 16]) if r >= 16 else 0 for r in rows]
    recs = []
    # directions whose run axis is x: thread-per-row masks are used as they are
    for d, stride, eqv, has_prev in ((0, 16, eqz, lambda r: r >= 16), (1, 16, eqz, lambda r: r >= 16),
                                     (2, 1, eqy, lambda r: (r & 15) > 0), (3, 1, eqy, lambda r: (r & 15) > 0)):
        cont = [(_continued(f[d][r - stride], eqx[r - stride], f[d][r], eqx[r], eqv[r]) if has_prev(r) else 0) for r in rows]
        for r in rows:
            s, e = _starts_ends(f[d][r], eqx[r])
            anchors = s & ~cont[r]
            left = (15 - (r >> 4)) if stride == 16 else (15 - (r & 15))  # rows after this one along the stack axis
            for u0 in range(16):
                if not (anchors >> u0) & 1:
                    continue
                w = ((e >> u0) & -(e >> u0)).bit_length()
                h = 1
                while h <= left and (cont[r + h * stride] >> u0) & 1:
                    h += 1
                recs.append((r * 16 + u0) | (d << 12) | ((w - 1) << 15) | ((h - 1) << 19))
    # RIGHT / LEFT: run axis y, stack axis z -> transposed masks, column t = x + 16 z, bit = y
    def transpose(m):  # m[row = y + 16 z] bit x  ->  t[x + 16 z] bit y
        return [sum(((m[y + 16 * (t >> 4)] >> (t & 15)) & 1) << y for y in range(16)) for t in range(256)]
    eq_run, eq_stack = transpose(eqy), transpose(eqz)
    for d in (4, 5):
        ft = transpose(f[d])
        cont = [(_continued(ft[t - 16], eq_run[t - 16], ft[t], eq_run[t], eq_stack[t]) if t >= 16 else 0) for t in range(256)]
        for t in range(256):
            s, e = _starts_ends(ft[t], eq_run[t])
            anchors = s & ~cont[t]
            x, z = t & 15, t >> 4
            for u0 in range(16):
                if not (anchors >> u0) & 1:
                    continue
                w = ((e >> u0) & -(e >> u0)).bit_length()
                h = 1
                while z + h <= 15 and (cont[t + 16 * h] >> u0) & 1:
                    h += 1
                recs.append((x + 16 * u0 + 256 * z) | (d << 12) | ((w - 1) << 15) | ((h - 1) << 19))
    recs.sort(key=lambda q: ((q & 0xFFF) << 3) | ((q >> 12) & 7))
    return np.array(recs, dtype=np.uint32)


def expand_records_to_unit_faces(recs: np.ndarray) -> np.ndarray:
    """bool (6, 4096): the unit faces covered by the quads (each must be covered exactly once)."""
    axes = ((1, 0, 2), (1, 0, 2), (2, 0, 1), (2, 0, 1), (0, 1, 2), (0, 1, 2))  # normal, run, stack
    cover = np.zeros((6, 4096), dtype=np.int32)
    for q in np.asarray(recs, dtype=np.uint32).tolist():
        idx, d, w, h = q & 0xFFF, (q >> 12) & 7, ((q >> 15) & 15) + 1, ((q >> 19) & 15) + 1
        c0 = [idx & 15, (idx >> 4) & 15, idx >> 8]
        _, au, av = axes[d]
        for j in range(h):
            for i in range(w):
                c = list(c0)
                c[au] += i
                c[av] += j
                assert c[au] < 16 and c[av] < 16
                cover[d, c[0] + 16 * c[1] + 256 * c[2]] += 1
    assert cover.max(initial=0) <= 1, "a unit face is covered by two quads"
    return cover.astype(bool)
