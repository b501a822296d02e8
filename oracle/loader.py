"""TEST INFRASTRUCTURE ONLY -- ctypes loader for ``oracle/liboracle.so`` (the C++ restatement).

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs only.  Nothing under ``despawnengine_b200/`` imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

VERTEX_DTYPE = np.dtype([("position", np.float32, 3), ("tex_coords", np.float32, 2)])
GEN_REFERENCE, GEN_NOISE, GEN_RANDOM, GEN_CHECKER = 0, 1, 2, 3
FAITHFUL, FAIR = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mesher_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "liboracle.so"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i32p, u16p, f32p, i64p, u8p = (C.POINTER(t) for t in (C.c_int32, C.c_uint16, C.c_float, C.c_int64, C.c_uint8))
        L.orc_generate.argtypes = [C.c_int, C.c_uint32, i32p, u16p]
        L.orc_generate.restype = None
        L.orc_terrain_height.argtypes = [C.c_uint32, C.c_int32, C.c_int32]
        L.orc_terrain_height.restype = C.c_int32
        L.orc_mesh_chunk.argtypes = [i32p, u16p, f32p, C.c_uint32, C.c_int, C.c_void_p, C.c_int64]
        L.orc_mesh_chunk.restype = C.c_int64
        L.orc_mesh_chunk_halo.argtypes = [i32p, u16p, u16p, u8p, f32p, C.c_uint32, C.c_void_p, C.c_int64]
        L.orc_mesh_chunk_halo.restype = C.c_int64
        u32p = C.POINTER(C.c_uint32)
        L.orc_mesh_chunk_ext.argtypes = [i32p, u16p, u16p, u8p, f32p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                         u32p, C.c_int64, i64p, u32p, C.c_int64, i64p]
        L.orc_mesh_chunk_ext.restype = C.c_int64
        L.orc_mesh_batch.argtypes = [C.c_int64, i32p, u16p, f32p, C.c_uint32, C.c_int, C.c_int, i64p, C.c_void_p, C.c_int64]
        L.orc_mesh_batch.restype = C.c_int64
        L.orc_time_pipeline.argtypes = [C.c_int64, i32p, C.c_int, C.c_uint32, f32p, C.c_uint32, C.c_int, C.c_int, C.c_int, i64p]
        L.orc_time_pipeline.restype = C.c_double
        L.orc_workload_create.argtypes = [C.c_int64, i32p, C.c_int, C.c_uint32, f32p, C.c_uint32]
        L.orc_workload_create.restype = C.c_void_p
        L.orc_workload_destroy.argtypes = [C.c_void_p]
        L.orc_workload_destroy.restype = None
        L.orc_workload_mesh.argtypes = [C.c_void_p, C.c_int, C.c_int, i64p]
        L.orc_workload_mesh.restype = C.c_double
        L.orc_hardware_threads.argtypes = []
        L.orc_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def generate(kind: int, seed: int, pos) -> np.ndarray:
    p = np.ascontiguousarray(pos, dtype=np.int32)
    ids = np.empty(4096, dtype=np.uint16)
    lib().orc_generate(kind, seed & 0xFFFFFFFF, _p(p, C.c_int32), _p(ids, C.c_uint16))
    return ids


def generate_batch(kind: int, seed: int, pos: np.ndarray) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
    out = np.empty((pos.shape[0], 4096), dtype=np.uint16)
    L = lib()
    for i in range(pos.shape[0]):
        L.orc_generate(kind, seed & 0xFFFFFFFF, _p(pos[i], C.c_int32), _p(out[i], C.c_uint16))
    return out


def terrain_height(seed: int, wx: int, wz: int) -> int:
    return int(lib().orc_terrain_height(seed & 0xFFFFFFFF, wx, wz))


def mesh_chunk(pos, ids: np.ndarray, uv: np.ndarray, flavour: int = FAITHFUL) -> np.ndarray:
    p = np.ascontiguousarray(pos, dtype=np.int32)
    ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(4096)
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 4)
    out = np.empty(12288 * 6, dtype=VERTEX_DTYPE)
    n = lib().orc_mesh_chunk(_p(p, C.c_int32), _p(ids, C.c_uint16), _p(uv, C.c_float), uv.shape[0], flavour,
                             out.ctypes.data_as(C.c_void_p), out.shape[0])
    assert n >= 0
    return out[:n].copy()


def mesh_chunk_halo(pos, ids, nbr_slabs, nbr_present, uv) -> np.ndarray:
    p = np.ascontiguousarray(pos, dtype=np.int32)
    ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(4096)
    slabs = np.ascontiguousarray(nbr_slabs, dtype=np.uint16).reshape(6, 256)
    present = np.ascontiguousarray(nbr_present, dtype=np.uint8).reshape(6)
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 4)
    out = np.empty(12288 * 6, dtype=VERTEX_DTYPE)
    n = lib().orc_mesh_chunk_halo(_p(p, C.c_int32), _p(ids, C.c_uint16), _p(slabs, C.c_uint16), _p(present, C.c_uint8),
                                  _p(uv, C.c_float), uv.shape[0], out.ctypes.data_as(C.c_void_p), out.shape[0])
    assert n >= 0
    return out[:n].copy()


def mesh_chunk_ext(pos, ids, uv, greedy: bool = False, indexed: bool = False, nbr_slabs=None, nbr_present=None):
    """Extension modes (builder-defined, see mesher_oracle.cpp): returns (vertices, indices u32, quad records u32)."""
    p = np.ascontiguousarray(pos, dtype=np.int32)
    ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(4096)
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 4)
    slabs = present = None
    if nbr_slabs is not None:
        slabs = np.ascontiguousarray(nbr_slabs, dtype=np.uint16).reshape(6, 256)
        present = np.ascontiguousarray(nbr_present, dtype=np.uint8).reshape(6)
    verts = np.empty(12288 * 6, dtype=VERTEX_DTYPE)
    inds = np.empty(12288 * 6, dtype=np.uint32)
    quads = np.empty(12288, dtype=np.uint32)
    n_i, n_q = C.c_int64(0), C.c_int64(0)
    n = lib().orc_mesh_chunk_ext(_p(p, C.c_int32), _p(ids, C.c_uint16), None if slabs is None else _p(slabs, C.c_uint16),
                                 None if present is None else _p(present, C.c_uint8), _p(uv, C.c_float), uv.shape[0],
                                 1 if greedy else 0, 1 if indexed else 0, verts.ctypes.data_as(C.c_void_p), verts.shape[0],
                                 _p(inds, C.c_uint32), inds.shape[0], C.byref(n_i), _p(quads, C.c_uint32), quads.shape[0], C.byref(n_q))
    assert n >= 0
    return verts[:n].copy(), inds[:n_i.value].copy(), quads[:n_q.value].copy()


def mesh_batch(pos: np.ndarray, ids: np.ndarray, uv: np.ndarray, flavour: int = FAIR, threads: int = 0):
    """Returns (counts[n] int64 vertices per chunk, vertices packed in chunk order)."""
    pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
    n = pos.shape[0]
    ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(n, 4096)
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 4)
    if threads <= 0:
        threads = max(1, lib().orc_hardware_threads())
    counts = np.zeros(n, dtype=np.int64)
    cap = int(n) * 12288 * 6
    # first pass counts only (cheap relative to allocation of the worst case), second pass fills
    total = lib().orc_mesh_batch(n, _p(pos, C.c_int32), _p(ids, C.c_uint16), _p(uv, C.c_float), uv.shape[0], flavour,
                                 threads, _p(counts, C.c_int64), None, 0)
    out = np.empty(int(total), dtype=VERTEX_DTYPE)
    got = lib().orc_mesh_batch(n, _p(pos, C.c_int32), _p(ids, C.c_uint16), _p(uv, C.c_float), uv.shape[0], flavour,
                               threads, _p(counts, C.c_int64), out.ctypes.data_as(C.c_void_p), int(total))
    assert got == total and cap >= total
    return counts, out


def time_pipeline(pos: np.ndarray, kind: int, seed: int, uv: np.ndarray, flavour: int, threads: int,
                  include_generation: bool = True):
    """Times generation+mesh (or mesh only) of the given chunk positions on `threads` host threads.
    Returns (seconds, total_vertices)."""
    pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 4)
    tot = C.c_int64(0)
    s = lib().orc_time_pipeline(pos.shape[0], _p(pos, C.c_int32), kind, seed & 0xFFFFFFFF, _p(uv, C.c_float), uv.shape[0],
                                flavour, threads, 1 if include_generation else 0, C.byref(tot))
    return float(s), int(tot.value)


class Workload:
    """Chunks prepared once (outside timing); ``mesh`` times one pass of build_chunk_mesh over all of them."""

    def __init__(self, pos: np.ndarray, kind: int, seed: int, uv: np.ndarray):
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 4)
        self.n = pos.shape[0]
        self._h = lib().orc_workload_create(self.n, _p(pos, C.c_int32), kind, seed & 0xFFFFFFFF, _p(uv, C.c_float), uv.shape[0])

    def mesh(self, flavour: int = FAITHFUL, threads: int = 1):
        tot = C.c_int64(0)
        s = lib().orc_workload_mesh(self._h, flavour, threads, C.byref(tot))
        return float(s), int(tot.value)

    def close(self):
        if self._h:
            lib().orc_workload_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())
