"""ctypes binding of the C ABI declared in ``include/despawn_b200.h``.

This is the only way Python reaches the GPU path; there is no torch extension and no fallback.
If the shared library is missing the import of ``lib()`` raises -- nothing silently degrades to a
CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSP_LIB_PATH") or os.path.join(_PKG, "lib", "libdespawn_b200.so")  # override: A/B builds

# dsp_status
OK, ERR_BAD_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_ARENA_FULL, ERR_STORE_FULL, ERR_NOT_RESIDENT, ERR_INTERNAL = range(8)
# dsp_gen_kind
GEN_REFERENCE, GEN_NOISE, GEN_RANDOM, GEN_CHECKER = 0, 1, 2, 3
# mesh modes
MESH_PARITY, MESH_HALO, MESH_ORDERED, MESH_GREEDY, MESH_INDEXED = 0, 1, 2, 4, 8
NO_SLOT = 0xFFFFFFFF
FACE_TOP, FACE_BOTTOM, FACE_REAR, FACE_FRONT, FACE_RIGHT, FACE_LEFT = range(6)
OPPOSITE_FACE = (1, 0, 3, 2, 5, 4)

SEAM_HANDLE_BYTES = 128
CREATE_EXPORTABLE_ARENA = 1
MESH_ENTRY_DTYPE = np.dtype([("offset_bytes", np.uint64), ("vertex_count", np.uint32), ("index_count", np.uint32)])
EDIT_DTYPE = np.dtype([("wx", np.int32), ("wy", np.int32), ("wz", np.int32), ("block", np.uint32)])
VERTEX_DTYPE = np.dtype([("position", np.float32, 3), ("tex_coords", np.float32, 2)])  # vertex.rs:5-12
SCENE_CONFIG_DTYPE = np.dtype([("horizontal_render_distance", np.uint32), ("vertical_render_distance", np.uint32), ("gen_kind", np.int32),
                               ("seed", np.uint32), ("mesh_mode", np.uint32), ("reserved", np.uint32)])
SCENE_STATS_DTYPE = np.dtype([("current_chunk", np.int32, 3), ("changed", np.uint32), ("n_visible", np.uint64), ("n_loaded_now", np.uint64),
                              ("n_unloaded_now", np.uint64), ("n_loaded_total", np.uint64), ("n_meshes", np.uint64), ("vertex_total", np.uint64),
                              ("index_total", np.uint64), ("arena_bytes_used", np.uint64)])
assert SCENE_CONFIG_DTYPE.itemsize == 24 and SCENE_STATS_DTYPE.itemsize == 80
assert MESH_ENTRY_DTYPE.itemsize == 16 and EDIT_DTYPE.itemsize == 16 and VERTEX_DTYPE.itemsize == 20

# Every symbol include/despawn_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _i32, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int
_P = C.POINTER
SYMBOLS = {
    "dsp_create": (_int, [_int, _u64, _u64, _P(_vp)]),
    "dsp_create_ex": (_int, [_int, _u64, _u64, _u32, _P(_vp)]),
    "dsp_destroy": (_int, [_vp]),
    "dsp_last_error": (C.c_char_p, [_vp]),
    "dsp_status_string": (C.c_char_p, [_int]),
    "dsp_abi_version": (_int, []),
    "dsp_set_uv_table": (_int, [_vp, _u32, _P(C.c_float)]),
    "dsp_generate_chunks": (_int, [_vp, _u64, _P(_i32), _int, _u32, _P(_u32)]),
    "dsp_regenerate_chunks": (_int, [_vp, _u64, _P(_i32), _int, _u32, _P(_u32)]),
    "dsp_generate_region": (_int, [_vp, _P(_i32), _P(_u32), _int, _u32, _P(_u32)]),
    "dsp_clone_region": (_int, [_vp, _u32, _P(_i32), _P(_u32)]),
    "dsp_invalidate_chunks": (_int, [_vp, _u64, _P(_u32)]),
    "dsp_load_chunks": (_int, [_vp, _u64, _P(_i32), _P(C.c_uint16), _P(C.c_uint16), _P(_u32), _P(_u32)]),
    "dsp_unload_chunks": (_int, [_vp, _u64, _P(_u32)]),
    "dsp_find_chunk": (_int, [_vp, _P(_i32), _P(_u32)]),
    "dsp_resident_chunks": (_u64, [_vp]),
    "dsp_mesh_chunks": (_int, [_vp, _u64, _P(_u32), _u32, _u64, _vp, _P(_u64)]),
    "dsp_mesh_range": (_int, [_vp, _u32, _u64, _u32, _u64, _vp, _P(_u64)]),
    "dsp_mesh_host_chunks": (_int, [_vp, _u64, _P(_i32), _P(C.c_uint16), _u32, _u64, _P(_u32), _vp, _P(_u64)]),
    "dsp_mesh_host_chunks_to_host": (_int, [_vp, _u64, _P(_i32), _P(C.c_uint16), _u32, _u64, _P(_u32), _vp, _P(_u64), _vp, _u64]),
    "dsp_mesh_host_chunks_ex": (_int, [_vp, _u64, _P(_i32), _P(C.c_uint16), _P(_u32), _u32, _u64, _P(_u32), _vp, _P(_u64), _vp, _u64]),
    "dsp_mesh_range_async": (_int, [_vp, _u32, _u64, _u32, _u64]),
    "dsp_mesh_chunks_async": (_int, [_vp, _u64, _P(_u32), _u32, _u64]),
    "dsp_sync": (_int, [_vp]),
    "dsp_mesh_result": (_int, [_vp, _vp, _P(_u64)]),
    "dsp_apply_edits": (_int, [_vp, _u64, _vp, _P(_u32), _u64, _P(_u64)]),
    "dsp_apply_edits_and_remesh": (_int, [_vp, _u64, _vp, _u32, _u64, _P(_u32), _u64, _P(_u64), _vp, _P(_u64)]),
    "dsp_get_block_world": (_int, [_vp, _i32, _i32, _i32, _P(C.c_uint16)]),
    "dsp_pack_border_slabs": (_int, [_vp, _u64, _P(_u32), _int, _vp]),
    "dsp_set_halo_slabs": (_int, [_vp, _u64, _P(_u32), _int, _vp]),
    "dsp_clear_halo_slabs": (_int, [_vp]),
    "dsp_seam_create": (_int, [_vp, _int, _u64, _P(_u32), _P(_u32), _vp]),
    "dsp_seam_connect": (_int, [_vp, _u32, _vp]),
    "dsp_seam_exchange_async": (_int, [_vp]),
    "dsp_seam_epoch": (_u32, [_vp]),
    "dsp_comm_unique_id": (_int, [_vp]),
    "dsp_comm_init": (_int, [_vp, _int, _int, _vp]),
    "dsp_seam_connect_nccl": (_int, [_vp, _u32, _int]),
    "dsp_visible_chunks": (_int, [_P(C.c_float), _u32, _u32, _u64, _P(_i32), _P(_u64)]),
    "dsp_scene_update": (_int, [_vp, _P(C.c_float), _vp, _vp]),
    "dsp_scene_draw_list": (_int, [_vp, _u64, _P(_i32), _vp, _P(_u64)]),
    "dsp_scene_reset": (_int, [_vp]),
    "dsp_download_arena": (_int, [_vp, _u64, _u64, _vp]),
    "dsp_download_chunk": (_int, [_vp, _u32, _P(C.c_uint16), _P(_i32)]),
    "dsp_arena_export_fd": (_int, [_vp, _P(_int), _P(_u64)]),
    "dsp_arena_import_fd": (_int, [_int, _int, _u64, _P(_vp), _P(_vp)]),
    "dsp_arena_import_read": (_int, [_vp, _u64, _u64, _vp]),
    "dsp_arena_import_close": (_int, [_vp]),
    "dsp_arena_device_ptr": (_vp, [_vp]),
    "dsp_arena_bytes": (_u64, [_vp]),
    "dsp_voxel_device_ptr": (_vp, [_vp]),
    "dsp_entries_device_ptr": (_vp, [_vp]),
    "dsp_stream": (_vp, [_vp]),
    "dsp_max_chunks": (_u64, [_vp]),
    "dsp_kernel_launches": (_u64, [_vp]),
    "dsp_host_upload_stats": (_int, [_vp, _P(_u64)]),
    "dsp_debug_pack_chunk": (_int, [_P(C.c_uint16), _P(C.c_uint8), _P(_u32), _int]),
    "dsp_debug_phase_cycles": (_int, [_vp, _P(_u64)]),
    "dsp_set_kernel_timing": (_int, [_vp, _int]),
    "dsp_mesh_kernel_stats": (_int, [_vp, _P(_u64), _P(C.c_double), _P(C.c_float)]),
}

PACK_FORMS = ("air", "uniform", "bits2", "bits4", "bits8", "raw")


def debug_pack_chunk(blocks: np.ndarray, portable: bool = False):
    """(descriptor, packed bytes of the form's size) of the host-side packer on one chunk (dsp_debug_pack_chunk; no GPU needed)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(4096)
    out = np.zeros(8192, dtype=np.uint8)
    desc = C.c_uint32(0)
    rc = lib().dsp_debug_pack_chunk(b.ctypes.data_as(C.POINTER(C.c_uint16)), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(desc), 1 if portable else 0)
    if rc != 0:
        raise RuntimeError(f"dsp_debug_pack_chunk failed: {rc}")
    size = (0, 0, 1024, 2048, 4096, 8192)[desc.value & 7]
    return int(desc.value), out[:size].copy()


_lib = None


def lib() -> C.CDLL:
    """Loads libdespawn_b200.so (built by ``despawnengine_b200.build``); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m despawnengine_b200.build` "
                "(there is no CPU fallback for the mesh path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.dsp_abi_version() != 1:
            raise RuntimeError("ABI version mismatch between capi.py and libdespawn_b200.so")
        _lib = L
    return _lib


def comm_unique_id() -> bytes:
    """A fresh NCCL unique id (128 bytes) for dsp_comm_init: made on one rank, handed to all."""
    buf = C.create_string_buffer(128)
    rc = lib().dsp_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc != OK:
        raise DspError(rc, lib().dsp_last_error(None).decode())
    return buf.raw


def visible_chunks(camera_pos, horizontal: int, vertical: int) -> np.ndarray:
    """get_all_chunk_pos_in_render (scene_game.rs:170-223) through the C ABI; host only, needs no GPU."""
    cam = np.ascontiguousarray(camera_pos, dtype=np.float32).reshape(3)
    n = C.c_uint64(0)
    rc = lib().dsp_visible_chunks(cam.ctypes.data_as(C.POINTER(C.c_float)), horizontal, vertical, 0, None, C.byref(n))
    if rc != OK:
        raise RuntimeError(f"dsp_visible_chunks: status {rc}")
    out = np.empty((int(n.value), 3), dtype=np.int32)
    rc = lib().dsp_visible_chunks(cam.ctypes.data_as(C.POINTER(C.c_float)), horizontal, vertical, out.shape[0],
                                  out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n))
    if rc != OK:
        raise RuntimeError(f"dsp_visible_chunks: status {rc}")
    return out


class ArenaImport:
    """CUDA-side importer of an exported arena (dsp_arena_import_fd): another process' view of the same HBM bytes."""

    def __init__(self, device: int, fd: int, allocation_bytes: int):
        self._h, ptr = C.c_void_p(), C.c_void_p()
        rc = lib().dsp_arena_import_fd(device, fd, allocation_bytes, C.byref(self._h), C.byref(ptr))
        if rc != OK:
            raise DspError(rc, lib().dsp_last_error(None).decode())
        self.device_ptr = int(ptr.value)

    def read(self, offset: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        rc = lib().dsp_arena_import_read(self._h, offset, nbytes, out.ctypes.data_as(C.c_void_p))
        if rc != OK:
            raise DspError(rc, lib().dsp_last_error(None).decode())
        return out

    def close(self):
        if self._h:
            lib().dsp_arena_import_close(self._h)
            self._h = C.c_void_p()


class DspError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[{status}: {lib().dsp_status_string(status).decode()}] {message}")
        self.status = status


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Context:
    """Thin RAII wrapper: one context per GPU, single caller (SURVEY.md section 8b 'Threading')."""

    def __init__(self, device: int = 0, max_chunks: int = 4096, arena_bytes: int = 1 << 30, exportable_arena: bool = False):
        self._h = C.c_void_p()
        rc = lib().dsp_create_ex(device, max_chunks, arena_bytes, CREATE_EXPORTABLE_ARENA if exportable_arena else 0, C.byref(self._h))
        if rc != OK:
            raise DspError(rc, lib().dsp_last_error(None).decode())
        self.device = device

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != OK:
            raise DspError(rc, lib().dsp_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().dsp_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- block_uvs -----------------------------------------------------------------------------------
    def set_uv_table(self, uv_rows):
        uv = np.ascontiguousarray(uv_rows, dtype=np.float32).reshape(-1, 4)
        self._check(lib().dsp_set_uv_table(self._h, uv.shape[0], _ptr(uv, C.c_float)))

    # -- residency -------------------------------------------------------------------------------------
    def generate_chunks(self, pos, kind: int = GEN_REFERENCE, seed: int = 0) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        slots = np.empty(pos.shape[0], dtype=np.uint32)
        self._check(lib().dsp_generate_chunks(self._h, pos.shape[0], _ptr(pos, C.c_int32), kind, seed & 0xFFFFFFFF, _ptr(slots, C.c_uint32)))
        return slots

    def regenerate_chunks(self, pos, kind: int = GEN_REFERENCE, seed: int = 0) -> np.ndarray:
        """dsp_generate_chunks that also runs the generator over resident positions (edits are discarded)."""
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        slots = np.empty(pos.shape[0], dtype=np.uint32)
        self._check(lib().dsp_regenerate_chunks(self._h, pos.shape[0], _ptr(pos, C.c_int32), kind, seed & 0xFFFFFFFF, _ptr(slots, C.c_uint32)))
        return slots

    def clone_region(self, src_first_slot: int, new_origin) -> int:
        o = np.ascontiguousarray(new_origin, dtype=np.int32).reshape(3)
        first = C.c_uint32(0)
        self._check(lib().dsp_clone_region(self._h, src_first_slot, _ptr(o, C.c_int32), C.byref(first)))
        return int(first.value)

    def invalidate_chunks(self, slots):
        s = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        self._check(lib().dsp_invalidate_chunks(self._h, s.shape[0], _ptr(s, C.c_uint32)))

    def generate_region(self, origin, dims, kind: int = GEN_NOISE, seed: int = 0) -> int:
        o = np.ascontiguousarray(origin, dtype=np.int32).reshape(3)
        d = np.ascontiguousarray(dims, dtype=np.uint32).reshape(3)
        first = C.c_uint32(0)
        self._check(lib().dsp_generate_region(self._h, _ptr(o, C.c_int32), _ptr(d, C.c_uint32), kind, seed & 0xFFFFFFFF, C.byref(first)))
        return int(first.value)

    def load_chunks(self, pos, blocks, palette_ids=None, palette_offsets=None) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        n = pos.shape[0]
        blocks = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(n, 4096)
        slots = np.empty(n, dtype=np.uint32)
        if palette_ids is not None:
            pal = np.ascontiguousarray(palette_ids, dtype=np.uint16).reshape(-1)
            off = np.ascontiguousarray(palette_offsets, dtype=np.uint32).reshape(n + 1)
            pp, po = _ptr(pal, C.c_uint16), _ptr(off, C.c_uint32)
        else:
            pp, po = None, None
        self._check(lib().dsp_load_chunks(self._h, n, _ptr(pos, C.c_int32), _ptr(blocks, C.c_uint16), pp, po, _ptr(slots, C.c_uint32)))
        return slots

    def load_chunks_ptr(self, n: int, pos: np.ndarray, blocks_host_ptr: int, slots_out: np.ndarray):
        """Raw-pointer form used by the bench (pinned host memory owned by the caller)."""
        self._check(lib().dsp_load_chunks(self._h, n, _ptr(pos, C.c_int32), C.cast(blocks_host_ptr, C.POINTER(C.c_uint16)), None, None,
                                          _ptr(slots_out, C.c_uint32)))

    def unload_chunks(self, slots):
        s = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        self._check(lib().dsp_unload_chunks(self._h, s.shape[0], _ptr(s, C.c_uint32)))

    def find_chunk(self, pos):
        p = np.ascontiguousarray(pos, dtype=np.int32).reshape(3)
        s = C.c_uint32(0)
        rc = lib().dsp_find_chunk(self._h, _ptr(p, C.c_int32), C.byref(s))
        return int(s.value) if rc == OK else None

    @property
    def max_chunks(self) -> int:
        return int(lib().dsp_max_chunks(self._h))

    @property
    def resident_chunks(self) -> int:
        return int(lib().dsp_resident_chunks(self._h))

    # -- meshing ---------------------------------------------------------------------------------------
    def mesh_chunks(self, slots, mode: int = MESH_PARITY, arena_offset: int = 0):
        s = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        entries = np.zeros(s.shape[0], dtype=MESH_ENTRY_DTYPE)
        total = C.c_uint64(0)
        self._check(lib().dsp_mesh_chunks(self._h, s.shape[0], _ptr(s, C.c_uint32), mode, arena_offset, entries.ctypes.data_as(C.c_void_p), C.byref(total)))
        return entries, int(total.value)

    def mesh_range(self, first_slot: int, n: int, mode: int = MESH_PARITY, arena_offset: int = 0, entries_out=None):
        entries = np.zeros(n, dtype=MESH_ENTRY_DTYPE) if entries_out is None else entries_out
        total = C.c_uint64(0)
        self._check(lib().dsp_mesh_range(self._h, first_slot, n, mode, arena_offset, entries.ctypes.data_as(C.c_void_p), C.byref(total)))
        return entries, int(total.value)

    def mesh_host_chunks(self, pos, blocks, mode: int = MESH_PARITY, arena_offset: int = 0):
        """Load host chunk data (global ids) and mesh it, pipelined; returns (slots, entries, total_bytes)."""
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        n = pos.shape[0]
        blocks = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(n, 4096)
        slots = np.empty(n, dtype=np.uint32)
        entries = np.zeros(n, dtype=MESH_ENTRY_DTYPE)
        total = C.c_uint64(0)
        self._check(lib().dsp_mesh_host_chunks(self._h, n, _ptr(pos, C.c_int32), _ptr(blocks, C.c_uint16), mode, arena_offset,
                                               _ptr(slots, C.c_uint32), entries.ctypes.data_as(C.c_void_p), C.byref(total)))
        return slots, entries, int(total.value)

    def mesh_host_chunks_ptr(self, n: int, pos: np.ndarray, blocks_host_ptr: int, entries_out: np.ndarray, mode: int = MESH_PARITY, arena_offset: int = 0) -> int:
        """Raw-pointer form (pinned host voxels owned by the caller); returns total bytes."""
        total = C.c_uint64(0)
        self._check(lib().dsp_mesh_host_chunks(self._h, n, _ptr(pos, C.c_int32), C.cast(blocks_host_ptr, C.POINTER(C.c_uint16)), mode, arena_offset,
                                               None, entries_out.ctypes.data_as(C.c_void_p), C.byref(total)))
        return int(total.value)

    def mesh_host_chunks_to_host_ptr(self, n: int, pos: np.ndarray, blocks_host_ptr: int, entries_out: np.ndarray, host_vertices_ptr: int,
                                     host_capacity: int, mode: int = MESH_PARITY, arena_offset: int = 0) -> int:
        """dsp_mesh_host_chunks_to_host with caller-owned (pinned) host buffers; returns total bytes."""
        total = C.c_uint64(0)
        self._check(lib().dsp_mesh_host_chunks_to_host(self._h, n, _ptr(pos, C.c_int32), C.cast(blocks_host_ptr, C.POINTER(C.c_uint16)), mode, arena_offset,
                                                       None, entries_out.ctypes.data_as(C.c_void_p), C.byref(total), C.c_void_p(host_vertices_ptr), host_capacity))
        return int(total.value)

    def mesh_host_chunks_ex_ptr(self, n: int, pos: np.ndarray, blocks_host_ptr: int, palette_lens, entries_out: np.ndarray, mode: int = MESH_PARITY,
                                arena_offset: int = 0, host_vertices_ptr: int = 0, host_capacity: int = 0, slots_out=None) -> int:
        """dsp_mesh_host_chunks_ex with caller-owned host buffers (pinned: the mesh kernel pulls the tiles itself); palette_lens: u32[n]
        or None; returns total bytes."""
        total = C.c_uint64(0)
        pl = None if palette_lens is None else _ptr(np.ascontiguousarray(palette_lens, dtype=np.uint32), C.c_uint32)
        self._check(lib().dsp_mesh_host_chunks_ex(self._h, n, _ptr(pos, C.c_int32), C.cast(blocks_host_ptr, C.POINTER(C.c_uint16)), pl, mode, arena_offset,
                                                  None if slots_out is None else _ptr(slots_out, C.c_uint32), entries_out.ctypes.data_as(C.c_void_p), C.byref(total),
                                                  C.c_void_p(host_vertices_ptr) if host_vertices_ptr else None, host_capacity))
        return int(total.value)

    def mesh_host_chunks_ex(self, pos, blocks, palette_lens=None, mode: int = MESH_PARITY, arena_offset: int = 0):
        """dsp_mesh_host_chunks_ex on a numpy array (pageable memory unless the caller pinned it); returns (slots, entries, total_bytes)."""
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        n = pos.shape[0]
        blocks = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(n, 4096)
        slots = np.empty(n, dtype=np.uint32)
        entries = np.zeros(n, dtype=MESH_ENTRY_DTYPE)
        total = self.mesh_host_chunks_ex_ptr(n, pos, blocks.ctypes.data, palette_lens, entries, mode, arena_offset, slots_out=slots)
        return slots, entries, total

    def mesh_host_chunks_to_host(self, pos, blocks, mode: int = MESH_PARITY, arena_offset: int = 0, capacity: int = 0):
        """Returns (entries, total bytes, host mirror of the arena window as a u8 array)."""
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        n = pos.shape[0]
        blocks = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(n, 4096)
        entries = np.zeros(n, dtype=MESH_ENTRY_DTYPE)
        host = np.empty(capacity or n * 1474560, dtype=np.uint8)
        total = C.c_uint64(0)
        self._check(lib().dsp_mesh_host_chunks_to_host(self._h, n, _ptr(pos, C.c_int32), _ptr(blocks, C.c_uint16), mode, arena_offset, None,
                                                       entries.ctypes.data_as(C.c_void_p), C.byref(total), host.ctypes.data_as(C.c_void_p), host.shape[0]))
        return entries, int(total.value), host[: int(total.value)]

    def mesh_range_async(self, first_slot: int, n: int, mode: int = MESH_PARITY, arena_offset: int = 0):
        self._check(lib().dsp_mesh_range_async(self._h, first_slot, n, mode, arena_offset))

    def mesh_chunks_async(self, slots, mode: int = MESH_PARITY, arena_offset: int = 0):
        s = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        self._check(lib().dsp_mesh_chunks_async(self._h, s.shape[0], _ptr(s, C.c_uint32), mode, arena_offset))

    def sync(self):
        self._check(lib().dsp_sync(self._h))

    def mesh_result(self, n: int, entries_ptr: int | None = None):
        total = C.c_uint64(0)
        if entries_ptr is None:
            entries = np.zeros(n, dtype=MESH_ENTRY_DTYPE)
            self._check(lib().dsp_mesh_result(self._h, entries.ctypes.data_as(C.c_void_p), C.byref(total)))
            return entries, int(total.value)
        self._check(lib().dsp_mesh_result(self._h, C.c_void_p(entries_ptr), C.byref(total)))
        return None, int(total.value)

    # -- region seams ------------------------------------------------------------------------------------
    def pack_border_slabs(self, slots, face: int, dst_device_ptr: int):
        s = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        self._check(lib().dsp_pack_border_slabs(self._h, s.shape[0], _ptr(s, C.c_uint32), face, C.c_void_p(dst_device_ptr)))

    def set_halo_slabs(self, slots, face: int, src_device_ptr: int):
        s = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        self._check(lib().dsp_set_halo_slabs(self._h, s.shape[0], _ptr(s, C.c_uint32), face, C.c_void_p(src_device_ptr)))

    def clear_halo_slabs(self):
        self._check(lib().dsp_clear_halo_slabs(self._h))

    def seam_create(self, face: int, slots):
        """Returns (seam id, 128-byte handle to give to the neighbouring context / rank)."""
        s = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        seam = C.c_uint32(0)
        handle = C.create_string_buffer(SEAM_HANDLE_BYTES)
        self._check(lib().dsp_seam_create(self._h, face, s.shape[0], _ptr(s, C.c_uint32), C.byref(seam), C.cast(handle, C.c_void_p)))
        return int(seam.value), handle.raw

    def seam_connect(self, seam: int, peer_handle: bytes):
        assert len(peer_handle) == SEAM_HANDLE_BYTES
        buf = C.create_string_buffer(peer_handle, SEAM_HANDLE_BYTES)
        self._check(lib().dsp_seam_connect(self._h, seam, C.cast(buf, C.c_void_p)))

    def comm_init(self, rank: int, n_ranks: int, unique_id: bytes):
        """NCCL communicator for the fallback seam transport (collective over the ranks)."""
        buf = C.create_string_buffer(unique_id, 128)
        self._check(lib().dsp_comm_init(self._h, rank, n_ranks, C.cast(buf, C.c_void_p)))

    def seam_connect_nccl(self, seam: int, peer_rank: int):
        self._check(lib().dsp_seam_connect_nccl(self._h, seam, peer_rank))

    def seam_exchange_async(self):
        self._check(lib().dsp_seam_exchange_async(self._h))

    @property
    def seam_epoch(self) -> int:
        return int(lib().dsp_seam_epoch(self._h))

    # -- edits -----------------------------------------------------------------------------------------
    def apply_edits(self, edits) -> np.ndarray:
        e = np.ascontiguousarray(edits, dtype=EDIT_DTYPE).reshape(-1)
        dirty = np.empty(max(1, e.shape[0]), dtype=np.uint32)
        nd = C.c_uint64(0)
        self._check(lib().dsp_apply_edits(self._h, e.shape[0], e.ctypes.data_as(C.c_void_p), _ptr(dirty, C.c_uint32), dirty.shape[0], C.byref(nd)))
        return dirty[: int(nd.value)].copy()

    def get_block_world(self, wx: int, wy: int, wz: int):
        """Global block id at world voxel coordinates, or None when the chunk is not resident (world.rs:50-65)."""
        out = C.c_uint16(0)
        rc = lib().dsp_get_block_world(self._h, wx, wy, wz, C.byref(out))
        if rc == ERR_NOT_RESIDENT:
            return None
        self._check(rc)
        return int(out.value)

    def apply_edits_and_remesh(self, edits, mode: int = MESH_PARITY, arena_offset: int = 0):
        """Returns (dirty slots ascending, their mesh entries, total bytes)."""
        e = np.ascontiguousarray(edits, dtype=EDIT_DTYPE).reshape(-1)
        cap = min(4 * e.shape[0], self.max_chunks) if e.shape[0] else 0  # halo mode also dirties border neighbours
        dirty = np.empty(max(cap, 1), dtype=np.uint32)
        entries = np.empty(max(cap, 1), dtype=MESH_ENTRY_DTYPE)
        nd, total = C.c_uint64(0), C.c_uint64(0)
        self._check(lib().dsp_apply_edits_and_remesh(self._h, e.shape[0], e.ctypes.data_as(C.c_void_p), mode, arena_offset, _ptr(dirty, C.c_uint32),
                                                     dirty.shape[0], C.byref(nd), entries.ctypes.data_as(C.c_void_p), C.byref(total)))
        k = int(nd.value)
        return dirty[:k].copy(), entries[:k].copy(), int(total.value)

    # -- scene (GameScene::update, scene_game.rs:39-82) ---------------------------------------------------
    def scene_update(self, camera_pos, horizontal: int, vertical: int, gen_kind: int = GEN_REFERENCE, seed: int = 0, mesh_mode: int = MESH_PARITY):
        cam = np.ascontiguousarray(camera_pos, dtype=np.float32).reshape(3)
        cfg = np.zeros(1, dtype=SCENE_CONFIG_DTYPE)
        cfg["horizontal_render_distance"], cfg["vertical_render_distance"] = horizontal, vertical
        cfg["gen_kind"], cfg["seed"], cfg["mesh_mode"] = gen_kind, seed & 0xFFFFFFFF, mesh_mode
        stats = np.zeros(1, dtype=SCENE_STATS_DTYPE)
        self._check(lib().dsp_scene_update(self._h, _ptr(cam, C.c_float), cfg.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p)))
        return stats[0]

    def scene_draw_list(self):
        n = C.c_uint64(0)
        self._check(lib().dsp_scene_draw_list(self._h, 0, None, None, C.byref(n)))
        pos = np.empty((int(n.value), 3), dtype=np.int32)
        entries = np.empty(int(n.value), dtype=MESH_ENTRY_DTYPE)
        self._check(lib().dsp_scene_draw_list(self._h, pos.shape[0], _ptr(pos, C.c_int32), entries.ctypes.data_as(C.c_void_p), C.byref(n)))
        return pos, entries

    def scene_reset(self):
        self._check(lib().dsp_scene_reset(self._h))

    # -- data access -------------------------------------------------------------------------------------
    def download_arena(self, offset: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        if nbytes:
            self._check(lib().dsp_download_arena(self._h, offset, nbytes, out.ctypes.data_as(C.c_void_p)))
        return out

    def download_arena_into(self, offset: int, nbytes: int, host_ptr: int):
        self._check(lib().dsp_download_arena(self._h, offset, nbytes, C.c_void_p(host_ptr)))

    def download_vertices(self, entry) -> np.ndarray:
        """The vertex stream of one mesh entry as a structured array (empty <=> the reference's None)."""
        n = int(entry["vertex_count"])
        return self.download_arena(int(entry["offset_bytes"]), n * 20).view(VERTEX_DTYPE)

    def download_indices(self, entry) -> np.ndarray:
        """DSP_MESH_INDEXED: the chunk-local u32 indices of one mesh entry (they follow the vertices, 128-byte aligned)."""
        n = int(entry["index_count"])
        off = int(entry["offset_bytes"]) + ((int(entry["vertex_count"]) * 20 + 127) & ~127)
        return self.download_arena(off, n * 4).view(np.uint32)

    def download_chunk(self, slot: int):
        ids = np.empty(4096, dtype=np.uint16)
        pos = np.zeros(3, dtype=np.int32)
        self._check(lib().dsp_download_chunk(self._h, slot, _ptr(ids, C.c_uint16), _ptr(pos, C.c_int32)))
        return ids, pos

    def arena_export_fd(self):
        """(fd, allocation bytes) of an exportable arena: pass the fd to the importer (Vulkan external memory, another CUDA
        process) and close it afterwards."""
        fd, size = C.c_int(-1), C.c_uint64(0)
        self._check(lib().dsp_arena_export_fd(self._h, C.byref(fd), C.byref(size)))
        return int(fd.value), int(size.value)

    @property
    def arena_device_ptr(self) -> int:
        return int(lib().dsp_arena_device_ptr(self._h) or 0)

    @property
    def arena_bytes(self) -> int:
        return int(lib().dsp_arena_bytes(self._h))

    @property
    def voxel_device_ptr(self) -> int:
        return int(lib().dsp_voxel_device_ptr(self._h) or 0)

    @property
    def stream(self) -> int:
        return int(lib().dsp_stream(self._h) or 0)

    def host_upload_stats(self) -> dict:
        """How the chunks of the last host-chunk call crossed the bus (dsp_host_upload_stats)."""
        out = np.zeros(8, dtype=np.uint64)
        self._check(lib().dsp_host_upload_stats(self._h, _ptr(out, C.c_uint64)))
        keys = ("air", "uniform", "bits2", "bits4", "bits8", "raw", "host_threads", "device_read_bytes")
        return {k: int(v) for k, v in zip(keys, out)}

    def debug_phase_cycles(self) -> np.ndarray:
        out = np.zeros(32, dtype=np.uint64)
        self._check(lib().dsp_debug_phase_cycles(self._h, _ptr(out, C.c_uint64)))
        return out

    def set_kernel_timing(self, on: bool):
        self._check(lib().dsp_set_kernel_timing(self._h, 1 if on else 0))

    def mesh_kernel_stats(self):
        """(launch count, summed device ms, newest launch ms) of the mesh kernel alone."""
        n, tot, last = C.c_uint64(0), C.c_double(0), C.c_float(0)
        self._check(lib().dsp_mesh_kernel_stats(self._h, C.byref(n), C.byref(tot), C.byref(last)))
        return int(n.value), float(tot.value), float(last.value)

    @property
    def kernel_launches(self) -> int:
        return int(lib().dsp_kernel_launches(self._h))
