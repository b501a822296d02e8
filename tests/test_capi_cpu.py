"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly the symbols
include/despawn_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "despawn_b200.h")


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsp_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_declare_the_same_symbols():
    assert declared_functions() == sorted(capi.SYMBOLS)


def test_integration_md_binds_every_declared_symbol():
    """INTEGRATION.md's Rust `extern "C"` block is the binding a maintainer of the reference would paste: it must name every entry point."""
    text = open(os.path.join(os.path.dirname(HEADER), "..", "INTEGRATION.md")).read()
    bound = set(re.findall(r"pub fn (dsp_[a-z_0-9]+)\(", text))
    assert bound == set(declared_functions()), sorted(set(declared_functions()) ^ bound)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    for name in declared_functions():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dsp_[a-z_0-9]+)", out))
    assert exported == set(declared_functions())


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_abi_version_and_status_strings():
    L = capi.lib()
    assert L.dsp_abi_version() == 1
    assert L.dsp_status_string(capi.OK) == b"ok"
    assert L.dsp_status_string(capi.ERR_ARENA_FULL) == b"vertex arena full"
    assert L.dsp_status_string(99) == b"unknown status"


def test_struct_layouts_match_the_header():
    assert capi.MESH_ENTRY_DTYPE.itemsize == 16 and capi.MESH_ENTRY_DTYPE.fields["vertex_count"][1] == 8
    assert capi.EDIT_DTYPE.itemsize == 16 and capi.VERTEX_DTYPE.itemsize == 20
    assert capi.VERTEX_DTYPE.fields["tex_coords"][1] == 12  # vertex.rs: R32G32B32 @0, R32G32 @12


def test_bad_arguments_are_reported_not_crashed():
    L = capi.lib()
    h = C.c_void_p()
    assert L.dsp_create(0, 0, 128, C.byref(h)) == capi.ERR_BAD_ARG          # max_chunks = 0
    assert L.dsp_create(0, 16, 100, C.byref(h)) == capi.ERR_BAD_ARG         # arena not a multiple of 128
    assert b"128" in L.dsp_last_error(None)
    assert L.dsp_destroy(None) == capi.OK
    assert L.dsp_sync(None) == capi.ERR_BAD_ARG


def test_no_gpu_means_an_error_not_a_fallback():
    """On a box without a CUDA device the product path must refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.DspError) as ei:
        capi.Context(0, 16, 1 << 20)
    assert ei.value.status == capi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "despawnengine_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "mesher_oracle" not in text.replace("oracle/mesher_oracle.cpp", ""), f


def test_registry_uv_table_and_host_chunk_palette():
    from despawnengine_b200.world import AIR_BLOCK_ID, AtlasUV, BlockRegistry, Chunk
    reg = BlockRegistry()
    t = reg.uv_table({"template:dirt": AtlasUV((0.0, 0.0), (0.5, 1.0))})
    assert np.array_equal(t[1], [0, 0, 0.5, 1]) and np.array_equal(t[2], [0, 0, 1, 1])  # engine missing -> fallback
    c = Chunk((0, 0, 0))
    c.set_block(1, 2, 3, "x:b")
    c.set_block(0, 0, 0, "x:a")
    c.set_block(1, 2, 3, AIR_BLOCK_ID)
    assert c.palette == [AIR_BLOCK_ID, "x:b", "x:a"] and c.blocks[0] == 2 and c.blocks[Chunk.index(1, 2, 3)] == 0
