"""The C++ host mirror (despawnengine_b200/host/despawn_world.hpp) over the C ABI: compiles everywhere, runs on the GPU."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_demo.cpp")


def build_demo(tmp_path):
    entry.build()
    exe = str(tmp_path / "host_mirror_demo")
    lib_dir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, SRC, "-L" + lib_dir, "-ldespawn_b200", "-Wl,-rpath," + lib_dir], check=True)
    return exe


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    exe = build_demo(tmp_path)
    assert os.path.exists(exe)
    import torch
    if not torch.cuda.is_available():  # without a GPU the mirror must surface the library's refusal, not compute anything
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_first_frame(tmp_path):
    """Default-settings world through World::load_chunk + build_chunk_mesh: 2,853 chunks, 1,268 meshes, 10,712,064
    vertices, and chunk (0,0,0) byte-identical to the committed golden stream."""
    exe = build_demo(tmp_path)
    dump = str(tmp_path / "origin.bin")
    r = subprocess.run([exe, dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].split() == ["2853", "1268", "10712064"]
    assert lines[1] == "72"  # two isolated voxels: 12 faces
    # GameScene::update: same first frame; one chunk to +x swaps 21 columns x 9 layers in and out, mesh count unchanged
    assert lines[2].split() == ["2853", "1268", "10712064", "189", "189", "1268"]
    assert lines[3].split() == ["6144", "9216", "0", "0"]  # build_all_host_chunks: flat 1,024 faces, full 1,536, never touched, emptied again
    golden = np.fromfile(os.path.join(ROOT, "tests", "golden", "flat_chunk_origin.bin"), dtype=np.uint8)
    assert np.fromfile(dump, dtype=np.uint8).tobytes() == golden.tobytes()
