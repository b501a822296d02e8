"""The consumer side of the path (chunk_mesh.rs:111-124 hands the mesh to a Vulkan vertex buffer):
(a) the vertex arena exported as a file descriptor and imported by ANOTHER PROCESS, which must see the same bytes;
(b) dsp_mesh_host_chunks_to_host: host voxels in, vertices in host memory out, three-stream pipeline."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from oracle import loader as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0xD5E50001

CHILD = r"""
import hashlib, sys
sys.path.insert(0, sys.argv[1])
from despawnengine_b200 import capi
fd, alloc, nbytes = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
imp = capi.ArenaImport(0, fd, alloc)
data = imp.read(0, nbytes)
print(hashlib.sha256(data.tobytes()).hexdigest(), hex(imp.device_ptr))
imp.close()
"""


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def test_exported_arena_is_the_same_memory_in_another_process(uv_default, tmp_path):
    with capi.Context(0, 512, 96 << 20, exportable_arena=True) as ctx:
        ctx.set_uv_table(uv_default)
        first = ctx.generate_region((0, 5, 0), (4, 6, 4), capi.GEN_NOISE, SEED)
        entries, total = ctx.mesh_range(first, 96)
        assert total > 0
        # the exportable (virtual-memory) arena behaves like the cudaMalloc one: byte-exact against the oracle
        pos = np.array([(a, 5 + b, c) for c in range(4) for b in range(6) for a in range(4)], dtype=np.int32)
        for i in (0, 17, 50, 95):
            want = orc.mesh_chunk(pos[i], orc.generate(orc.GEN_NOISE, SEED, pos[i]), uv_default)
            assert ctx.download_vertices(entries[i]).tobytes() == want.tobytes()
        mine = hashlib.sha256(ctx.download_arena(0, total).tobytes()).hexdigest()
        fd, alloc = ctx.arena_export_fd()
        assert fd >= 0 and alloc >= ctx.arena_bytes
        script = tmp_path / "importer.py"
        script.write_text(CHILD)
        r = subprocess.run([sys.executable, str(script), ROOT, str(fd), str(alloc), str(total)], pass_fds=(fd,), capture_output=True, text=True, timeout=300)
        os.close(fd)
        assert r.returncode == 0, r.stderr[-2000:]
        theirs, their_ptr = r.stdout.split()
        assert theirs == mine, "the importing process read different bytes"
        # in-process import as well: a second mapping of the same allocation
        fd2, alloc2 = ctx.arena_export_fd()
        imp = capi.ArenaImport(0, fd2, alloc2)
        os.close(fd2)
        assert imp.device_ptr != ctx.arena_device_ptr
        assert hashlib.sha256(imp.read(0, total).tobytes()).hexdigest() == mine
        e2, t2 = ctx.mesh_range(first, 96, capi.MESH_GREEDY | capi.MESH_INDEXED)  # the mapping follows what the kernels write next
        assert imp.read(0, t2).tobytes() == ctx.download_arena(0, t2).tobytes()
        imp.close()
    with capi.Context(0, 8, 1 << 20) as plain:
        with pytest.raises(capi.DspError) as ei:
            plain.arena_export_fd()
        assert ei.value.status == capi.ERR_BAD_ARG


@pytest.mark.parametrize("mode", [capi.MESH_PARITY, capi.MESH_GREEDY | capi.MESH_INDEXED, capi.MESH_INDEXED, capi.MESH_ORDERED])
def test_mesh_host_chunks_to_host_mirrors_the_arena(uv_default, mode):
    dims = (16, 10, 14)  # 2,240 chunks: several sub-batches
    pos = np.array([(a, 3 + b, c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)
    ids = orc.generate_batch(orc.GEN_NOISE, SEED, pos)
    with capi.Context(0, 4096, 1 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        off = 1 << 20
        entries, total, host = ctx.mesh_host_chunks_to_host(pos, ids, mode, arena_offset=off, capacity=600 << 20)
        assert host.shape[0] == total and host.tobytes() == ctx.download_arena(off, total).tobytes()
        greedy, indexed = bool(mode & capi.MESH_GREEDY), bool(mode & capi.MESH_INDEXED)
        for i in range(0, len(pos), 97):
            e = entries[i]
            o = int(e["offset_bytes"]) - off
            nv = int(e["vertex_count"]) * 20
            if greedy or indexed:
                v, idx, _ = orc.mesh_chunk_ext(pos[i], ids[i], uv_default, greedy=greedy, indexed=indexed)
                assert host[o:o + nv].tobytes() == v.tobytes()
                io = o + ((nv + 127) & ~127)
                assert host[io:io + 4 * int(e["index_count"])].tobytes() == idx.tobytes()
            else:
                assert host[o:o + nv].tobytes() == orc.mesh_chunk(pos[i], ids[i], uv_default).tobytes()
        # a host buffer that is too small is reported with the size needed, after copying what fits
        ctx.unload_chunks(np.arange(len(pos), dtype=np.uint32))
        with pytest.raises(capi.DspError) as ei:
            ctx.mesh_host_chunks_to_host(pos, ids, mode, capacity=4096)
        assert ei.value.status == capi.ERR_BAD_ARG and str(total) in str(ei.value)


def test_mesh_host_chunks_to_host_pinned_destination_is_device_driven(uv_default):
    """Pinned host buffers: the sub-batch windows are stored by a copy kernel over PCIe, no host round trip between sub-batches;
    same bytes as the arena, for the reference layout and for merged indexed quads."""
    import torch
    dims = (16, 12, 16)  # 3,072 chunks
    pos = np.array([(a, 2 + b, c) for c in range(dims[2]) for b in range(dims[1]) for a in range(dims[0])], dtype=np.int32)
    n = len(pos)
    vox = torch.empty((n, 4096), dtype=torch.int16).pin_memory()
    vox.numpy().view(np.uint16)[:] = orc.generate_batch(orc.GEN_NOISE, SEED, pos)
    cap = 400 << 20
    host = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    entries = np.zeros(n, dtype=capi.MESH_ENTRY_DTYPE)
    with capi.Context(0, 4096, 1 << 30) as ctx:
        ctx.set_uv_table(uv_default)
        for mode in (capi.MESH_PARITY, capi.MESH_GREEDY | capi.MESH_INDEXED):
            launches0 = ctx.kernel_launches
            total = ctx.mesh_host_chunks_to_host_ptr(n, pos, vox.data_ptr(), entries, host.data_ptr(), cap, mode, arena_offset=256)
            assert ctx.kernel_launches - launches0 >= 4 * 3  # per sub-batch: positions, mesh, cursor, copy-out
            assert total > 0 and host.numpy()[:total].tobytes() == ctx.download_arena(256, total).tobytes()
            e = entries[n // 2]
            o = int(e["offset_bytes"]) - 256
            if mode == capi.MESH_PARITY:
                want = orc.mesh_chunk(pos[n // 2], vox.numpy().view(np.uint16)[n // 2], uv_default)
                assert host.numpy()[o:o + want.nbytes].tobytes() == want.tobytes()
            ctx.unload_chunks(np.arange(n, dtype=np.uint32))
