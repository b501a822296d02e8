"""Region seams inside the library (dsp_seam_*): peer-memory exchange + in-kernel wait, against a single context that holds the
whole world.  The C program drives two contexts through the C ABI only; the Python tests cover the slab partition helpers
and the error paths.  Multi-PROCESS seams (CUDA IPC, one rank per GPU) are exercised by tools/world_bench.py --seam c under
`gpurun --gpus 2` and by bench.py --gpus N (cfg5)."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from despawnengine_b200 import capi
from despawnengine_b200 import partition as part
from despawnengine_b200.halo import RegionOnDevice

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0xD5E50001


@pytest.fixture(scope="module", autouse=True)
def _built():
    entry.build()


def test_c_program_drives_two_contexts_through_the_c_abi_only(tmp_path):
    exe = str(tmp_path / "seam_two_contexts")
    lib_dir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "seam_two_contexts.c"),
                    "-L" + lib_dir, "-ldespawn_b200", "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([exe, "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    tag, devices, compared, culled = r.stdout.split()
    assert tag == "ok" and int(compared) == 4 * 72 and int(culled) > 0


def test_three_slabs_on_one_gpu_match_one_context(uv_default):
    """Three contexts (a middle rank has two seams), several epochs with edits in between, greedy + indexed mode as well."""
    world = part.Region((-4, 5, 1), (7, 4, 3))
    sp = part.SlabPartition(world, 3)
    with capi.Context(0, 256, 256 << 20) as whole, capi.Context(0, 128, 128 << 20) as c0, capi.Context(0, 128, 128 << 20) as c1, \
            capi.Context(0, 128, 128 << 20) as c2:
        ctxs = [c0, c1, c2]
        for c in [whole] + ctxs:
            c.set_uv_table(uv_default)
        ref = RegionOnDevice(whole, world, capi.GEN_NOISE, SEED)
        devs = [RegionOnDevice(c, sp.region(r), capi.GEN_NOISE, SEED) for r, c in enumerate(ctxs)]
        published, seams = {}, {}
        for r, dev in enumerate(devs):
            for t in sp.seam_transfers(r):
                seam, handle = dev.ctx.seam_create(t.send_face, dev.slots_of(t.positions))
                published[(r, t.send_face)] = handle
                seams[(r, t.send_face)] = (seam, t.peer)
        assert len(seams) == 4
        for (r, face), (seam, peer) in seams.items():
            devs[r].ctx.seam_connect(seam, published[(peer, part.OPPOSITE_FACE[face])])
        rng = np.random.default_rng(8)
        for epoch in range(1, 5):
            if epoch > 1:  # random edits concentrated around the two seams
                k = 400
                e = np.zeros(k, dtype=capi.EDIT_DTYPE)
                seam_x = rng.choice([sp.region(1).origin[0] * 16, sp.region(2).origin[0] * 16], k)
                e["wx"] = seam_x + rng.integers(-2, 2, k)
                e["wy"] = rng.integers(world.origin[1] * 16, (world.origin[1] + world.dims[1]) * 16, k)
                e["wz"] = rng.integers(world.origin[2] * 16, (world.origin[2] + world.dims[2]) * 16, k)
                e["block"] = rng.integers(0, 3, k)
                for c in [whole] + ctxs:
                    c.apply_edits(e)
            for c in ctxs:
                c.seam_exchange_async()
            # (HALO | ORDERED: the count pass of the ordered mode reads the seam inboxes too, behind its own wait)
            for mode in (capi.MESH_HALO, capi.MESH_HALO | capi.MESH_GREEDY | capi.MESH_INDEXED, capi.MESH_HALO | capi.MESH_ORDERED):
                e_ref, _ = whole.mesh_range(ref.first_slot, world.n_chunks, mode)
                for r, dev in enumerate(devs):
                    reg = sp.region(r)
                    dev.ctx.mesh_range_async(dev.first_slot, reg.n_chunks, mode)
                for r, dev in enumerate(devs):
                    reg = sp.region(r)
                    dev.ctx.sync()
                    e, total = dev.ctx.mesh_result(reg.n_chunks)
                    for i, p in enumerate(reg.positions()):
                        w = e_ref[world.index_of(p)]
                        assert int(e[i]["vertex_count"]) == int(w["vertex_count"]) and int(e[i]["index_count"]) == int(w["index_count"]), (epoch, mode, r, tuple(p))
                        assert dev.ctx.download_vertices(e[i]).tobytes() == whole.download_vertices(w).tobytes(), (epoch, mode, r, tuple(p))
            assert c1.seam_epoch == epoch


def test_seam_errors(uv_default):
    with capi.Context(0, 64, 16 << 20) as a, capi.Context(0, 64, 16 << 20) as b:
        for c in (a, b):
            c.set_uv_table(uv_default)
        fa = a.generate_region((0, 6, 0), (2, 2, 2), capi.GEN_NOISE, SEED)
        fb = b.generate_region((2, 6, 0), (2, 2, 2), capi.GEN_NOISE, SEED)
        sa, ha = a.seam_create(capi.FACE_LEFT, [fa + 1, fa + 3, fa + 5, fa + 7])
        sb, hb = b.seam_create(capi.FACE_RIGHT, [fb, fb + 2, fb + 4, fb + 6])
        s3, h3 = b.seam_create(capi.FACE_TOP, [fb + 2, fb + 3])
        with pytest.raises(capi.DspError):
            a.seam_exchange_async()        # not connected yet
        with pytest.raises(capi.DspError):
            a.seam_connect(sa, h3)         # different chunk count
        with pytest.raises(capi.DspError):
            a.seam_connect(sa, ha)         # same face on both sides
        with pytest.raises(capi.DspError):
            a.seam_connect(sa, b"\0" * 128)  # not a handle
        a.seam_connect(sa, hb)
        with pytest.raises(capi.DspError):
            a.seam_connect(sa, hb)         # twice
        with pytest.raises(capi.DspError):
            a.unload_chunks([fa + 1])      # seam chunks stay resident
        # a peer that never pushes: bounded wait, then an error -- not a hang
        a.seam_exchange_async()
        with pytest.raises(capi.DspError) as ei:
            a.mesh_range(fa, 8, capi.MESH_HALO)
        assert ei.value.status == capi.ERR_INTERNAL and "seam" in str(ei.value)


def test_nccl_fallback_entry_points(uv_default):
    """The NCCL transport is exercised across processes by tools/world_bench.py --seam nccl (2 GPUs); here: the library finds
    libnccl.so.2 by itself (no link-time dependency), makes unique ids, and refuses misuse."""
    a, b = capi.comm_unique_id(), capi.comm_unique_id()
    assert len(a) == 128 and a != b
    with capi.Context(0, 64, 16 << 20) as ctx:
        ctx.set_uv_table(uv_default)
        f = ctx.generate_region((0, 6, 0), (2, 2, 2), capi.GEN_NOISE, SEED)
        seam, _ = ctx.seam_create(capi.FACE_LEFT, [f + 1, f + 3, f + 5, f + 7])
        with pytest.raises(capi.DspError) as ei:
            ctx.seam_connect_nccl(seam, 1)     # no communicator yet
        assert ei.value.status == capi.ERR_BAD_ARG
        ctx.comm_init(0, 1, a)                 # a communicator of one rank: enough to check the plumbing
        with pytest.raises(capi.DspError):
            ctx.comm_init(0, 1, b)             # twice
        with pytest.raises(capi.DspError):
            ctx.seam_connect_nccl(seam, 0)     # a seam with oneself
        with pytest.raises(capi.DspError):
            ctx.seam_connect_nccl(99, 1)
