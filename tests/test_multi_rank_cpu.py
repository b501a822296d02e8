"""N > 1 host logic on CPU: world partitioning and the region-seam (halo) exchange, two ranks over gloo.

Each rank owns one X-slab of a small world, holds its chunks as numpy arrays (from the oracle's generator), packs
the border slabs its neighbour needs, exchanges them with torch.distributed (gloo), and then meshes its seam
chunks in halo mode with the extended oracle using ONLY local chunks + received slabs.  The result must equal the
single-process answer where every neighbour is resident."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from despawnengine_b200 import partition as part
from oracle import loader as orc

SEED = 0xD5E50003
WORLD = part.Region((-2, 0, 1), (5, 3, 3))
UV = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
DIRS = [(0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0)]


def slab_of(ids, face):
    """Border layer of a chunk at `face` in the layout of dsp_pack_border_slabs (+-Y [z][x], +-Z [y][x], +-X [z][y])."""
    v = ids.reshape(16, 16, 16)  # [z][y][x]
    return [v[:, 15, :], v[:, 0, :], v[0, :, :], v[15, :, :], v[:, :, 0], v[:, :, 15]][face].reshape(256)


def halo_mesh(pos, chunks, halo):
    """Halo-mode oracle mesh of `pos` from resident chunks `chunks` and received slabs `halo[(pos, face)]`."""
    slabs = np.zeros((6, 256), dtype=np.uint16)
    present = np.zeros(6, dtype=np.uint8)
    for f, d in enumerate(DIRS):
        q = (pos[0] + d[0], pos[1] + d[1], pos[2] + d[2])
        if q in chunks:
            slabs[f], present[f] = slab_of(chunks[q], part.OPPOSITE_FACE[f]), 1
        elif (pos, f) in halo:
            slabs[f], present[f] = halo[(pos, f)], 1
    return orc.mesh_chunk_halo(pos, chunks[pos], slabs, present, UV)


def test_split_extent_and_regions():
    assert part.split_extent(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert part.split_extent(512, 8) == [(64 * i, 64) for i in range(8)]
    p = part.SlabPartition(part.Region((0, 0, 0), (512, 32, 512)), 8)
    assert p.region(3) == part.Region((192, 0, 0), (64, 32, 512)) and p.owner_of((255, 5, 5)) == 3
    assert p.seam_bytes(0) == 512 * 32 * 512 and p.seam_bytes(4) == 2 * 512 * 32 * 512  # 8 MiB per seam per direction
    r = part.Region((1, 2, 3), (2, 2, 2))
    assert r.positions().tolist()[:3] == [[1, 2, 3], [2, 2, 3], [1, 3, 3]] and r.index_of((2, 3, 4)) == 7
    assert r.face_positions(part.FACE_LEFT).tolist() == [[2, 2, 3], [2, 3, 3], [2, 2, 4], [2, 3, 4]]
    with pytest.raises(ValueError):
        part.SlabPartition(part.Region((0, 0, 0), (2, 1, 1)), 3)


def _rank_main(rank, world_size, port, out_dir):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    try:
        sp = part.SlabPartition(WORLD, world_size)
        reg = sp.region(rank)
        chunks = {tuple(p): orc.generate(orc.GEN_RANDOM, SEED, tuple(p)) for p in reg.positions().tolist()}
        halo = {}

        def pack(positions, face):
            return torch.from_numpy(np.stack([slab_of(chunks[tuple(p)], face) for p in positions.tolist()]).astype(np.uint16).view(np.uint8))

        def unpack(positions, face, t):
            for p, row in zip(positions.tolist(), t.numpy().view(np.uint16)):
                halo[(tuple(p), face)] = row

        sent = part.exchange_halos(sp.seam_transfers(rank), pack, torch.empty_like, unpack, dist=dist)
        assert sent == sp.seam_bytes(rank)
        # single-process truth: every chunk of the world resident
        full = {tuple(p): orc.generate(orc.GEN_RANDOM, SEED, tuple(p)) for p in WORLD.positions().tolist()}
        n_seam = 0
        for p in chunks:
            got = halo_mesh(p, chunks, halo)
            want = halo_mesh(p, full, {})
            assert got.tobytes() == want.tobytes(), (rank, p)
            n_seam += any((p, f) in halo for f in range(6))
        # in-library seams (dsp_seam_*): the set-up routing -- who creates which inbox, whose handle is mapped where -- is host
        # logic shared with the GPU path; here the "handles" just name their inbox
        created, connected = [], {}

        def create(face, positions):
            created.append((face, len(positions)))
            return len(created) - 1, f"inbox(rank={rank},face={face},n={len(positions)})".encode().ljust(128, b"\0")

        def all_gather(obj):
            out = [None] * world_size
            dist.all_gather_object(out, obj)
            return out

        seams = part.connect_seams(sp.seam_transfers(rank), rank, create, lambda seam, handle: connected.__setitem__(seam, handle.rstrip(b"\0").decode()), all_gather)
        assert seams == list(range(len(sp.seam_transfers(rank))))
        per_seam = WORLD.dims[1] * WORLD.dims[2]
        for t, seam in zip(sp.seam_transfers(rank), seams):
            assert connected[seam] == f"inbox(rank={t.peer},face={part.OPPOSITE_FACE[t.send_face]},n={per_seam})"
        total = torch.tensor([len(chunks), n_seam], dtype=torch.int64)
        dist.all_reduce(total)
        assert int(total[0]) == WORLD.n_chunks and int(total[1]) == 2 * (world_size - 1) * WORLD.dims[1] * WORLD.dims[2]
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world_size", [2, 3])
def test_halo_exchange_over_gloo(tmp_path, world_size):
    orc.lib()  # build the oracle before forking
    mp.spawn(_rank_main, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok{r}" for r in range(world_size)]


def test_edits_are_routed_to_the_owning_rank_in_order():
    """SURVEY.md section 8e, incremental case: edits go to the rank owning the chunk (by chunk X), keep their order, and
    an edit on a seam-facing border layer flags the slab the neighbour holds as stale (halo mode)."""
    from despawnengine_b200 import capi
    sp = part.SlabPartition(WORLD, 3)  # X cuts of 5 chunks starting at -2: [-2,-1], [0,1], [2]
    rng = np.random.default_rng(3)
    e = np.zeros(2000, dtype=capi.EDIT_DTYPE)
    e["wx"] = rng.integers(-3 * 16, 4 * 16, 2000)   # some outside the world on both sides
    e["wy"] = rng.integers(-8, 3 * 16 + 8, 2000)
    e["wz"] = rng.integers(16 - 8, 4 * 16 + 8, 2000)
    e["block"] = rng.integers(0, 3, 2000)
    per_rank, seam = sp.route_edits(e, halo=True)
    want = [[], [], []]
    for ed in e:
        c = (int(ed["wx"]) >> 4, int(ed["wy"]) >> 4, int(ed["wz"]) >> 4)
        try:
            WORLD.index_of(c)
        except KeyError:
            continue
        want[sp.owner_of(c)].append(tuple(ed.tolist()))
    for r in range(3):
        assert [tuple(x.tolist()) for x in per_rank[r]] == want[r]
    assert sum(len(p) for p in per_rank) < 2000 and all(len(p) > 0 for p in per_rank)
    # seam flags: rank 0's +X seam layer is world x = -1 (voxel wx in [-16, -1], local x == 15 -> wx == -1)
    assert seam[0] == bool(((e["wx"] == -1) & np.isin(np.arange(2000), np.nonzero([t in set(want[0]) for t in map(lambda x: tuple(x.tolist()), e)])[0])).any())
    only_interior = e[(e["wx"] & 15 != 0) & (e["wx"] & 15 != 15)]
    assert sp.route_edits(only_interior, halo=True)[1] == [False, False, False]
    assert sp.route_edits(e, halo=False)[1] == [False, False, False]
