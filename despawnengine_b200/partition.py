"""World partitioning across the GPUs of one box and the region-seam (halo) exchange plan.

The reference is single-process; its mesher never looks outside a chunk
(/root/reference/src/content/world/chunks/chunk_mesh.rs:46-51), so in the reference's own mode the world
shards by chunk region with no exchange at all.  The cross-chunk culling extension (DSP_MESH_HALO) needs, at
each seam between two ranks, the one-voxel-thick border layer of the chunks on the other side: 512 bytes per
chunk face (16x16 u16).  This module is pure host logic -- who owns what, who sends which slabs to whom, in which
order -- shared by the GPU path (NCCL over NVLink through torch.distributed) and the CPU tests (gloo).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

FACE_TOP, FACE_BOTTOM, FACE_REAR, FACE_FRONT, FACE_RIGHT, FACE_LEFT = range(6)  # chunk_mesh.rs:46-51 order
OPPOSITE_FACE = (FACE_BOTTOM, FACE_TOP, FACE_FRONT, FACE_REAR, FACE_LEFT, FACE_RIGHT)
SLAB_ELEMS = 256  # 16 x 16 u16


def split_extent(n: int, parts: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split of range(n) into `parts` (start, count) pieces; earlier pieces get the remainder."""
    if parts < 1:
        raise ValueError("parts must be >= 1")
    base, rem = divmod(n, parts)
    out, start = [], 0
    for r in range(parts):
        cnt = base + (1 if r < rem else 0)
        out.append((start, cnt))
        start += cnt
    return out


@dataclass(frozen=True)
class Region:
    """A box of chunks: minimum corner `origin` (chunk coordinates) and `dims` chunks along x, y, z."""
    origin: Tuple[int, int, int]
    dims: Tuple[int, int, int]

    @property
    def n_chunks(self) -> int:
        return self.dims[0] * self.dims[1] * self.dims[2]

    def positions(self) -> np.ndarray:
        """(n,3) int32 in the slot order of dsp_generate_region: x fastest, then y, then z."""
        a, b, c = (np.arange(d, dtype=np.int32) for d in self.dims)
        z, y, x = np.meshgrid(c, b, a, indexing="ij")
        return np.stack([x.ravel() + self.origin[0], y.ravel() + self.origin[1], z.ravel() + self.origin[2]], axis=1).astype(np.int32)

    def index_of(self, pos) -> int:
        a, b, c = (int(pos[i]) - self.origin[i] for i in range(3))
        if not (0 <= a < self.dims[0] and 0 <= b < self.dims[1] and 0 <= c < self.dims[2]):
            raise KeyError(tuple(pos))
        return a + self.dims[0] * (b + self.dims[1] * c)

    def face_positions(self, face: int) -> np.ndarray:
        """Positions of the chunks on the boundary layer of the region at `face`, in canonical order (z outer,
        then y, then x over the two free axes) so that sender and receiver enumerate a seam identically."""
        ox, oy, oz = self.origin
        dx, dy, dz = self.dims
        xs, ys, zs = np.arange(ox, ox + dx), np.arange(oy, oy + dy), np.arange(oz, oz + dz)
        if face == FACE_RIGHT:
            xs = xs[:1]
        elif face == FACE_LEFT:
            xs = xs[-1:]
        elif face == FACE_BOTTOM:
            ys = ys[:1]
        elif face == FACE_TOP:
            ys = ys[-1:]
        elif face == FACE_REAR:
            zs = zs[:1]
        elif face == FACE_FRONT:
            zs = zs[-1:]
        else:
            raise ValueError(face)
        z, y, x = np.meshgrid(zs, ys, xs, indexing="ij")
        return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int32)


@dataclass(frozen=True)
class SeamTransfer:
    """One direction of one seam: this rank packs `send_face` of `send_positions` for `peer`, and registers what it
    receives from `peer` as the neighbour layer across `send_face` of the same chunks (the peer packed the
    opposite face of ITS boundary chunks, enumerated in the same canonical order)."""
    peer: int
    send_face: int
    positions: np.ndarray  # this rank's boundary chunks at the seam, canonical order


class SlabPartition:
    """X-slab partition: the world's X extent is cut into `world_size` contiguous slabs;
    rank r owns slab r, its seams are with r-1 (across its RIGHT/-X face) and r+1 (across its LEFT/+X face)."""

    def __init__(self, world: Region, world_size: int):
        if world.dims[0] < world_size:
            raise ValueError(f"cannot cut {world.dims[0]} chunks of X extent into {world_size} slabs")
        self.world, self.world_size = world, world_size
        self._cuts = split_extent(world.dims[0], world_size)
        self._transfers = {}  # rank -> seam transfers (the position lists are thousands of entries: built once)

    def region(self, rank: int) -> Region:
        start, cnt = self._cuts[rank]
        return Region((self.world.origin[0] + start, self.world.origin[1], self.world.origin[2]), (cnt, self.world.dims[1], self.world.dims[2]))

    def owner_of(self, pos) -> int:
        a = int(pos[0]) - self.world.origin[0]
        for r, (start, cnt) in enumerate(self._cuts):
            if start <= a < start + cnt:
                return r
        raise KeyError(tuple(pos))

    def seam_transfers(self, rank: int) -> List[SeamTransfer]:
        cached = self._transfers.get(rank)
        if cached is not None:
            return cached
        out = self._transfers[rank] = self._seam_transfers(rank)
        return out

    def _seam_transfers(self, rank: int) -> List[SeamTransfer]:
        reg, out = self.region(rank), []
        if rank > 0:
            out.append(SeamTransfer(rank - 1, FACE_RIGHT, reg.face_positions(FACE_RIGHT)))
        if rank < self.world_size - 1:
            out.append(SeamTransfer(rank + 1, FACE_LEFT, reg.face_positions(FACE_LEFT)))
        return out

    def route_edits(self, edits: np.ndarray, halo: bool = False):
        """SURVEY.md section 8e, incremental case: every edit goes to the rank that owns its chunk (by chunk X); edits outside
        the world are dropped.  Returns (per_rank, seam_touched): per_rank[r] = the edits of rank r in their original order
        (so "the last edit of a voxel wins" holds per rank), seam_touched[r] = True when, with cross-chunk culling (`halo`),
        rank r must repack its border slabs for a neighbour because one of its edits lies on a seam-facing border layer.
        `edits` is a structured array with fields wx, wy, wz (world voxel coordinates) and block (capi.EDIT_DTYPE)."""
        e = np.asarray(edits)
        cx = (e["wx"].astype(np.int64) >> 4) - self.world.origin[0]          # div_euclid(16), world.rs:69-74
        cy = (e["wy"].astype(np.int64) >> 4) - self.world.origin[1]
        cz = (e["wz"].astype(np.int64) >> 4) - self.world.origin[2]
        inside = (cx >= 0) & (cx < self.world.dims[0]) & (cy >= 0) & (cy < self.world.dims[1]) & (cz >= 0) & (cz < self.world.dims[2])
        starts = np.array([c[0] for c in self._cuts], dtype=np.int64)
        owner = np.searchsorted(starts, cx, side="right") - 1
        per_rank, seam_touched = [], []
        for r, (start, cnt) in enumerate(self._cuts):
            sel = inside & (owner == r)
            per_rank.append(e[sel])
            touched = False
            if halo and sel.any():
                lx = e["wx"][sel].astype(np.int64) & 15
                c = cx[sel]
                touched = bool(((r > 0) & (c == start) & (lx == 0)).any() or ((r < self.world_size - 1) & (c == start + cnt - 1) & (lx == 15)).any())
            seam_touched.append(touched)
        return per_rank, seam_touched

    def seam_bytes(self, rank: int) -> int:
        """Bytes this rank sends (== receives) per exchange."""
        return sum(len(t.positions) * SLAB_ELEMS * 2 for t in self.seam_transfers(rank))


def exchange_halos(transfers: Sequence[SeamTransfer], pack: Callable[[np.ndarray, int], "object"], alloc_like: Callable[["object"], "object"],
                   unpack: Callable[[np.ndarray, int, "object"], None], dist=None, sync: Optional[Callable[[], None]] = None) -> int:
    """Runs one halo exchange.  `pack(positions, face)` returns a tensor of border slabs (n x 512 uint8 = n x 256 u16) of this rank's
    chunks; `alloc_like(t)` returns an empty tensor of the same shape/device; `unpack(positions, face, tensor)`
    registers received slabs as the neighbour layer across `face`.  `dist` is torch.distributed (NCCL on GPUs, gloo in
    the CPU tests).  All sends and receives of a rank are posted together (batch_isend_irecv), so neighbouring ranks
    cannot deadlock on ordering.  Returns the bytes sent."""
    if not transfers:
        return 0
    import os
    import time
    trace = os.environ.get("DSP_TRACE") is not None
    t0 = time.perf_counter()
    send_bufs = [pack(t.positions, t.send_face) for t in transfers]
    if sync is not None:
        sync()  # packing ran on the context's stream; the collective runs on torch's
    t1 = time.perf_counter()
    recv_bufs = [alloc_like(b) for b in send_bufs]
    ops = []
    for t, sb, rb in zip(transfers, send_bufs, recv_bufs):
        ops.append(dist.P2POp(dist.isend, sb, t.peer))
        ops.append(dist.P2POp(dist.irecv, rb, t.peer))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    if sync is not None:
        sync()
    t2 = time.perf_counter()
    for t, rb in zip(transfers, recv_bufs):
        unpack(t.positions, t.send_face, rb)
    if trace:
        import sys
        print(f"[trace] halo exchange: pack {1e3 * (t1 - t0):.2f} ms, send/recv {1e3 * (t2 - t1):.2f} ms, register {1e3 * (time.perf_counter() - t2):.2f} ms", file=sys.stderr)
    return sum(int(b.numel()) * int(b.element_size()) for b in send_bufs)


def connect_seams(transfers: Sequence[SeamTransfer], rank: int, create: Callable[[int, np.ndarray], Tuple[int, bytes]],
                  connect: Callable[[int, bytes], None], all_gather: Callable[[dict], List[dict]]) -> List[int]:
    """Sets up the in-library seams (dsp_seam_*) of `rank`: one seam per transfer, created with `create(face, positions)
    -> (seam id, handle)`; every rank publishes {(rank, face): handle} through `all_gather` (torch.distributed
    all_gather_object on GPUs and in the gloo tests; any host channel will do -- it runs once, at set-up) and connects
    each of its seams to the handle its peer published for the OPPOSITE face.  Returns the seam ids in transfer order.
    After this, an exchange is `dsp_seam_exchange_async` on every rank: device-to-device stores over NVLink, no host
    in the data path."""
    mine, seams = {}, []
    for t in transfers:
        seam, handle = create(t.send_face, t.positions)
        mine[(rank, t.send_face)] = handle
        seams.append(seam)
    published = {}
    for d in all_gather(mine):
        published.update(d)
    for t, seam in zip(transfers, seams):
        key = (t.peer, OPPOSITE_FACE[t.send_face])
        if key not in published:
            raise KeyError(f"rank {t.peer} published no seam handle for face {OPPOSITE_FACE[t.send_face]} (wanted by rank {rank})")
        connect(seam, published[key])
    return seams
