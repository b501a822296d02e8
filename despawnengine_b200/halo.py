"""GPU side of the region-seam exchange: border slabs are packed by a CUDA kernel into a torch tensor (device memory),
moved between ranks by torch.distributed (NCCL over NVLink/NVSwitch) and registered with the receiving context.
torch is plumbing here (device memory + the collective); the pack and the culling are libdespawn_b200's kernels."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import capi
from .partition import Region, SlabPartition, connect_seams, exchange_halos


class RegionOnDevice:
    """One rank's slab of the world, generated on its GPU into consecutive slots."""

    def __init__(self, ctx: capi.Context, region: Region, gen_kind: int, seed: int):
        self.ctx, self.region = ctx, region
        self.first_slot = ctx.generate_region(region.origin, region.dims, gen_kind, seed)
        self.seam_slots = {}  # (face, n) -> slots of this region's boundary chunks at that face (X-slab partition: one seam per face)

    def slots_of(self, positions: np.ndarray) -> np.ndarray:
        p = np.asarray(positions, dtype=np.int64).reshape(-1, 3) - np.array(self.region.origin, dtype=np.int64)
        d = self.region.dims
        if len(p) and (p.min() < 0 or (p >= np.array(d)).any()):
            raise KeyError("position outside the region")
        return (self.first_slot + p[:, 0] + d[0] * (p[:, 1] + d[1] * p[:, 2])).astype(np.uint32)


def exchange_region_halos(dev: RegionOnDevice, partition: SlabPartition, rank: int, dist) -> int:
    """One halo exchange for `rank`; afterwards DSP_MESH_HALO on this context culls seam faces against the neighbouring
    ranks' border layers.  Returns bytes sent."""
    import torch
    ctx = dev.ctx
    device = torch.device("cuda", ctx.device)

    def slots(positions, face):  # a seam's slot list does not change between exchanges
        key = (face, len(positions))
        if key not in dev.seam_slots:
            dev.seam_slots[key] = dev.slots_of(positions)
        return dev.seam_slots[key]

    def pack(positions, face):
        t = torch.empty((len(positions), 512), dtype=torch.uint8, device=device)  # NCCL has no 16-bit integer type: ship bytes
        ctx.pack_border_slabs(slots(positions, face), face, t.data_ptr())
        return t

    def unpack(positions, face, t):
        ctx.set_halo_slabs(slots(positions, face), face, t.data_ptr())
        ctx.sync()  # the copy out of `t` is enqueued on the context's stream; `t` may be freed after this returns

    def sync():
        ctx.sync()
        torch.cuda.synchronize(device)

    return exchange_halos(partition.seam_transfers(rank), pack, torch.empty_like, unpack, dist=dist, sync=sync)


def connect_region_seams(dev: RegionOnDevice, partition: SlabPartition, rank: int, dist=None):
    """In-library seams (dsp_seam_*, the product path): creates this rank's inboxes, swaps the 128-byte handles with the
    neighbouring ranks once (all_gather_object; `dist` None = single process, nothing to connect) and maps the peers' inboxes.
    From then on `dev.ctx.seam_exchange_async()` followed by a DSP_MESH_HALO call is all an epoch takes: the border layers
    travel by peer stores over NVLink and the wait for them sits inside the mesh kernel, behind the interior chunks."""
    transfers = partition.seam_transfers(rank)
    if not transfers:
        return []

    def all_gather(obj):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out

    return connect_seams(transfers, rank, lambda face, positions: dev.ctx.seam_create(face, dev.slots_of(positions)),
                         dev.ctx.seam_connect, all_gather)


def connect_region_seams_nccl(dev: RegionOnDevice, partition: SlabPartition, rank: int, dist):
    """The same seams over the library's NCCL fallback transport (dsp_comm_init + dsp_seam_connect_nccl): for GPUs whose memory
    cannot be mapped into each other.  The unique id is made on rank 0 and broadcast once; ncclSend / ncclRecv are issued from
    inside the library on the context's stream."""
    from . import capi
    box = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    dev.ctx.comm_init(rank, dist.get_world_size(), box[0])
    seams = []
    for t in partition.seam_transfers(rank):
        seam, _ = dev.ctx.seam_create(t.send_face, dev.slots_of(t.positions))
        dev.ctx.seam_connect_nccl(seam, t.peer)
        seams.append(seam)
    return seams
