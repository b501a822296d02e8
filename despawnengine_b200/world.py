"""Host-side mirror of the reference's chunk / world / mesh interface over the C ABI.

Same names and argument meaning as the Rust items this path replaces (paths under
/root/reference/src), so the parity tests read like tests of the reference:

    Chunk             content/world/chunks/chunk.rs:15-82     (position, blocks, palette, palette_map)
    World             content/world/world.rs:10-88            (get_chunk, load_chunk, unload_chunk)
    build_chunk_mesh  content/world/chunks/chunk_mesh.rs:13   (chunk, block_uvs) -> Option<vertex buffer>
    AtlasUV           engine/rendering/texture_atlas.rs:15-19

The only host work here is what the reference's host also does *around* the hot loop: resolving
block-id strings to numbers.  Voxel generation, culling and vertex emission all happen in
libdespawn_b200.so on the GPU; nothing in this module computes a face.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import capi

CHUNK_SIZE = 16            # chunk.rs:7
MAX_CHUNK_INDEX = 4095     # chunk.rs:9
AIR_BLOCK_ID = "base:air"  # chunk.rs:12


@dataclass(frozen=True)
class AtlasUV:  # texture_atlas.rs:15-19
    uv_min: Tuple[float, float]
    uv_max: Tuple[float, float]


class BlockRegistry:
    """Block-id string -> global numeric id (0 = air).  Stands in for Registry<Block>
    (utils/registry.rs:4-24); ids are assigned in registration order."""

    def __init__(self, ids: Iterable[str] = ("template:dirt", "template:engine")):
        self._ids: Dict[str, int] = {AIR_BLOCK_ID: 0}
        self._names: List[str] = [AIR_BLOCK_ID]
        for b in ids:
            self.register(b)

    def register(self, block_id: str) -> int:
        if block_id not in self._ids:
            if len(self._names) >= 65535:
                raise ValueError("block registry full (u16 ids)")
            self._ids[block_id] = len(self._names)
            self._names.append(block_id)
        return self._ids[block_id]

    def id_of(self, block_id: str) -> int:
        return self._ids[block_id] if block_id in self._ids else self.register(block_id)

    def name_of(self, gid: int) -> str:
        return self._names[gid]

    def __len__(self) -> int:
        return len(self._names)

    def uv_table(self, block_uvs: Dict[str, AtlasUV]) -> np.ndarray:
        """(n,4) rows for dsp_set_uv_table.  A block absent from block_uvs gets the reference's
        fallback (0,0)-(1,1) (chunk_mesh.rs:53-56)."""
        t = np.zeros((len(self._names), 4), dtype=np.float32)
        t[:, 2:] = 1.0
        for name, uv in block_uvs.items():
            if name in self._ids:
                t[self._ids[name]] = (*uv.uv_min, *uv.uv_max)
        return t


@dataclass
class Chunk:
    """Host-resident chunk data with the reference's per-chunk string palette (chunk.rs:15-68).
    Only needed when the caller owns voxel data on the host; generated chunks never leave the GPU."""
    position: Tuple[int, int, int]
    blocks: np.ndarray = field(default_factory=lambda: np.zeros(4096, dtype=np.uint16))
    palette: List[str] = field(default_factory=lambda: [AIR_BLOCK_ID])
    palette_map: Dict[str, int] = field(default_factory=lambda: {AIR_BLOCK_ID: 0})

    @staticmethod
    def index(x: int, y: int, z: int) -> int:  # chunk.rs:44-46
        return x + y * CHUNK_SIZE + z * CHUNK_SIZE * CHUNK_SIZE

    def is_air_at_idx(self, idx: int) -> bool:  # chunk.rs:39-41
        return self.blocks[idx] == 0

    def set_block(self, x: int, y: int, z: int, block_id: str) -> None:  # chunk.rs:49-68
        idx = self.index(x, y, z)
        if block_id == AIR_BLOCK_ID:
            self.blocks[idx] = 0
            return
        i = self.palette_map.get(block_id)
        if i is None:
            i = len(self.palette)
            self.palette.append(block_id)
            self.palette_map[block_id] = i
        self.blocks[idx] = i


@dataclass
class ChunkMesh:
    """What build_chunk_mesh returns in place of Subbuffer<[BlockVertex]>: a window of the device arena."""
    world: "World"
    offset_bytes: int
    vertex_count: int
    index_count: int = 0  # > 0 with MESH_INDEXED: draw_indexed(index_count, ...) instead of draw(vertex_count, ...)

    def __len__(self) -> int:  # mesh.len() as used by draw(), scene_game.rs:118-122
        return self.vertex_count

    @property
    def device_ptr(self) -> int:
        return self.world.ctx.arena_device_ptr + self.offset_bytes

    def to_host(self) -> np.ndarray:
        return self.world.ctx.download_arena(self.offset_bytes, self.vertex_count * 20).view(capi.VERTEX_DTYPE)

    @property
    def index_offset_bytes(self) -> int:
        return self.offset_bytes + ((self.vertex_count * 20 + 127) & ~127)

    def indices_to_host(self) -> np.ndarray:
        return self.world.ctx.download_arena(self.index_offset_bytes, self.index_count * 4).view(np.uint32)


class World:
    """world.rs:10-88 over a device-resident voxel store."""

    def __init__(self, device: int = 0, max_chunks: int = 4096, arena_bytes: int = 1 << 28, registry: Optional[BlockRegistry] = None,
                 gen_kind: int = capi.GEN_REFERENCE, seed: int = 0):
        self.ctx = capi.Context(device, max_chunks, arena_bytes)
        self.registry = registry or BlockRegistry()
        self.chunks: Dict[Tuple[int, int, int], int] = {}         # pos -> slot   (world.rs:11)
        self.loaded_chunks: Dict[Tuple[int, int, int], int] = {}  # world.rs:12
        self.gen_kind, self.seed = gen_kind, seed
        self._arena_top = 0
        self._uv_key = None

    def close(self):
        self.ctx.close()

    # world.rs:28-47 -- cached, else generated (on the device) by the rule for this chunk position
    def get_chunk(self, pos) -> int:
        return int(self.get_chunks([pos])[0])

    def get_chunks(self, positions: Sequence[Sequence[int]]) -> np.ndarray:
        positions = [tuple(int(c) for c in p) for p in positions]
        missing = [p for p in dict.fromkeys(positions) if p not in self.chunks]
        if missing:
            slots = self.ctx.generate_chunks(np.array(missing, dtype=np.int32), self.gen_kind, self.seed)
            for p, s in zip(missing, slots):
                self.chunks[p] = int(s)
        return np.array([self.chunks[p] for p in positions], dtype=np.uint32)

    def put_chunk(self, chunk: Chunk) -> int:
        """Uploads host chunk data (per-chunk palette resolved to global ids on the device)."""
        pal = np.array([self.registry.id_of(b) for b in chunk.palette], dtype=np.uint16)
        off = np.array([0, len(pal)], dtype=np.uint32)
        slot = int(self.ctx.load_chunks([chunk.position], chunk.blocks, pal, off)[0])
        self.chunks[tuple(chunk.position)] = slot
        return slot

    def get_block_world(self, wx: int, wy: int, wz: int) -> Optional[str]:
        """world.rs:50-65: the block id string at world voxel coordinates; None when the chunk is not in `chunks`."""
        gid = self.ctx.get_block_world(int(wx), int(wy), int(wz))
        return None if gid is None else self.registry.name_of(gid)

    def load_chunk(self, chunk_pos) -> None:  # world.rs:80-84
        p = tuple(int(c) for c in chunk_pos)
        self.loaded_chunks[p] = self.get_chunk(p)

    def unload_chunk(self, chunk_pos) -> None:  # world.rs:85-87
        self.loaded_chunks.pop(tuple(int(c) for c in chunk_pos), None)

    def set_block_uvs(self, block_uvs: Dict[str, AtlasUV]) -> None:
        key = (tuple(sorted(block_uvs.items())), len(self.registry))
        if key != self._uv_key:
            self.ctx.set_uv_table(self.registry.uv_table(block_uvs))
            self._uv_key = key


def build_chunk_meshes(world: World, positions: Sequence[Sequence[int]], block_uvs: Dict[str, AtlasUV],
                       mode: int = capi.MESH_PARITY) -> List[Optional[ChunkMesh]]:
    """Batched build_chunk_mesh: GameScene::update's load loop (scene_game.rs:56-64) as one call."""
    world.set_block_uvs(block_uvs)
    slots = world.get_chunks(positions)
    entries, total = world.ctx.mesh_chunks(slots, mode, world._arena_top)
    world._arena_top += total
    return [ChunkMesh(world, int(e["offset_bytes"]), int(e["vertex_count"]), int(e["index_count"])) if e["vertex_count"] else None for e in entries]


def build_chunk_mesh(world: World, chunk_pos, block_uvs: Dict[str, AtlasUV]) -> Optional[ChunkMesh]:
    """chunk_mesh.rs:13-17 -- None when the chunk is all air / produces no faces."""
    return build_chunk_meshes(world, [chunk_pos], block_uvs)[0]


def build_all_host_chunks(world: World, chunks: Sequence[Chunk], block_uvs: Dict[str, AtlasUV], mode: int = capi.MESH_PARITY,
                          blocks_out: Optional[np.ndarray] = None) -> List[Optional[ChunkMesh]]:
    """GameScene::build_all_loaded_chunks (scene_game.rs:231-244) for chunks the caller holds on the HOST: every `Chunk` (blocks +
    string palette, chunk.rs:15-20) is loaded and meshed by one dsp_mesh_host_chunks_ex call.  The host does what the reference's
    host does around the loop -- palette strings to numbers -- and hands over the palette lengths, so that a chunk whose palette
    is just "base:air" is never fetched (chunk_mesh.rs:18-20).  `blocks_out` (n x 4096 u16, e.g. a pinned torch buffer's numpy
    view) receives the global-id voxels; with pinned memory the mesh kernel pulls the tiles itself."""
    world.set_block_uvs(block_uvs)
    n = len(chunks)
    gids = blocks_out if blocks_out is not None else np.empty((n, 4096), dtype=np.uint16)
    pos = np.empty((n, 3), dtype=np.int32)
    palette_lens = np.empty(n, dtype=np.uint32)
    for i, c in enumerate(chunks):
        pos[i] = c.position
        palette_lens[i] = len(c.palette)
        gids[i] = np.array([world.registry.id_of(b) for b in c.palette], dtype=np.uint16)[c.blocks]  # per-chunk palette index -> global id (chunk.rs:58-65)
    entries = np.zeros(n, dtype=capi.MESH_ENTRY_DTYPE)
    slots = np.zeros(n, dtype=np.uint32)
    total = world.ctx.mesh_host_chunks_ex_ptr(n, pos, gids.ctypes.data, palette_lens, entries, mode, world._arena_top, slots_out=slots)
    for c, s in zip(chunks, slots):
        world.chunks[tuple(int(k) for k in c.position)] = int(s)
    world._arena_top += total
    return [ChunkMesh(world, int(e["offset_bytes"]), int(e["vertex_count"]), int(e["index_count"])) if e["vertex_count"] else None for e in entries]
