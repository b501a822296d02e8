"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the caller of the hot path: GameScene::update's visible set and
load / unload bookkeeping (/root/reference/src/engine/scenes/scene_game.rs:39-82, 170-223) over the mesher oracle.
Only tests/ may import this module.  PARITY UNPINNED by reference tests (there are none); the known answer is the
default-settings world of SURVEY.md section 3.4 (2,853 chunks, 1,268 meshes, 10,712,064 vertices)."""
from __future__ import annotations

import math

import numpy as np

from . import loader as orc

CHUNK_SIZE = 16


def rust_f32_as_i32(v: float) -> int:
    """`x as i32` for an f32: truncates toward zero, saturates, NaN -> 0."""
    v = float(np.float32(v))
    if math.isnan(v):
        return 0
    return max(-2 ** 31, min(2 ** 31 - 1, int(v)))


def current_chunk(camera_pos) -> tuple:
    """scene_game.rs:41-46: (camera as i32).div_euclid(CHUNK_SIZE) per axis."""
    return tuple(rust_f32_as_i32(c) // CHUNK_SIZE for c in camera_pos)  # Python // is floor == div_euclid for a positive divisor


def visible_chunks(camera_pos, hori_dist: int, vert_dist: int) -> list:
    """scene_game.rs:170-223, statement order preserved: loops over min-1 .. max (exclusive max+1), x outermost."""
    cx, cy, cz = current_chunk(camera_pos)
    is_in_vert = lambda p: abs(p[1] - cy) <= vert_dist
    is_in_hori = lambda p: abs(p[0] - cx) ** 2 + abs(p[2] - cz) ** 2 <= hori_dist ** 2
    max_x, min_x = cx + hori_dist, cx - hori_dist
    max_z, min_z = cz + hori_dist, cz - hori_dist
    max_y, min_y = cy + vert_dist, cy - vert_dist
    out = []
    for x in range(min_x - 1, max_x + 1):
        for y in range(min_y - 1, max_y + 1):
            for z in range(min_z - 1, max_z + 1):
                p = (x, y, z)
                if is_in_vert(p) and is_in_hori(p):
                    out.append(p)
    return out


class GameSceneOracle:
    """world.loaded_chunks / chunk_meshes / last_chunk_pos of GameScene, meshes from the mesher oracle."""

    def __init__(self, uv, hori_dist: int, vert_dist: int, gen_kind: int = orc.GEN_REFERENCE, seed: int = 0):
        self.uv, self.h, self.v, self.kind, self.seed = uv, hori_dist, vert_dist, gen_kind, seed
        self.loaded = {}        # pos -> ids
        self.chunk_meshes = {}  # pos -> vertices or None
        self.last_chunk_pos = None

    def update(self, camera_pos):
        """Returns (changed, newly loaded positions in order, unloaded positions)."""
        cur = current_chunk(camera_pos)
        if cur == self.last_chunk_pos:  # scene_game.rs:48
            return False, [], []
        visible = visible_chunks(camera_pos, self.h, self.v)
        visible_set = set(visible)
        new = []
        for p in visible:  # scene_game.rs:56-64
            if p not in self.loaded:
                ids = orc.generate(self.kind, self.seed, p)
                self.loaded[p] = ids
                m = orc.mesh_chunk(p, ids, self.uv)
                self.chunk_meshes[p] = m if m.shape[0] else None
                new.append(p)
        gone = [p for p in list(self.loaded) if p not in visible_set]  # scene_game.rs:67-73
        for p in gone:
            del self.loaded[p]
            del self.chunk_meshes[p]
        self.last_chunk_pos = cur
        return True, new, gone

    def amount_of_chunk_meshes(self) -> int:  # scene_game.rs:272-274
        return sum(1 for m in self.chunk_meshes.values() if m is not None)
