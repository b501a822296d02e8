"""Synthetic inputs of the BASELINE.json configs, shared by bench.py, tools/ and tests/ so that the
sizes the bench reports are the sizes the parity tests check.  Host-side numpy only; nothing here touches the device."""
from __future__ import annotations

import numpy as np

SEED_NOISE = 0xD5E50001   # config 2 terrain
SEED_RANDOM = 0xD5E50003  # config 3b Bernoulli(0.5)
SEED_EDITS = 0xD5E50004   # config 4 edit stream
UV_DEFAULT = np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1, _M2 = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)


def splitmix64(state: int, n: int):
    """The next n outputs of SplitMix64 from `state`; returns (new state, outputs u64[n]).  Vectorised: output k is the
    mix of state + (k + 1) * gamma."""
    with np.errstate(over="ignore"):
        s = np.uint64(state) + np.arange(1, n + 1, dtype=np.uint64) * _GAMMA
        z = (s ^ (s >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return int(s[-1]), z


def edit_frame(state: int, n_edits: int, dims_chunks, origin_chunks=(0, 0, 0)):
    """Config 4: `n_edits` edits uniform over the world box, 50 % destroy (block 0) / 50 % place (block 1), from
    SplitMix64.  Returns (new state, edits as an array of dsp_edit {wx, wy, wz, block})."""
    from .capi import EDIT_DTYPE
    state, r = splitmix64(state, 4 * n_edits)
    r = r.reshape(n_edits, 4)
    e = np.zeros(n_edits, dtype=EDIT_DTYPE)
    for a, name in enumerate(("wx", "wy", "wz")):
        e[name] = (r[:, a] % np.uint64(dims_chunks[a] * 16)).astype(np.int64) + origin_chunks[a] * 16
    e["block"] = (r[:, 3] & np.uint64(1)).astype(np.uint32)
    return state, e


def box_positions(origin, dims) -> np.ndarray:
    """(n, 3) i32 chunk positions of a box in the slot order of dsp_generate_region: x fastest, then y, then z."""
    a, b, c = (np.arange(d, dtype=np.int32) for d in dims)
    z, y, x = np.meshgrid(c, b, a, indexing="ij")
    return np.stack([x.ravel() + origin[0], y.ravel() + origin[1], z.ravel() + origin[2]], axis=1).astype(np.int32)


def replay_edits_on_chunk(ids: np.ndarray, chunk_pos, edits: np.ndarray) -> np.ndarray:
    """Applies, in order, the edits that fall into chunk `chunk_pos` to its voxels (world.rs:69-74 addressing: chunk =
    div_euclid(w, 16), local = rem_euclid(w, 16); Chunk::set_block, chunk.rs:49-68).  Checker-side helper."""
    sel = ((edits["wx"] >> 4) == chunk_pos[0]) & ((edits["wy"] >> 4) == chunk_pos[1]) & ((edits["wz"] >> 4) == chunk_pos[2])
    out = ids.copy()
    for e in edits[sel]:
        out[(int(e["wx"]) & 15) + 16 * (int(e["wy"]) & 15) + 256 * (int(e["wz"]) & 15)] = e["block"]
    return out
