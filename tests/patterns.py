"""Structured voxel patterns that exercise the merge logic of the greedy mode: runs that change block id, runs that shift by
one voxel from row to row, partial overlaps, cavities, thin plates in every orientation.  Shared by the CPU and GPU tests."""
import numpy as np


def _grid():
    z, y, x = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
    return x, y, z


def structured_chunks():
    x, y, z = _grid()
    out = {}
    out["stripes_x_ids"] = np.where(y < 9, 1 + (x // 3) % 3, 0)                       # id changes along the run axis of TOP/REAR/FRONT
    out["stripes_y_ids"] = np.where(x < 13, 1 + (y // 2) % 2, 0)                      # ... along the run axis of RIGHT/LEFT
    out["stripes_z_ids"] = np.where(y < 5, 1 + (z // 4) % 4, 0)                       # ... along a stack axis
    out["staircase"] = np.where(y <= (x + z) // 2, 1, 0)                               # runs shift by one voxel per row
    out["shear"] = np.where((y < 8) & ((x + z) % 16 < 10), 2, 0)                       # equal-length runs with different starts
    out["hollow_box"] = np.where((np.minimum.reduce([x, y, z, 15 - x, 15 - y, 15 - z]) >= 2) &
                                 (np.minimum.reduce([x, y, z, 15 - x, 15 - y, 15 - z]) < 4), 1, 0)
    out["plates_xy"] = np.where(z % 4 == 1, 1 + (z // 4) % 2, 0)
    out["plates_yz"] = np.where(x % 5 == 2, 3, 0)
    out["plates_xz"] = np.where(y % 3 == 0, 1, 0)
    out["pillars"] = np.where((x % 4 == 1) & (z % 4 == 2), 1 + (y // 8), 0)
    out["l_shapes"] = np.where(((x % 8 < 5) & (z % 8 == 0)) | ((x % 8 == 0) & (z % 8 < 5)), np.where(y < 12, 1, 0), 0)
    out["cavity_grid"] = np.where((x % 4 == 2) & (y % 4 == 2) & (z % 4 == 2), 0, 2)    # solid with isolated holes
    out["two_halves"] = np.where(x < 8, 1, 2)                                          # uniform occupancy, two ids: nothing may merge across x = 8
    out["fallback_ids"] = np.where(y < 6, 1 + (x + y + z) % 5, 0)                      # ids beyond the uv table (fallback uv), dense id changes
    out["diagonal_plane"] = np.where((x + y + z) % 7 == 0, 1, 0)
    out["checker_2x2x2"] = np.where(((x // 2) + (y // 2) + (z // 2)) % 2 == 0, 1, 0)
    return {k: v.astype(np.uint16).reshape(4096) for k, v in out.items()}
