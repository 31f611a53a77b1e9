"""Screen-tile partition across the GPUs of one box, one process per GPU.

Tiles are independent (a pixel depends only on the ordered list of splats covering it), so the path
shards by interleaved bands of tile rows: rank g owns tile rows ty with (ty // group) % G == g.  Every rank
holds all splat attributes, culls to its rows, sorts / bins / blends only those, and writes zeros to
rows it does not own.  The only exchange step is the framebuffer combine below: one NCCL reduction per
frame onto the display rank (exact, because every pixel is non-zero on exactly one rank: x + 0 == x).
The reference has no multi-GPU path at all.
"""
from __future__ import annotations

import numpy as np

TILE = 16


def owned_tile_rows(height: int, rank: int, world: int, group: int = 1) -> list[int]:
    tiles_y = (height + TILE - 1) // TILE
    return [ty for ty in range(tiles_y) if (ty // max(group, 1)) % world == rank]


def owned_scanlines(height: int, rank: int, world: int, group: int = 1) -> np.ndarray:
    """bool[height]: scanlines (bottom-up, GL order) this rank renders."""
    y = np.arange(height)
    return (((y // TILE) // max(group, 1)) % world) == rank


def default_row_group(height: int, world: int, bands_per_rank: int = 8) -> int:
    """Band height in tile rows: wide enough that few splats straddle two ranks (each straddler is projected,
    sorted and binned twice), narrow enough that every rank still gets >= bands_per_rank bands for load balance."""
    tiles_y = (height + TILE - 1) // TILE
    return max(1, tiles_y // (max(world, 1) * bands_per_rank))


def combine_on_root(fb, rank: int, world: int, root: int = 0, group=None):
    """fb: [H,W,4] float32 torch tensor holding this rank's rows (zeros elsewhere).  After the call the
    root holds the whole frame.  NCCL on GPUs (NVLink 5 / NVSwitch), gloo in the CPU tests."""
    if world == 1:
        return fb
    import torch.distributed as dist
    dist.reduce(fb, dst=root, op=dist.ReduceOp.SUM, group=group)
    return fb
