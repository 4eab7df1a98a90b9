"""Row sharding of one render over ranks / GPUs (reference analogue: `Threads.@threads for i in 1:image_height`,
src/render.jl:23).  Every (pixel, sample) path is independent given the path-keyed Philox stream, so the image is
partitioned by rows with no data-path collective; the only exchange is one gather of the finished row tiles.
Rows are INTERLEAVED (row r -> rank r mod G) so that cheap sky rows and expensive ground rows balance.

Tile layout (what rtw_render_rows_device writes and rtw_assemble_tiles_device reads):
    tile_g[k, j, c] = image row (g + k*G), column j, channel c        k = 0 .. rows_of(g)-1
    gathered[g] = tile_g padded to rows_pad = ceil(H / G) rows
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def rows_pad(height: int, world: int) -> int:
    return (height + world - 1) // world


def rows_for_rank(height: int, rank: int, world: int) -> range:
    """image rows (0-based, top = 0) rendered by `rank`"""
    return range(rank, height, world)


def assemble_tiles_host(tiles: Sequence[np.ndarray], height: int, width: int) -> np.ndarray:
    """NumPy statement of rtw_assemble_tiles_device (used by the CPU tests of the multi-rank logic):
    un-interleave G padded tiles into the (H, W, 3) image."""
    world = len(tiles)
    out = np.zeros((height, width, 3), dtype=tiles[0].dtype)
    for g, t in enumerate(tiles):
        n = len(rows_for_rank(height, g, world))
        out[g::world] = np.asarray(t)[:n]
    return out


def gather_tiles(tile, rank: int, world: int, dst: int = 0):
    """One collective: gather every rank's padded tile on `dst` (torch.distributed; NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return [tile]
    bufs: List = [torch.empty_like(tile) for _ in range(world)] if rank == dst else None
    dist.gather(tile, bufs, dst=dst)
    return bufs
