"""world_size-2 test of the multi-rank host logic on CPU (gloo): interleaved row sharding, one gather of the
padded tiles to rank 0, assembly -- bit-identical to the single-rank image.  The per-rank "renderer" here is the
oracle's row-subset render (test infrastructure); on GPUs the same sharding code drives rtw_render_rows_device."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_path, width, spp):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist

    import rtw_b200 as R
    from oracle import binding as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        R.reseed()
        scene = R.flatten_scene(R.scene_random_spheres())
        cam = R.t_cam1()
        H = R.image_height(width)
        pad = R.sharding.rows_pad(H, world)
        rows = R.sharding.rows_for_rank(H, rank, world)
        img, _, st = O.render(*scene, cam.as_array(), width, spp, max_depth=8, seed=1, n_threads=2, row_start=rank,
                              row_stride=world)
        tile = torch.zeros((pad, width, 3), dtype=torch.float32)
        tile[:len(rows)] = torch.from_numpy(np.ascontiguousarray(img[rank::world]))
        tiles = R.sharding.gather_tiles(tile, rank, world, dst=0)
        segs = torch.tensor([float(st["ray_segments"])], dtype=torch.float64)
        dist.all_reduce(segs, op=dist.ReduceOp.SUM)
        if rank == 0:
            full = R.sharding.assemble_tiles_host([t.numpy() for t in tiles], H, width)
            np.savez(out_path, image=full, segments=segs.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_render_equals_single_rank(tmp_path, world, oracle, rtw):
    import torch.multiprocessing as mp

    width, spp = 64, 2
    out_path = str(tmp_path / "sharded.npz")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, out_path, width, spp), nprocs=world, join=True)
    got = np.load(out_path)
    rtw.reseed()
    scene = rtw.flatten_scene(rtw.scene_random_spheres())
    ref, _, st = oracle.render(*scene, rtw.t_cam1().as_array(), width, spp, max_depth=8, seed=1)
    assert np.array_equal(got["image"], ref)
    assert int(got["segments"][0]) == st["ray_segments"]


def test_sharding_helpers(rtw):
    S = rtw.sharding
    for H in (0, 1, 9, 54, 1080):
        for G in (1, 2, 3, 4, 8):
            rows = [list(S.rows_for_rank(H, g, G)) for g in range(G)]
            assert sorted(sum(rows, [])) == list(range(H))
            assert max(len(r) for r in rows) <= S.rows_pad(H, G)
    img = np.arange(7 * 5 * 3, dtype=np.float32).reshape(7, 5, 3)
    tiles = []
    for g in range(3):
        t = np.zeros((S.rows_pad(7, 3), 5, 3), np.float32)
        t[:len(S.rows_for_rank(7, g, 3))] = img[g::3]
        tiles.append(t)
    assert np.array_equal(S.assemble_tiles_host(tiles, 7, 5), img)
