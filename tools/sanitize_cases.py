"""Small renders through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck).
Usage: compute-sanitizer --tool racecheck python tools/sanitize_cases.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import rtw_b200 as R  # noqa: E402

R.reseed()
small = R.flatten_scene(R.scene_random_spheres())            # 484 spheres: single resident tile
R.reseed()
big = R.flatten_scene(R.scene_random_spheres(half_extent=17))  # > 1024 spheres: streamed TMA tiles
cam = R.t_cam1()
with R.Renderer([0]) as r:
    r.set_option(R.RTW_OPT_SMALL_RENDER, 0)  # the persistent kernels first
    for scene, W, spp in ((small, 64, 2), (big, 32, 1)):
        for coop, tail, walk, mode in ((2, 2, 1, 0), (2, 2, 2, 0), (4, 2, 1, 0), (4, 2, 2, 0), (2, 1, 0, 0), (2, 0, 0, 2), (2, 0, 0, 1), (2, 0, 0, 3)):
            if mode in (1, 2) and len(scene[2]) > 1024:
                continue
            if not R.has_variants() and (mode == 2 or (mode == 0 and (coop != 2 or tail != 2 or (walk == 1 and len(scene[2]) <= 1024)))):
                continue  # needs RTW_BUILD_VARIANTS=1
            r.set_option(R.RTW_OPT_COOP, coop)
            r.set_option(R.RTW_OPT_TAIL, tail)
            r.set_option(R.RTW_OPT_WALK, walk)
            r.set_option(R.RTW_OPT_MODE, mode)
            r.render(cam, W, spp, max_depth=8, scene=scene)
            print("ok", len(scene[2]), coop, tail, walk, mode, r.last_stats["ray_segments"], flush=True)
    for o in (R.RTW_OPT_COOP, R.RTW_OPT_TAIL, R.RTW_OPT_WALK, R.RTW_OPT_MODE):
        r.set_option(o, 0)
    r.set_scene(small)
    r.accumulate(cam, 48, 0, 1, 2, max_depth=8)
    r.accumulate(cam, 48, 1, 1, 2, max_depth=8)
    r.resolve_rgb8()
    s64 = R.flatten_scene(R.scene_4_spheres(elem_type=np.float64), np.float64)
    r.render(R.t_default_cam(np.float64), 48, 2, max_depth=8, scene=s64)
    R.reseed()
    r64 = R.flatten_scene(R.scene_random_spheres(elem_type=np.float64), np.float64)  # 484 spheres: cooperative f64 sweep, ragged
    r.render(R.t_cam1(np.float64), 48, 2, max_depth=8, scene=r64)
    print("ok progressive + f64", flush=True)
    # the latency path (one launch), the device scene generator, a large list through the grid (loose lists + culled sweep)
    r.set_option(R.RTW_OPT_SMALL_RENDER, 1)
    r.render(cam, 64, 3, max_depth=8, scene=small)
    assert r.last_stats["kernel_launches"] == 1
    r.render(cam, 64, 3, max_depth=8, scene=small)
    R.reseed()
    g = r.generate_random_spheres(40)
    print("ok small render + scene generator", len(g[2]), flush=True)
    r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
    r.render(cam, 96, 2, max_depth=12)
    print("ok grid on", r.n_spheres, "spheres:", r.last_stats["grid_loose_cells"], "loose cells,", r.last_stats["grid_fallback_rays"], "sweeps", flush=True)
