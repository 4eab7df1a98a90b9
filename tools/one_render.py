"""One render of scene_random_spheres through the C-ABI (for ncu captures).
Usage: python tools/one_render.py W spp depth [rays] [sweep] [reps] [half_extent] [coop] [mode] [tail]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import rtw_b200 as R  # noqa: E402

W, spp, depth = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rays = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sweep = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
half = int(sys.argv[7]) if len(sys.argv) > 7 else 11
coop = int(sys.argv[8]) if len(sys.argv) > 8 else 0
mode = int(sys.argv[9]) if len(sys.argv) > 9 else 0
tail = int(sys.argv[10]) if len(sys.argv) > 10 else 0
R.reseed()
scene = R.flatten_scene(R.scene_random_spheres(half_extent=half))
with R.Renderer([0]) as r:
    r.set_option(R.RTW_OPT_RAYS_PER_LANE, rays)
    r.set_option(R.RTW_OPT_SWEEP, sweep)
    r.set_option(R.RTW_OPT_COOP, coop)
    r.set_option(R.RTW_OPT_MODE, mode)
    r.set_option(R.RTW_OPT_TAIL, tail)
    r.set_scene(scene)
    for _ in range(reps):
        r.render(R.t_cam1(), W, spp, max_depth=depth, seed=1)
        st = r.last_stats
        print({k: st[k] for k in ("paths", "ray_segments", "ms_trace", "ms_total", "n_spheres")},
              "Mrays/s", st["ray_segments"] / st["ms_trace"] / 1e3,
              "fp32 T/s", st["sphere_tests"] * 11 / st["ms_trace"] / 1e9, flush=True)
