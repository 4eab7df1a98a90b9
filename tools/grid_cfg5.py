"""RTW_MODE_GRID on BASELINE configs[4] (~100k spheres, 1920x1080) and on the headline scene: time + tier counters.
Usage (GPU box): python tools/grid_cfg5.py [spp5] [spp3]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rtw_b200 as R  # noqa: E402

spp5 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spp3 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
with R.Renderer([0]) as r:
    r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
    for half, spp in ((158, spp5), (11, spp3)):
        R.reseed()
        scene = R.flatten_scene(R.scene_random_spheres(half_extent=half))
        t0 = time.perf_counter()
        r.set_scene(scene)
        for rep in range(2):
            r.render(R.t_cam1(), 1920, spp, max_depth=50, seed=1)
        st = r.last_stats
        if len(sys.argv) > 3:  # compare with the linear sweep at a reduced sample count
            cspp = int(sys.argv[3])
            a = r.render(R.t_cam1(), 1920, cspp, max_depth=50, seed=1).copy()
            sa = dict(r.last_stats)
            r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_FUSED)
            b = r.render(R.t_cam1(), 1920, cspp, max_depth=50, seed=1).copy()
            sb = dict(r.last_stats)
            r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
            print(f"  vs linear sweep at {cspp} spp: identical image {bool((a == b).all())}, segments {sa['ray_segments']} / {sb['ray_segments']}, "
                  f"linear trace {sb['ms_trace']:.0f} ms", flush=True)
        print(f"n={len(scene[2])} spp={spp}: trace {st['ms_trace']:.1f} ms  {st['ray_segments'] / st['ms_trace'] / 1e3:.0f} Mrays/s  "
              f"segments {st['ray_segments']}  loose {st['grid_loose_cells']}  sweep {st['grid_fallback_rays']}  "
              f"cells/segment {st['grid_cells'] / st['ray_segments']:.2f}  tests/segment {st['grid_tests'] / st['ray_segments']:.2f}", flush=True)
