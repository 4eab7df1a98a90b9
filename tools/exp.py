"""Quick A/B of kernel configurations on a slice of the headline workload.
Usage (GPU box): python tools/exp.py [spp] [cfg ...]   cfg = bps:coop:walk (tail is the unified one)"""
import hashlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import rtw_b200 as R  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfgs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]] or [(0, 2, 1), (0, 2, 2), (0, 4, 2), (2, 4, 2)]
R.reseed()
scene = R.flatten_scene(R.scene_random_spheres())
cam = R.t_cam1()
with R.Renderer([0]) as r:
    peak, _ = r.measure_fp32_peak(0)
    r.set_scene(scene)
    for bps, coop, walk in cfgs:
        r.set_option(R.RTW_OPT_TAIL, 2)
        r.set_option(R.RTW_OPT_WALK, walk)
        r.set_option(R.RTW_OPT_BLOCKS_PER_SM, bps)
        r.set_option(R.RTW_OPT_COOP, coop)
        best = None
        for rep in range(3):
            img = r.render(cam, 1920, spp, max_depth=50, seed=1)
            st = dict(r.last_stats)
            if best is None or st["ms_trace"] < best["ms_trace"]:
                best = st
        fp32 = best["sphere_tests"] * 11 / (best["ms_trace"] * 1e-3)
        print(f"bps={bps} coop={coop} walk={walk}: {best['ms_trace']:.2f} ms  {fp32 / 1e12:.2f} T  frac {fp32 / peak:.4f}  "
              f"sha {hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()[:12]}", flush=True)
