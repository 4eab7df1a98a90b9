"""Times the Float64 path on the headline scene.  Usage: python tools/f64_render.py [spp]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rtw_b200 as R
R.reseed(); scene = R.flatten_scene(R.scene_random_spheres(elem_type=np.float64), np.float64)
cam = R.t_cam1(np.float64)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
with R.Renderer([0]) as r:
    for _ in range(2):
        r.render(cam, 1920, spp, max_depth=50, scene=scene); st = r.last_stats
        print(spp, st["ms_trace"], "Mrays/s", st["ray_segments"] / st["ms_trace"] / 1e3, "T fp64 instr/s", st["sphere_tests"] * 11 / st["ms_trace"] / 1e9, flush=True)
