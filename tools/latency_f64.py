import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import rtw_b200 as R
s64 = R.flatten_scene(R.scene_2_spheres(elem_type=np.float64), np.float64)
cam64 = R.t_default_cam(np.float64)
with R.Renderer([0]) as r:
    for spp in (16, 1):
        best = 1e9
        for _ in range(50):
            t0 = time.perf_counter(); r.render(cam64, 96, spp, scene=s64); best = min(best, time.perf_counter() - t0)
        print("f64 96x54x%d: %.0f us wall, launches %d, trace %.0f us" % (spp, best * 1e6, r.last_stats["kernel_launches"], r.last_stats["ms_trace"] * 1e3))
