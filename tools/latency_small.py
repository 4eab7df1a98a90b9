"""Latency of small renders through the host-buffer API (the reference's own smoke/benchmark sizes).
Usage: python tools/latency_small.py"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rtw_b200 as R  # noqa: E402

cases = [("scene_2_spheres 96x54x16 (test/runtests.jl:194; reference: 951 us, 16 threads, Float64)", R.scene_2_spheres(), R.t_default_cam(), 96, 16),
         ("scene_2_spheres 96x54x1 (reference: 101 us)", R.scene_2_spheres(), R.t_default_cam(), 96, 1)]
R.reseed()
rnd = R.scene_random_spheres()
cases += [("random_spheres 200x112x32 (src/proto/proto.jl:195-200; reference: 296.8 ms)", rnd, R.t_cam1(), 200, 32),
          ("random_spheres 96x54x1 (reference: 2.04 ms)", rnd, R.t_cam1(), 96, 1)]
with R.Renderer([0]) as r:
    for name, scene, cam, W, spp in cases:
        flat = R.flatten_scene(scene)
        best_wall, best = 1e9, None
        for _ in range(30):
            t0 = time.perf_counter()
            r.render(cam, W, spp, scene=flat)  # rtw_render_scene: what the Julia render(scene, cam, W, spp) binds to
            wall = time.perf_counter() - t0
            if wall < best_wall:
                best_wall, best = wall, dict(r.last_stats)
        print(f"{name}: wall {best_wall * 1e6:.0f} us (device total {best['ms_total'] * 1e3:.0f} us, trace {best['ms_trace'] * 1e3:.0f} us)", flush=True)
