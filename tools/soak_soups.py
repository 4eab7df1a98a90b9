"""Soak test on the GPU box: N random adversarial sphere soups (tests/test_gpu_parity.py::_random_soup, and every 25th
trial a _large_soup of 5k-100k spheres), rendered through the kernel configurations the loaded library contains (linear
default, grid, split wavefront; with RTW_BUILD_VARIANTS=1 also 4 cooperating lanes, per-slot walk, split tail) and
compared with the CPU oracle: equal ray-segment counts, Linf < 1e-6.  Usage: python tools/soak_soups.py [seed] [N]"""
import sys, time, numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import rtw_b200 as rtw
from oracle import binding as O
from test_gpu_parity import _large_soup, _random_soup
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 99)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 150
bad = 0
t0 = time.time()
with rtw.Renderer([0]) as r:
    for trial in range(N):
        large = trial % 25 == 24
        if large:
            scene, cam = _large_soup(rtw, rng, int(rng.choice([5000, 20000, 100000])), trial)
        else:
            scene, cam = _random_soup(rtw, rng, trial)
        ref, _, ost = O.render(*scene, cam.as_array(), 64, 3, max_depth=10, seed=trial)
        configs = [("linear", {}), ("grid", {rtw.RTW_OPT_MODE: 3})]
        if len(scene[2]) <= 1024:
            configs.append(("wavefront", {rtw.RTW_OPT_MODE: 1}))
        if rtw.has_variants() and not large:
            configs += [("coop4own", {rtw.RTW_OPT_COOP: 4, rtw.RTW_OPT_WALK: 2}), ("coop2slots", {rtw.RTW_OPT_WALK: 1}),
                        ("split", {rtw.RTW_OPT_TAIL: 1})]
        for name, opts in configs:
            for k, v in opts.items():
                r.set_option(k, v)
            img = np.array(r.render(cam, 64, 3, max_depth=10, seed=trial, scene=scene))
            segs = r.last_stats["ray_segments"]
            for k in opts:
                r.set_option(k, 0)
            d = np.abs(img.astype(np.float64) - ref).max() if img.size else 0.0
            if segs != ost["ray_segments"] or not (d < 1e-6):
                bad += 1
                print("MISMATCH", trial, name, len(scene[2]), segs, ost["ray_segments"], d, flush=True)
print("trials", N, "bad", bad, "sec", round(time.time() - t0, 1))
