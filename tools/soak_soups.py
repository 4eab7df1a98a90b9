"""Soak test on the GPU box: N random adversarial sphere soups (tests/test_gpu_parity.py::_random_soup), rendered through
five kernel configurations (linear default, grid, 4 cooperating lanes + own-ray walk, 2 lanes + own-ray walk, split tail)
and compared with the CPU oracle: equal ray-segment counts, Linf < 1e-6.  Usage: python tools/soak_soups.py [seed] [N]"""
import sys, time, numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import rtw_b200 as rtw
from oracle import binding as O
from test_gpu_parity import _random_soup
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 99)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 150
bad = 0
t0 = time.time()
with rtw.Renderer([0]) as r:
    for trial in range(N):
        scene, cam = _random_soup(rtw, rng, trial)
        ref, _, ost = O.render(*scene, cam.as_array(), 64, 3, max_depth=10, seed=trial)
        for name, opts in (("linear", {}), ("grid", {rtw.RTW_OPT_MODE: 3}), ("coop4own", {rtw.RTW_OPT_COOP: 4, rtw.RTW_OPT_WALK: 2}),
                           ("coop2own", {rtw.RTW_OPT_WALK: 2}), ("split", {rtw.RTW_OPT_TAIL: 1})):
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
