"""Soak test of the Float64 instantiation on the GPU box: N random adversarial sphere soups (tests/test_gpu_parity.py::
_random_soup converted to Float64) against the oracle's Float64 path: equal ray-segment counts, Linf < 1e-9.
Usage: python tools/soak_soups_f64.py [seed] [N]"""
import sys, time, numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import rtw_b200 as rtw
from oracle import binding as O
from test_gpu_parity import _random_soup
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 400
bad = 0
t0 = time.time()
with rtw.Renderer([0]) as r:
    for trial in range(N):
        scene, _ = _random_soup(rtw, rng, trial)
        scene = tuple(a.astype(np.float64) if a.dtype == np.float32 else a for a in scene)
        cam = rtw.default_camera([0.1 * (trial % 7), 0.2, 9 - trial % 5], [0, 0.2, 0], [0, 1, 0], 60, 16 / 9, 0.05 * (trial % 2), 9.0,
                                 elem_type=np.float64)
        img = np.array(r.render(cam, 48, 3, max_depth=10, seed=trial, scene=scene))
        segs = r.last_stats["ray_segments"]
        ref, _, ost = O.render(*scene, cam.as_array(), 48, 3, max_depth=10, seed=trial, f64=True)
        d = np.abs(img - ref).max()
        if segs != ost["ray_segments"] or not (d < 1e-9):
            bad += 1
            print("MISMATCH", trial, len(scene[2]), segs, ost["ray_segments"], d, flush=True)
print("f64 trials", N, "bad", bad, "sec", round(time.time() - t0, 1))
