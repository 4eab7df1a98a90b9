"""Finds the paths on which RTW_MODE_GRID and the linear sweep disagree (per-sample passes through the progressive
API) and prints the oracle's record of those paths.  Usage (GPU box): python tools/grid_debug.py [half_extent] [spp]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rtw_b200 as R  # noqa: E402
from oracle import binding as O  # noqa: E402

half = int(sys.argv[1]) if len(sys.argv) > 1 else 158
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W, depth = 1920, 50
R.reseed()
scene = R.flatten_scene(R.scene_random_spheres(half_extent=half))
cam = R.t_cam1()
found = []
with R.Renderer([0]) as r:
    r.set_scene(scene)
    for s in range(spp):
        acc = {}
        for mode in (R.RTW_MODE_FUSED, R.RTW_MODE_GRID):
            r.set_option(R.RTW_OPT_MODE, mode)
            # sample s alone: a fresh image whose only pass is sample index s is not expressible (first must be 0), so
            # accumulate 0..s and subtract the 0..s-1 state kept from the previous round
            r.accumulate(cam, W, 0, s + 1, spp, max_depth=depth, seed=1)
            acc[mode] = r.accumulator_read().copy()
        diff = np.argwhere((acc[R.RTW_MODE_FUSED] != acc[R.RTW_MODE_GRID]).any(axis=2))
        for i0, j0 in diff:
            if (i0, j0) not in [(a, b) for a, b, _ in found]:
                found.append((int(i0), int(j0), s))
        print(f"samples 0..{s}: {len(diff)} differing pixels", flush=True)
        if len(found) >= 4:
            break
    r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_FUSED)
np.set_printoptions(precision=9, suppress=False, linewidth=200)
for i0, j0, s in found:
    print(f"--- pixel row {i0} col {j0}, first differing at sample {s}")
    rgb, tr = O.path_trace(*scene, cam.as_array(), W, i0, j0, s, max_depth=depth, seed=1)
    for k, row in enumerate(tr):
        kk = int(row[6])
        g = scene[0][kk] if kk >= 0 else None
        print(f"  seg {k}: o={row[0:3]} d={row[3:6]} |d|^2-1={float(np.dot(row[3:6], row[3:6]) - 1):.3e} hit={kk} t={row[7]:.6f} sphere={g}")

# which segment diverges: the same pixel with max_depth = 1, 2, ... in both modes
with R.Renderer([0]) as r:
    r.set_scene(scene)
    for i0, j0, s in found:
        for D in range(1, 14):
            px = {}
            for mode in (R.RTW_MODE_FUSED, R.RTW_MODE_GRID):
                r.set_option(R.RTW_OPT_MODE, mode)
                a = r.accumulate(cam, W, 0, s + 1, spp, max_depth=D, seed=1)
                acc_s = r.accumulator_read()[i0, j0].copy()
                if s > 0:
                    r.accumulate(cam, W, 0, s, spp, max_depth=D, seed=1)
                    acc_s = acc_s - r.accumulator_read()[i0, j0]
                px[mode] = acc_s
            same = bool((px[R.RTW_MODE_FUSED] == px[R.RTW_MODE_GRID]).all())
            print(f"pixel ({i0},{j0}) sample {s} max_depth {D}: linear {px[R.RTW_MODE_FUSED][:3]} grid {px[R.RTW_MODE_GRID][:3]} {'same' if same else 'DIFFERENT'}", flush=True)
