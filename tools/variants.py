"""Times the kernel variants (rays per lane x sweep) on a slice of the headline workload and the FP32 peaks.
Usage (on the GPU box): python tools/variants.py [spp] ; writes gpurun_out/variants.json"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import rtw_b200 as R  # noqa: E402


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    out = {"spp": spp, "variants": []}
    R.reseed()
    scene = R.flatten_scene(R.scene_random_spheres())
    n = len(scene[2])
    cam = R.t_cam1()
    with R.Renderer([0]) as r:
        for v in (0, 1, 2, 3, 4):
            rate, ms = r.measure_fp32_peak(v)
            out[f"fp32_peak_variant{v}_Tinstr_s"] = rate / 1e12
            out[f"fp32_peak_variant{v}_ms"] = ms
            print(f"fp32 peak variant {v}: {rate / 1e12:.2f} T lane-instr/s ({ms:.3f} ms)", flush=True)
        r.set_scene(scene)
        base = None
        full = len(sys.argv) > 2 and sys.argv[2] == "all"
        legacy = [(1, 1, 0, 1, 1, 0), (1, 2, 0, 1, 1, 0), (1, 3, 0, 1, 1, 0), (2, 3, 0, 1, 1, 0), (1, 3, 0, 4, 1, 0)] if full else []
        for rays, sweep, bps, coop, tail, walk in legacy + [(1, 3, 0, 2, 1, 0), (1, 3, 0, 2, 2, 1), (1, 3, 0, 4, 2, 1),
                                                            (1, 3, 0, 2, 2, 2), (1, 3, 0, 4, 2, 2), (1, 3, 2, 4, 2, 2)]:
            r.set_option(R.RTW_OPT_TAIL, tail)
            r.set_option(R.RTW_OPT_WALK, walk)
            r.set_option(R.RTW_OPT_RAYS_PER_LANE, rays)
            r.set_option(R.RTW_OPT_SWEEP, sweep)
            r.set_option(R.RTW_OPT_BLOCKS_PER_SM, bps)
            r.set_option(R.RTW_OPT_COOP, coop)
            best = None
            for rep in range(3):
                img = r.render(cam, 1920, spp, max_depth=50, seed=1)
                st = dict(r.last_stats)
                if best is None or st["ms_trace"] < best["ms_trace"]:
                    best = st
            if base is None:
                base = np.array(img)
            same = bool(np.array_equal(base, img))
            mrays = best["ray_segments"] / best["ms_trace"] / 1e3
            fp32 = best["sphere_tests"] * 11 / (best["ms_trace"] * 1e-3) / 1e12
            rec = {"rays_per_lane": rays, "sweep": sweep, "blocks_per_sm": bps, "coop": coop, "tail": tail, "walk": walk, "ms_trace": best["ms_trace"],
                   "Mrays_s": mrays, "fp32_Tinstr_s": fp32, "segments_per_path": best["ray_segments"] / best["paths"],
                   "identical_image": same, "n_spheres": n}
            out["variants"].append(rec)
            print(rec, flush=True)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "variants.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
