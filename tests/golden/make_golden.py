"""Generates the committed golden fixtures from the CPU oracle (the reference itself is Julia and cannot run
in this image -- SURVEY.md section 8c).  Run from the repo root:  python tests/golden/make_golden.py
The fixtures pin (a) the oracle against accidental change and (b) the CUDA path on boxes where only the
fixtures travel."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import rtw_b200 as R  # noqa: E402
from oracle import binding as O  # noqa: E402

HERE = Path(__file__).resolve().parent


def main():
    # cfg1: scene_2_spheres, 96x54, 16 spp, 4 bounces, Float32, 1 CPU thread, fixed seed
    g, m, k = R.flatten_scene(R.scene_2_spheres())
    cam = R.t_default_cam()
    img, lin, st = O.render(g, m, k, cam.as_array(), 96, 16, max_depth=4, seed=1, n_threads=1, want_linear=True)
    np.savez_compressed(HERE / "cfg1_scene_2_spheres_96x54_16spp_d4_seed1.npz", image=img, geom=g, mat=m, kind=k,
                        camera=cam.as_array(), ray_segments=np.uint64(st["ray_segments"]))
    # per-path vectors on the random scene: (i0, j0, s0) -> linear rgb, segment count
    R.reseed()
    g, m, k = R.flatten_scene(R.scene_random_spheres())
    cam = R.t_cam1()
    rng = np.random.default_rng(2024)
    rows = []
    for _ in range(256):
        i0, j0, s0 = int(rng.integers(0, 225)), int(rng.integers(0, 400)), int(rng.integers(0, 64))
        rgb, nseg = O.path(g, m, k, cam.as_array(), 400, i0, j0, s0, max_depth=16, seed=1)
        rows.append([i0, j0, s0, nseg, *rgb])
    np.savez_compressed(HERE / "random_spheres_paths_400w_d16_seed1.npz", paths=np.array(rows, dtype=np.float64),
                        geom=g, mat=m, kind=k, camera=cam.as_array())
    # a small crop-sized image of the random scene for a fast oracle regression check
    img, _, st = O.render(g, m, k, cam.as_array(), 64, 4, max_depth=16, seed=1)
    np.savez_compressed(HERE / "random_spheres_64x36_4spp_d16_seed1.npz", image=img,
                        ray_segments=np.uint64(st["ray_segments"]))
    # RNG stream vectors: first 8 uniforms of three paths
    streams = np.stack([O.path_stream(1, p, s, e, 8) for p, s, e in [(0, 0, 0), (12345, 7, 1), (2073599, 999, 50)]])
    np.save(HERE / "philox_path_streams_seed1.npy", streams)
    # the reference's own smoke render, in its own element type: render(scene_2_spheres(Float64), default_cam, 96, 16)
    # (test/runtests.jl:190-194), depth 16 -- pins the Float64 instantiation of the oracle and of the CUDA path
    g, m, k = R.flatten_scene(R.scene_2_spheres(elem_type=np.float64), np.float64)
    cam = R.t_default_cam(np.float64)
    img, _, st = O.render(g, m, k, cam.as_array(), 96, 16, max_depth=16, seed=1, n_threads=1, f64=True)
    np.savez_compressed(HERE / "runtests194_scene_2_spheres_f64_96x54_16spp_d16_seed1.npz", image=img, geom=g, mat=m, kind=k,
                        camera=cam.as_array(), ray_segments=np.uint64(st["ray_segments"]))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
