"""The oracle reproduces the committed golden fixtures bit for bit (guards the checker itself).  CPU only."""
from pathlib import Path

import numpy as np

GOLD = Path(__file__).parent / "golden"


def test_cfg1_fixture(oracle, rtw):
    gold = np.load(GOLD / "cfg1_scene_2_spheres_96x54_16spp_d4_seed1.npz")
    g, m, k = rtw.flatten_scene(rtw.scene_2_spheres())
    assert np.array_equal(g, gold["geom"]) and np.array_equal(m, gold["mat"]) and np.array_equal(k, gold["kind"])
    assert np.array_equal(rtw.t_default_cam().as_array(), gold["camera"])
    img, _, st = oracle.render(g, m, k, gold["camera"], 96, 16, max_depth=4, seed=1, n_threads=3)
    assert np.array_equal(img, gold["image"]) and st["ray_segments"] == int(gold["ray_segments"])


def test_random_scene_fixtures(oracle, rtw):
    gold = np.load(GOLD / "random_spheres_paths_400w_d16_seed1.npz")
    rtw.reseed()
    g, m, k = rtw.flatten_scene(rtw.scene_random_spheres())
    assert np.array_equal(g, gold["geom"]) and np.array_equal(m, gold["mat"]) and np.array_equal(k, gold["kind"])
    for row in gold["paths"][:64]:
        i0, j0, s0, nseg = (int(x) for x in row[:4])
        rgb, n = oracle.path(g, m, k, gold["camera"], 400, i0, j0, s0, max_depth=16, seed=1)
        assert n == nseg and rgb.tolist() == row[4:].tolist()
    small = np.load(GOLD / "random_spheres_64x36_4spp_d16_seed1.npz")
    img, _, st = oracle.render(g, m, k, gold["camera"], 64, 4, max_depth=16, seed=1)
    assert np.array_equal(img, small["image"]) and st["ray_segments"] == int(small["ray_segments"])


def test_stream_fixture(oracle):
    gold = np.load(GOLD / "philox_path_streams_seed1.npy")
    for row, (p, s, e) in zip(gold, [(0, 0, 0), (12345, 7, 1), (2073599, 999, 50)]):
        assert np.array_equal(oracle.path_stream(1, p, s, e, 8), row)


def test_float64_fixture_of_the_reference_smoke_render(oracle, rtw):
    # test/runtests.jl:190-194: render(scene_2_spheres(Float64), default_camera, 96, 16)
    gold = np.load(GOLD / "runtests194_scene_2_spheres_f64_96x54_16spp_d16_seed1.npz")
    g, m, k = rtw.flatten_scene(rtw.scene_2_spheres(elem_type=np.float64), np.float64)
    assert g.dtype == np.float64 and np.array_equal(g, gold["geom"]) and np.array_equal(m, gold["mat"])
    img, _, st = oracle.render(g, m, k, gold["camera"], 96, 16, max_depth=16, seed=1, n_threads=2, f64=True)
    assert img.dtype == np.float64 and np.array_equal(img, gold["image"]) and st["ray_segments"] == int(gold["ray_segments"])
