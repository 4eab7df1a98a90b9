"""Structural checks of the oracle's render() restatement (src/render.jl:8-44).  CPU only, seconds."""
import numpy as np
import pytest


def _cam(rtw, name="default"):
    return {"default": rtw.t_default_cam, "cam1": rtw.t_cam1, "cam2": rtw.t_cam2}[name]().as_array()


def test_image_height(oracle, rtw):
    # image_width div (16//9), src/render.jl:11-12 ; SURVEY 3.1: 96->54, 400->225, 1920->1080, 200->112
    for w, h in [(96, 54), (400, 225), (1920, 1080), (200, 112), (1, 0), (16, 9), (17, 9)]:
        assert oracle.load().rtwo_image_height(w) == h == rtw.image_height(w)


def test_empty_scene_is_sky_closed_form(oracle, rtw):
    # no spheres: every path is one segment that misses => pixel = sqrt(mean skycolor) (src/ray_color.jl:36)
    geom = np.zeros((0, 4), np.float32)
    img, lin, st = oracle.render(geom, geom, np.zeros(0, np.uint32), _cam(rtw), 32, 1, n_threads=1, want_linear=True)
    assert st["ray_segments"] == st["paths"] == 32 * 18
    cam = rtw.t_default_cam()
    W, H = 32, 18
    for (i, j) in [(0, 0), (5, 7), (17, 31)]:
        u = np.float32((j + 1) / W)
        v = np.float32((H - 1 - i) / H)
        d = (cam.lower_left_corner + u * cam.horizontal + v * cam.vertical - cam.origin).astype(np.float64)
        d /= np.linalg.norm(d)
        t = 0.5 * (d[1] + 1.0)
        exp = (1 - t) * np.array([1.0, 1.0, 1.0]) + t * np.array([0.5, 0.7, 1.0])
        assert np.allclose(lin[i, j], exp, atol=2e-6)
        assert np.allclose(img[i, j], np.sqrt(exp), atol=2e-6)


def test_thread_count_independence_and_row_subsets(oracle, rtw, scenes):
    g, m, k = scenes["four"]
    a, _, sa = oracle.render(g, m, k, _cam(rtw), 64, 4, max_depth=8, n_threads=1)
    b, _, sb = oracle.render(g, m, k, _cam(rtw), 64, 4, max_depth=8, n_threads=5)
    assert np.array_equal(a, b) and sa["ray_segments"] == sb["ray_segments"]
    # interleaved row subsets reproduce exactly the same pixels (the multi-GPU split)
    parts = np.zeros_like(a)
    seg = 0
    for r in range(3):
        c, _, sc = oracle.render(g, m, k, _cam(rtw), 64, 4, max_depth=8, row_start=r, row_stride=3)
        parts[r::3] = c[r::3]
        seg += sc["ray_segments"]
    assert np.array_equal(parts, a) and seg == sa["ray_segments"]


def test_same_seed_same_image_other_seed_differs(oracle, rtw, scenes):
    # reseed!() at the top of render (src/render.jl:21): identical image every call
    g, m, k = scenes["two"]
    a, _, _ = oracle.render(g, m, k, _cam(rtw), 48, 4, seed=1)
    b, _, _ = oracle.render(g, m, k, _cam(rtw), 48, 4, seed=1)
    c, _, _ = oracle.render(g, m, k, _cam(rtw), 48, 4, seed=2)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_depth_semantics(oracle, rtw, scenes):
    # depth<=0 => black without a hit test (src/ray_color.jl:15-17); depth D => at most D segments per path
    g, m, k = scenes["two"]
    img, _, st = oracle.render(g, m, k, _cam(rtw), 32, 2, max_depth=0)
    assert st["ray_segments"] == 0 and not img.any()
    img, _, st = oracle.render(g, m, k, _cam(rtw), 32, 2, max_depth=1)
    assert st["ray_segments"] == st["paths"]
    for depth in (2, 5):
        _, _, st = oracle.render(g, m, k, _cam(rtw), 32, 2, max_depth=depth)
        assert st["paths"] <= st["ray_segments"] <= depth * st["paths"]


def test_energy_bounds_and_first_sample_centred(oracle, rtw, scenes):
    g, m, k = scenes["random"]
    img, lin, _ = oracle.render(g, m, k, _cam(rtw, "cam1"), 64, 8, want_linear=True)
    assert np.isfinite(img).all() and lin.min() >= 0.0 and lin.max() <= 1.0 + 1e-6
    # spp=1: no jitter (src/render.jl:30-31); only the disk sample and the bounces consume randomness
    one, _, _ = oracle.render(g, m, k, _cam(rtw, "cam1"), 64, 1)
    rgb, nseg = oracle.path(g, m, k, _cam(rtw, "cam1"), 64, 10, 20, 0)
    assert np.allclose(one[10, 20], np.sqrt(rgb).astype(np.float32), atol=0) and nseg >= 1


def test_mirror_symmetry_of_symmetric_scene(oracle, rtw, scenes):
    # scene_blue_red_spheres geometry (src/scenes.jl:41-47) with both albedos grey is mirror-symmetric in x.
    # u = j/W has no half-pixel offset (src/render.jl:26), so column j0 mirrors onto column W-2-j0.
    g, m, k = scenes["bluered"]
    m = m.copy()
    m[:, :3] = 0.5
    W = 64
    img, lin, _ = oracle.render(g, m, k, _cam(rtw), W, 128, max_depth=8, want_linear=True)
    left = lin[:, 0:W - 1]
    right = lin[:, [W - 2 - j for j in range(W - 1)]]
    assert float(np.abs(left - right).mean()) < 0.02
    assert float(np.abs(left.mean(axis=(0, 1)) - right.mean(axis=(0, 1))).max()) < 1e-3 + 0.01


def test_xoroshiro_stream_converges_to_philox_stream(oracle, rtw, scenes):
    # Tier 3 of SURVEY 8c: the reference-shaped sequential stream and the path-keyed Philox stream estimate
    # the same image; the difference shrinks like 1/sqrt(spp)
    g, m, k = scenes["two"]
    diffs = []
    for spp in (4, 64):
        a, _, _ = oracle.render(g, m, k, _cam(rtw), 48, spp, max_depth=8, rng_mode=oracle.RNG_PHILOX)
        b, _, _ = oracle.render(g, m, k, _cam(rtw), 48, spp, max_depth=8, rng_mode=oracle.RNG_XOROSHIRO, n_threads=1)
        diffs.append(float(np.abs(a - b).mean()))
    assert diffs[1] < 0.45 * diffs[0]  # 16x the samples => ~4x smaller
    assert diffs[1] < 0.02


def test_float64_instantiation_close_to_float32(oracle, rtw, scenes):
    # the reference is generic over T; same stream shape => visually identical estimate at moderate spp
    g, m, k = scenes["two"]
    a, _, _ = oracle.render(g, m, k, _cam(rtw), 32, 32, max_depth=8)
    cam64 = rtw.default_camera((0, 0, 0), elem_type=np.float64).as_array()
    b, _, _ = oracle.render(g, m, k, cam64, 32, 32, max_depth=8, f64=True)
    assert b.dtype == np.float64 and float(np.abs(a - b).mean()) < 0.05


def test_bad_arguments(oracle, rtw, scenes):
    g, m, k = scenes["two"]
    for kw in ({"n_samples": 0}, {"image_width": 0}, {"max_depth": -1}, {"row_stride": 0}, {"rng_mode": 7}):
        args = {"image_width": 32, "n_samples": 1}
        args.update(kw)
        with pytest.raises(ValueError):
            oracle.render(g, m, k, _cam(rtw), args.pop("image_width"), args.pop("n_samples"), **args)


def test_thread_count_and_work_sharing_do_not_change_the_image(oracle, rtw):
    # production stream: every draw is addressed, and the threads are dealt 16-pixel blocks round-robin -- any thread count
    # gives the same bits, also for a one-row slice of a wide image (what the headline parity tests render)
    g, m, k = rtw.flatten_scene(rtw.scene_4_spheres())
    cam = rtw.t_default_cam().as_array()
    a, _, sa = oracle.render(g, m, k, cam, 200, 3, max_depth=8, seed=5, n_threads=1)
    b, _, sb = oracle.render(g, m, k, cam, 200, 3, max_depth=8, seed=5, n_threads=7)
    assert np.array_equal(a, b) and sa["ray_segments"] == sb["ray_segments"]
    row, _, sr = oracle.render(g, m, k, cam, 200, 3, max_depth=8, seed=5, n_threads=5, row_start=60, row_stride=112)
    assert np.array_equal(row[60], a[60]) and sr["paths"] == 200 * 3


def test_path_trace_records_the_segments_of_a_path(oracle, rtw):
    g, m, k = rtw.flatten_scene(rtw.scene_2_spheres())
    cam = rtw.t_default_cam().as_array()
    rgb, seg = oracle.path(g, m, k, cam, 96, 30, 48, 1, max_depth=6, seed=1)
    rgb2, tr = oracle.path_trace(g, m, k, cam, 96, 30, 48, 1, max_depth=6, seed=1)
    assert np.array_equal(rgb, rgb2) and len(tr) == seg
    assert all(abs(np.dot(r[3:6], r[3:6]) - 1.0) < 1e-5 for r in tr)        # unit directions in this scene
    assert tr[-1][6] == -1 or len(tr) == 6                                    # ends in the sky or at the depth limit
    assert all(r[6] in (-1.0, 0.0, 1.0) for r in tr)
