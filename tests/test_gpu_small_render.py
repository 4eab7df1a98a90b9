"""The latency path (csrc/rtw_small.cu): small renders through rtw_render / rtw_render_scene run as ONE kernel launch.
Same image bits as the persistent kernel and as the oracle, same ray-segment counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fast(rtw):
    with rtw.Renderer([0]) as r:  # library defaults: RTW_OPT_SMALL_RENDER = 1
        yield r


CASES = [("two", "default", 96, 16, 4, 1),     # BASELINE configs[0] / test/runtests.jl:194
         ("two", "default", 96, 1, 16, 1),     # src/proto/proto.jl:87-89
         ("random", "cam1", 96, 1, 16, 1),     # src/proto/proto.jl:142-144
         ("four", "default", 77, 3, 16, 3),    # group of 4 lanes, one idle
         ("diel", "cam2", 64, 33, 16, 7),      # 32-lane groups looping over the samples, depth of field
         ("bubble", "default", 50, 90, 50, 2),
         ("bluered", "cam2", 33, 7, 1, 5),     # depth 1
         ("random", "cam1", 40, 5, 50, 9)]


@pytest.mark.parametrize("name,cam_name,W,spp,depth,seed", CASES)
def test_small_render_matches_oracle_and_persistent_kernel(rtw, oracle, renderer, fast, scenes, name, cam_name, W, spp, depth, seed):
    cam = {"default": rtw.t_default_cam, "cam1": rtw.t_cam1, "cam2": rtw.t_cam2}[cam_name]()
    img = np.array(fast.render(cam, W, spp, max_depth=depth, seed=seed, scene=scenes[name]))
    st = dict(fast.last_stats)
    assert st["kernel_launches"] == 1, "the render did not take the single-launch path"
    big = np.array(renderer.render(cam, W, spp, max_depth=depth, seed=seed, scene=scenes[name]))
    sb = dict(renderer.last_stats)
    assert sb["kernel_launches"] >= 3
    assert np.array_equal(img, big)  # integer pixel sums: the same bits whichever kernel added them
    assert st["ray_segments"] == sb["ray_segments"] and st["paths"] == sb["paths"]
    ref, _, ost = oracle.render(*scenes[name], cam.as_array(), W, spp, max_depth=depth, seed=seed)
    assert st["ray_segments"] == ost["ray_segments"]
    assert float(np.abs(img.astype(np.float64) - ref).max()) < 1e-6


def test_small_render_back_to_back_and_after_large_renders(rtw, fast, scenes):
    # the kernel re-zeroes its device counters itself; a persistent-kernel render in between leaves them dirty
    cam = rtw.t_default_cam()
    fast.set_scene(scenes["two"])
    a = np.array(fast.render(cam, 96, 16, max_depth=4))
    seg = fast.last_stats["ray_segments"]
    for _ in range(3):
        assert np.array_equal(np.array(fast.render(cam, 96, 16, max_depth=4)), a)
        assert fast.last_stats["ray_segments"] == seg and fast.last_stats["kernel_launches"] == 1
    big = fast.render(cam, 640, 64, max_depth=4)  # 14.7 M paths: persistent kernel
    assert fast.last_stats["kernel_launches"] >= 3 and big.shape == (360, 640, 3)
    assert np.array_equal(np.array(fast.render(cam, 96, 16, max_depth=4)), a)
    assert fast.last_stats["ray_segments"] == seg and fast.last_stats["kernel_launches"] == 1
    # the same scene passed again with the call (what the drop-in render(scene, cam, ...) does) is not re-uploaded
    assert np.array_equal(np.array(fast.render(cam, 96, 16, max_depth=4, scene=scenes["two"])), a)
    # max_depth 0: black image, no segments
    z = np.array(fast.render(cam, 32, 2, max_depth=0))
    assert not z.any() and fast.last_stats["ray_segments"] == 0


def test_small_render_is_off_for_variants_and_other_modes(rtw, fast, scenes):
    cam = rtw.t_default_cam()
    fast.set_scene(scenes["two"])
    fast.set_option(rtw.RTW_OPT_MODE, rtw.RTW_MODE_GRID)
    fast.render(cam, 96, 1)
    assert fast.last_stats["kernel_launches"] >= 3
    fast.set_option(rtw.RTW_OPT_MODE, rtw.RTW_MODE_FUSED)
    fast.set_option(rtw.RTW_OPT_SMALL_RENDER, 0)
    fast.render(cam, 96, 1)
    assert fast.last_stats["kernel_launches"] >= 3
    fast.set_option(rtw.RTW_OPT_SMALL_RENDER, 1)
    fast.render(cam, 96, 1)
    assert fast.last_stats["kernel_launches"] == 1


@pytest.mark.parametrize("half", [11, 3, 26, 158])
def test_device_scene_generator_reproduces_the_host_builder(rtw, fast, half):
    # scene_random_spheres (src/scenes.jl:49-84) built on the device: the same list, bit for bit, as the sequential host
    # loop (host.py mirrors the reference's loop and its Xoroshiro128Plus stream), and the same generator state after it
    rtw.reseed()
    for _ in range(5):
        rtw.trand()  # "draws from the calling thread's TRNG in whatever state it is"
    x0, y0 = rtw.TRNG[0].x, rtw.TRNG[0].y
    g_ref, m_ref, k_ref = rtw.flatten_scene(rtw.scene_random_spheres(half_extent=half))
    after = (rtw.TRNG[0].x, rtw.TRNG[0].y)
    rtw.TRNG[0].x, rtw.TRNG[0].y = x0, y0
    g, m, k = fast.generate_random_spheres(half)
    assert (rtw.TRNG[0].x, rtw.TRNG[0].y) == after
    assert len(k) == len(k_ref) and fast.n_spheres == len(k)
    assert np.array_equal(k, k_ref)
    assert np.array_equal(g.view(np.uint32), g_ref.view(np.uint32))
    assert np.array_equal(m.view(np.uint32), m_ref.view(np.uint32))
    # the generated list is installed: rendering it equals rendering the host-built one
    cam = rtw.t_cam1()
    a = np.array(fast.render(cam, 64, 2, max_depth=8))
    b = np.array(fast.render(cam, 64, 2, max_depth=8, scene=(g_ref, m_ref, k_ref)))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name,cam_name,W,spp,depth", [("two", "default", 96, 16, 16),   # test/runtests.jl:190-194, as is
                                                       ("two", "default", 96, 1, 16),
                                                       ("diel", "cam2", 64, 33, 16), ("bubble", "default", 50, 7, 50),
                                                       ("random", "cam1", 40, 3, 12)])
def test_small_render_float64(rtw, oracle, renderer, fast, name, cam_name, W, spp, depth):
    # the reference's own smoke test renders Float64: the Float64 latency path against the Float64 oracle and against the
    # persistent Float64 kernel
    from test_gpu_parity import _f64_scene
    scene = rtw.flatten_scene(_f64_scene(rtw, name), np.float64)
    cam = {"default": rtw.t_default_cam, "cam1": rtw.t_cam1, "cam2": rtw.t_cam2}[cam_name](np.float64)
    img = np.array(fast.render(cam, W, spp, max_depth=depth, seed=4, scene=scene))
    st = dict(fast.last_stats)
    assert st["kernel_launches"] == 1 and img.dtype == np.float64
    big = np.array(renderer.render(cam, W, spp, max_depth=depth, seed=4, scene=scene))
    assert renderer.last_stats["kernel_launches"] >= 2 and renderer.last_stats["ray_segments"] == st["ray_segments"]
    assert np.array_equal(img, big)
    ref, _, ost = oracle.render(*scene, cam.as_array(), W, spp, max_depth=depth, seed=4, f64=True)
    assert st["ray_segments"] == ost["ray_segments"]
    assert float(np.abs(img - ref).max()) <= 1e-9
