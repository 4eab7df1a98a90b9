"""Host-side mirror of camera.jl / scenes.jl / structs.jl (code that stays on the host in the design)."""
import numpy as np
import pytest


def test_default_camera_matches_reference_formulas(rtw):
    # default_camera(SA[0f0,0f0,0f0]) -- test/runtests.jl:182: vfov 90, aspect 16/9, focus 1, no aperture
    c = rtw.t_default_cam()
    assert c.origin.tolist() == [0, 0, 0] and c.lens_radius == 0
    assert np.allclose(c.horizontal, [2 * 16 / 9, 0, 0]) and np.allclose(c.vertical, [0, 2, 0])
    assert np.allclose(c.lower_left_corner, [-16 / 9, -1, -1])
    assert c.u.tolist() == [1, 0, 0] and c.v.tolist() == [0, 1, 0] and c.w.tolist() == [0, 0, 1]
    # t_cam1 (src/proto/proto.jl:19): orthonormal basis, lens radius = aperture/2, plane at focus distance
    c = rtw.t_cam1()
    assert c.lens_radius == np.float32(0.05)
    for a, b in [(c.u, c.v), (c.u, c.w), (c.v, c.w)]:
        assert abs(float(np.dot(a, b))) < 1e-6
    centre = c.lower_left_corner + c.horizontal / 2 + c.vertical / 2
    assert np.allclose(centre, c.origin - 10.0 * c.w, atol=1e-5)
    assert np.linalg.norm(c.vertical) == pytest.approx(2 * 10 * np.tan(np.radians(10)), rel=1e-6)
    c64 = rtw.t_cam1(np.float64)
    assert c64.as_array().dtype == np.float64 and np.allclose(c64.as_array(), c.as_array(), atol=1e-6)


def test_scene_builders(rtw):
    assert len(rtw.scene_2_spheres()) == 2 and len(rtw.scene_4_spheres()) == 4
    assert len(rtw.scene_diel_spheres()) == 4 and len(rtw.scene_blue_red_spheres()) == 2
    assert rtw.scene_diel_spheres(-0.4)[2].radius == np.float32(-0.4)
    rtw.reseed()
    a = rtw.flatten_scene(rtw.scene_random_spheres())
    rtw.reseed()
    b = rtw.flatten_scene(rtw.scene_random_spheres())
    for x, y in zip(a, b):
        assert np.array_equal(x, y)  # reseed!() => identical scene
    geom, mat, kind = a
    n = len(kind)
    assert 440 <= n <= 488  # 1 ground + <=484 small + 3 big (src/scenes.jl:49-84)
    assert geom[0].tolist() == [0, -1000, -1, 1000] and kind[0] == 0
    assert geom[-3:, 3].tolist() == [1, 1, 1] and kind[-3:].tolist() == [2, 0, 1]
    small = slice(1, n - 3)
    assert np.all(geom[small, 3] == np.float32(0.2)) and np.all(geom[small, 1] == np.float32(0.2))
    # the (4, 0.2, 0) exclusion zone, src/scenes.jl:61
    assert np.all(np.linalg.norm(geom[small, :3] - np.array([4, 0.2, 0], np.float32), axis=1) >= 0.9 - 1e-6)
    frac = np.bincount(kind[small], minlength=3) / (n - 4)
    assert 0.7 < frac[0] < 0.9 and 0.08 < frac[1] < 0.22 and 0.01 < frac[2] < 0.1
    metal = kind == 1
    assert np.all(mat[metal][:-1, :3] >= 0.5) and np.all(mat[metal][:-1, 3] < 5.0)  # fuzz in [0,5), src/scenes.jl:70
    assert np.all(mat[kind == 2][:, 3] == 1.5)


def test_large_synthetic_scene(rtw):
    rtw.reseed()
    geom, mat, kind = rtw.flatten_scene(rtw.scene_random_spheres(half_extent=20))
    assert 1500 < len(kind) <= 1604


def test_flatten_rejects_unknown(rtw):
    with pytest.raises(TypeError):
        rtw.flatten_scene([object()])
    with pytest.raises(TypeError):
        rtw.flatten_scene([rtw.Sphere(rtw.Vec3(0, 0, 0), 1.0, "wood")])
    with pytest.raises(rtw.RtwError):
        rtw.api._camera_struct(rtw.t_cam1(np.float64))  # the Float32 entry points refuse a Float64 camera (render() dispatches)


def test_trand_range_and_reseed(rtw):
    rtw.reseed()
    a = [float(rtw.trand()) for _ in range(1000)]
    rtw.reseed()
    b = [float(rtw.trand()) for _ in range(1000)]
    assert a == b and 0.0 <= min(a) and max(a) < 1.0 and 0.4 < np.mean(a) < 0.6
    x = rtw.random_between(np.float32(0.5), np.float32(1.0))
    assert 0.5 <= x < 1.0
