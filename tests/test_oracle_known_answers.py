"""Pins the CPU oracle against every known answer the reference's own tests / notebook hold for the
hot path.  CPU only."""
import math

import numpy as np
import pytest


def test_reflect_known_answer(oracle):
    # test/runtests.jl:180 and src/pluto_RayTracingWeekend.jl:382 -- exact equality in Float64
    out = oracle.reflect([0.6, -0.8, 0.0], [0.0, 1.0, 0.0])
    assert out.tolist() == [0.6, 0.8, 0.0]


def test_refract_known_answers(oracle):
    # src/pluto_RayTracingWeekend.jl:603-615 (commented copy: test/runtests.jl:203-211)
    d, n = [0.6, -0.8, 0.0], [0.0, 1.0, 0.0]
    assert oracle.refract(d, n, 1.0).tolist() == [0.6, -0.8, 0.0]  # unchanged angle: exact ==
    assert np.allclose(oracle.refract(d, n, 2.0), [0.87519, -0.483779, 0.0], atol=1e-3, rtol=0)  # wider
    assert np.allclose(oracle.refract(d, n, 0.5), [0.3, -0.953939, 0.0], atol=1e-3, rtol=0)  # narrower
    # the Float32 instantiation agrees with the Float64 one to Float32 precision
    for ratio in (1.0, 2.0, 0.5):
        assert np.allclose(oracle.refract(d, n, ratio, np.float32), oracle.refract(d, n, ratio), atol=1e-6)


def test_near_zero_known_answer(oracle):
    # test/runtests.jl:131  @test !near_zero(SA[0.4,0.5,0.1]);  src/vec.jl:20 threshold on the SQUARED length
    assert not oracle.near_zero([0.4, 0.5, 0.1])
    assert oracle.near_zero([1e-3, 1e-3, 1e-3])  # 3e-6 < 1e-5
    assert not oracle.near_zero([3e-3, 1e-3, 1e-3])  # 1.1e-5
    assert oracle.near_zero([1e-3, 1e-3, 1e-3], np.float32)


def test_hit_sphere_matches_hit_sphere2_closed_form(oracle):
    # test/runtests.jl:99-111: t = (-b - sqrt(b^2 - 4ac)) / 2a with b = 2 oc.d, a = 1
    rng = np.random.default_rng(7)
    center, radius = np.array([0.0, 0.0, -1.0]), 0.5
    hits = 0
    for _ in range(500):
        d = rng.normal(size=3) * [0.3, 0.3, 1.0]
        d[2] = -abs(d[2])
        d /= np.linalg.norm(d)
        o = rng.normal(size=3) * 0.05
        oc = o - center
        b = 2 * oc.dot(d)
        c = oc.dot(oc) - radius ** 2
        disc = b * b - 4 * c
        got = oracle.hit_sphere(center, radius, o, d, 1e-4, math.inf)
        if disc < 0:
            assert got is None
            continue
        t_ref = (-b - math.sqrt(disc)) / 2
        if t_ref < 1e-4:
            continue
        hits += 1
        t, p, n, front = got
        assert t == pytest.approx(t_ref, rel=1e-12, abs=1e-12)
        assert np.allclose(p, o + t * d, atol=1e-12)
        assert np.allclose(n, (p - center) / radius, atol=1e-12) and front
    assert hits > 100


def test_hit_sphere_far_root_inside_and_negative_radius(oracle):
    # src/hit.jl:23-29: origin inside the sphere => near root < tmin => far root; front_face False, normal flipped
    t, p, n, front = oracle.hit_sphere([0, 0, 0], 1.0, [0, 0, 0], [0, 0, -1], 1e-4, math.inf)
    assert t == 1.0 and not front and n.tolist() == [0.0, 0.0, 1.0]
    # negative radius (hollow glass, src/scenes.jl:35-36) flips the outward normal via the division (src/hit.jl:33)
    t, p, n, front = oracle.hit_sphere([0, 0, -2], -0.5, [0, 0, 0], [0, 0, -1], 1e-4, math.inf)
    assert t == 1.5 and not front and n.tolist() == [0.0, 0.0, 1.0]
    # tmax is inclusive (src/hit.jl:24: `tmax < root` rejects) -- a later sphere at the same t wins (src/hit.jl:44-46)
    assert oracle.hit_sphere([0, 0, -2], 0.5, [0, 0, 0], [0, 0, -1], 1e-4, 1.5) is not None
    assert oracle.hit_sphere([0, 0, -2], 0.5, [0, 0, 0], [0, 0, -1], 1e-4, 1.4999) is None
    # behind the ray
    assert oracle.hit_sphere([0, 0, 2], 0.5, [0, 0, 0], [0, 0, -1], 1e-4, math.inf) is None


def test_skycolor_closed_form(oracle):
    # src/ray_color.jl:1-6: (1-t)*white + t*skyblue with t = 0.5*(dir.y+1)
    assert np.allclose(oracle.skycolor([0, 1, 0]), [0.5, 0.7, 1.0])
    assert np.allclose(oracle.skycolor([0, -1, 0]), [1.0, 1.0, 1.0])
    assert np.allclose(oracle.skycolor([1, 0, 0]), [0.75, 0.85, 1.0])


def test_reflectance_schlick(oracle):
    # src/light.jl:19-25: r0 + (1-r0)(1-cos)^5, r0 = ((1-q)/(1+q))^2
    for cos_t, q in [(1.0, 1.5), (0.0, 1.5), (0.3, 1 / 1.5), (0.9, 0.7)]:
        r0 = ((1 - q) / (1 + q)) ** 2
        assert oracle.reflectance(cos_t, q) == pytest.approx(r0 + (1 - r0) * (1 - cos_t) ** 5, rel=1e-12)
        assert oracle.reflectance(cos_t, q, np.float32) == pytest.approx(r0 + (1 - r0) * (1 - cos_t) ** 5, rel=1e-5)


def test_philox_known_answer_vectors(oracle):
    # Random123 kat_vectors for philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    # ... and for philox4x32-7, the round count of the production stream (RTWO_PHILOX_ROUNDS / kPhiloxRounds)
    assert oracle.philox4x32_7([0, 0, 0, 0], [0, 0]) == [0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48]
    assert oracle.philox4x32_7([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x5207DDC2, 0x45165E59, 0x4D8EE751, 0x8C52F662]
    assert oracle.philox4x32_7([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0x4DFCCABA, 0x190A87F0, 0xC47362BA, 0xB6B5242A]


def test_path_stream_layout(oracle):
    # draw n of an event = word (n mod 4) of block (n div 4); counter = (block, sample, pixel, event); f32 = (w>>9)*2^-23
    seed, pixel, sample = 0x1234567800000001, 4321, 17
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for event in (0, 1, 7):
        got = oracle.path_stream(seed, pixel, sample, event, 12)
        exp = []
        for blk in range(3):
            exp += [np.float32(w >> 9) * np.float32(2.0 ** -23) for w in oracle.philox4x32_7([blk, sample, pixel, event], key)]
        assert got.tolist() == [float(x) for x in exp]
        assert got.min() >= 0.0 and got.max() < 1.0


def test_addressed_draws_drive_the_path(oracle, rtw):
    # the documented addresses: event 0 draws 0,1 = jitter, disk attempt k = draws 2+2k,3+2k;
    # event e draws 4a..4a+2 = ball attempt a.  Re-derive one primary ray by hand from the stream.
    cam = rtw.t_cam1()
    W, H, i0, j0, s0, seed = 64, 36, 5, 9, 3, 1
    st = oracle.path_stream(seed, i0 * W + j0, s0, 0, 16)
    u = np.float32((j0 + 1) / W) + st[0] / np.float32(W)
    v = np.float32((H - 1 - i0) / H) + st[1] / np.float32(H)
    k = 0
    while True:
        px, py = np.float32(st[2 + 2 * k] * 2 - 1), np.float32(st[3 + 2 * k] * 2 - 1)
        if px * px + py * py <= 1:
            break
        k += 1
    rd = cam.lens_radius * np.array([px, py], np.float32)
    off = cam.u * rd[0] + cam.v * rd[1]
    d = (cam.lower_left_corner + u * cam.horizontal + v * cam.vertical - cam.origin - off).astype(np.float64)
    d /= np.linalg.norm(d)
    # empty scene: the path's colour is the sky colour of exactly this direction
    empty = (np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), np.zeros(0, np.uint32))
    rgb, nseg = oracle.path(*empty, cam.as_array(), W, i0, j0, s0, seed=seed)
    t = 0.5 * (d[1] + 1.0)
    assert nseg == 1 and np.allclose(rgb, (1 - t) * np.ones(3) + t * np.array([0.5, 0.7, 1.0]), atol=1e-5)


def test_xoroshiro_matches_host_mirror(oracle, rtw):
    # two independent restatements (C and Python) of the published algorithm agree; NOT verified against Julia
    for seed in (1, 2, 16):
        g = rtw.Xoroshiro128Plus(seed)
        assert [g.next_u64() for _ in range(64)] == [int(x) for x in oracle.xoroshiro_u64(seed, 64)]
        g = rtw.Xoroshiro128Plus(seed)
        assert [float(g.rand(np.float32)) for _ in range(64)] == oracle.xoroshiro_f32(seed, 64).tolist()
