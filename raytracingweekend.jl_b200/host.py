"""Host-side mirror of the parts of RayTracingWeekend.jl that STAY host code in the B200 design:
structs (src/structs.jl), camera set-up (src/camera.jl:1-41), per-thread RNG used by the scene builders
(src/init.jl, src/rand.jl) and the scene builders themselves (src/scenes.jl).

In the drop-in deployment this layer is the unchanged Julia package; Julia is not
installed in this image, so the same API is mirrored here in Python (same names, argument order and
meaning) so that tests and the benchmark drive the C-ABI exactly as the Julia shim does.

Everything here is set-up code that runs once per scene -- none of it is on the per-sample hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

F32 = np.float32
F64 = np.float64

# ----------------------------------------------------------------------------- vec.jl


def Vec3(x, y, z, elem_type=F32) -> np.ndarray:
    """Vec3{T} = SVector{3,T}, src/vec.jl:3"""
    return np.array([x, y, z], dtype=elem_type)


def squared_length(v: np.ndarray):
    """src/vec.jl:19"""
    return v.dtype.type(np.dot(v, v))


def near_zero(v: np.ndarray) -> bool:
    """src/vec.jl:20"""
    return float(np.dot(v.astype(F64), v.astype(F64))) < 1e-5 if v.dtype == F64 else float(squared_length(v)) < 1e-5


def _normalize(v: np.ndarray) -> np.ndarray:
    t = v.dtype.type
    return (v * (t(1) / t(np.sqrt(t(np.dot(v, v)))))).astype(v.dtype)


# ----------------------------------------------------------------------------- init.jl / rand.jl

_MASK64 = (1 << 64) - 1


class Xoroshiro128Plus:
    """RandomNumbers.jl 1.5.3 Xoroshiro128Plus(seed) -- restated from the published xoroshiro128+
    algorithm (constants 55, 14, 36; SplitMix64 seed expansion; one warm-up step).  UNVERIFIED against
    Julia (the package is not vendored in the reference and Julia is not installed here).  It only feeds
    the host-side scene builders; the device stream is Philox."""

    def __init__(self, seed: int):
        self.seed(seed)

    @staticmethod
    def _splitmix64(s: int) -> Tuple[int, int]:
        s = (s + 0x9E3779B97F4A7C15) & _MASK64
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
        return s, z ^ (z >> 31)

    def seed(self, seed: int) -> None:
        s = seed & _MASK64
        s, self.x = self._splitmix64(s)
        s, self.y = self._splitmix64(s)
        self.next_u64()

    def next_u64(self) -> int:
        s0, s1 = self.x, self.y
        r = (s0 + s1) & _MASK64
        s1 ^= s0
        self.x = (((s0 << 55) | (s0 >> 9)) & _MASK64) ^ s1 ^ ((s1 << 14) & _MASK64)
        self.y = ((s1 << 36) | (s1 >> 28)) & _MASK64
        return r

    def rand(self, elem_type=F32):
        u = self.next_u64()
        if elem_type == F32:
            return F32((u & 0xFFFFFFFF) >> 9) * F32(2.0 ** -23)
        return F64(u >> 12) * F64(2.0 ** -52)


# const TRNG = Xoroshiro128Plus[] ; one per thread, seeded with the thread id (src/init.jl:2-12).
# The Python host is single-threaded: TRNG has exactly one entry, "thread 1".
TRNG: List[Xoroshiro128Plus] = [Xoroshiro128Plus(1)]


def reseed() -> None:
    """reseed!(), src/rand.jl:2"""
    for i, g in enumerate(TRNG):
        g.seed(i + 1)


def trand(elem_type=F32):
    """trand(T), src/rand.jl:10-13"""
    return TRNG[0].rand(elem_type)


def random_between(lo, hi):
    """random_between(min, max) = trand(T)*(max-min) + min, src/rand.jl:24"""
    t = type(lo)
    return t(trand(t) * (hi - lo) + lo)


# ----------------------------------------------------------------------------- structs.jl / material.jl


@dataclass(frozen=True)
class Lambertian:
    """src/material.jl:3-5"""

    albedo: np.ndarray


@dataclass(frozen=True)
class Metal:
    """src/material.jl:25-29 (fuzz defaults to 0)"""

    albedo: np.ndarray
    fuzz: float = 0.0


@dataclass(frozen=True)
class Dielectric:
    """src/material.jl:37-39"""

    ir: float


@dataclass(frozen=True)
class Sphere:
    """src/structs.jl:31-35"""

    center: np.ndarray
    radius: float
    mat: object


class HittableList(list):
    """const HittableList = Vector{Hittable}, src/structs.jl:10"""


KIND_LAMBERTIAN, KIND_METAL, KIND_DIELECTRIC = 0, 1, 2


def flatten_scene(scene: Sequence[Sphere], elem_type=F32) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Flatten a HittableList into the SoA arrays of the C-ABI (include/rtw_b200.h, rtw_set_scene):
    geom4 = n x {cx,cy,cz,r}, mat4 = n x {albedo rgb, fuzz|ir|0}, kind = n x u32.  List order is kept.
    This is the one piece of glue the Julia shim also needs (Sphere.mat is abstract => not isbits)."""
    n = len(scene)
    F32 = elem_type  # noqa: N806 -- element type of the flattened arrays (Float32 by default, Float64 for the f64 path)
    geom = np.zeros((n, 4), dtype=F32)
    mat = np.zeros((n, 4), dtype=F32)
    kind = np.zeros((n,), dtype=np.uint32)
    for i, s in enumerate(scene):
        if not isinstance(s, Sphere):
            raise TypeError(f"only Sphere hittables are supported (src/structs.jl:31), got {type(s).__name__}")
        geom[i, :3] = np.asarray(s.center, dtype=F32)
        geom[i, 3] = F32(s.radius)
        m = s.mat
        if isinstance(m, Lambertian):
            kind[i] = KIND_LAMBERTIAN
            mat[i, :3] = np.asarray(m.albedo, dtype=F32)
        elif isinstance(m, Metal):
            kind[i] = KIND_METAL
            mat[i, :3] = np.asarray(m.albedo, dtype=F32)
            mat[i, 3] = F32(m.fuzz)
        elif isinstance(m, Dielectric):
            kind[i] = KIND_DIELECTRIC
            mat[i, :3] = 1.0
            mat[i, 3] = F32(m.ir)
        else:
            raise TypeError(f"unsupported material {type(m).__name__} (Lambertian/Metal/Dielectric only)")
    return geom, mat, kind


# ----------------------------------------------------------------------------- camera.jl


@dataclass(frozen=True)
class Camera:
    """Camera{T}, src/camera.jl:1-10"""

    origin: np.ndarray
    lower_left_corner: np.ndarray
    horizontal: np.ndarray
    vertical: np.ndarray
    u: np.ndarray
    v: np.ndarray
    w: np.ndarray
    lens_radius: float
    elem_type: type = F32

    def as_array(self) -> np.ndarray:
        """22 x T in the field order of the Julia struct (the ABI layout)."""
        t = self.elem_type
        return np.concatenate(
            [self.origin, self.lower_left_corner, self.horizontal, self.vertical, self.u, self.v, self.w,
             np.array([self.lens_radius], dtype=t)]
        ).astype(t)


def _tand(x):
    """tand(x): tangent of an angle in degrees (Julia Base).  Exact at the multiples of 45 Julia special-cases."""
    t = type(x)
    xf = float(x)
    if xf % 180.0 == 45.0:
        return t(1)
    if xf % 180.0 == 135.0:
        return t(-1)
    if xf % 180.0 == 0.0:
        return t(0)
    return t(math.tan(math.radians(xf)))


def default_camera(lookfrom=(0, 0, 0), lookat=(0, 0, -1), vup=(0, 1, 0), vfov=90, aspect_ratio=16 / 9, aperture=0,
                   focus_dist=1, *, elem_type=F32) -> Camera:
    """default_camera(lookfrom, lookat, vup, vfov, aspect_ratio, aperture, focus_dist; elem_type),
    src/camera.jl:18-41.  All arithmetic in T, in the reference's order."""
    t = elem_type
    lookfrom = np.asarray(lookfrom, dtype=t)
    lookat = np.asarray(lookat, dtype=t)
    vup = np.asarray(vup, dtype=t)
    vfov, aspect_ratio, aperture, focus_dist = t(vfov), t(aspect_ratio), t(aperture), t(focus_dist)
    viewport_height = t(2) * _tand(t(vfov / t(2)))           # :23
    viewport_width = t(aspect_ratio * viewport_height)         # :24
    w = _normalize((lookfrom - lookat).astype(t))              # :26
    u = _normalize(np.cross(vup, w).astype(t))                 # :27
    v = np.cross(w, u).astype(t)                               # :28
    origin = lookfrom                                          # :30
    horizontal = (t(focus_dist * viewport_width) * u).astype(t)   # :31
    vertical = (t(focus_dist * viewport_height) * v).astype(t)    # :32
    llc = (origin - horizontal / t(2) - vertical / t(2) - focus_dist * w).astype(t)  # :33
    lens_radius = t(aperture / t(2))                           # :34
    return Camera(origin, llc, horizontal, vertical, u, v, w, lens_radius, t)


# ----------------------------------------------------------------------------- scenes.jl


def scene_2_spheres(*, elem_type=F32) -> HittableList:
    """src/scenes.jl:2-11"""
    t = elem_type
    return HittableList([
        Sphere(Vec3(0, 0, -1, t), t(0.5), Lambertian(Vec3(0.7, 0.3, 0.3, t))),
        Sphere(Vec3(0, -100.5, -1, t), t(100), Lambertian(Vec3(0.8, 0.8, 0.0, t))),
    ])


def scene_4_spheres(*, elem_type=F32) -> HittableList:
    """src/scenes.jl:16-23"""
    t = elem_type
    scene = scene_2_spheres(elem_type=t)
    scene.append(Sphere(Vec3(-1, 0, -1, t), t(0.5), Metal(Vec3(0.8, 0.8, 0.8, t), t(0.3))))
    scene.append(Sphere(Vec3(1, 0, -1, t), t(0.5), Metal(Vec3(0.8, 0.6, 0.2, t), t(0.8))))
    return scene


def scene_diel_spheres(left_radius=0.5, *, elem_type=F32) -> HittableList:
    """src/scenes.jl:25-39 (a negative left_radius makes the hollow-glass bubble)"""
    t = elem_type
    return HittableList([
        Sphere(Vec3(0, 0, -1, t), t(0.5), Lambertian(Vec3(0.1, 0.2, 0.5, t))),
        Sphere(Vec3(0, -100.5, -1, t), t(100), Lambertian(Vec3(0.8, 0.8, 0.0, t))),
        Sphere(Vec3(-1, 0, -1, t), t(left_radius), Dielectric(t(1.5))),
        Sphere(Vec3(1, 0, -1, t), t(0.5), Metal(Vec3(0.8, 0.6, 0.2, t), t(0))),
    ])


def scene_blue_red_spheres(*, elem_type=F32) -> HittableList:
    """src/scenes.jl:41-47"""
    t = elem_type
    r = math.cos(math.pi / 4)
    return HittableList([
        Sphere(Vec3(-r, 0, -1, t), t(r), Lambertian(Vec3(0, 0, 1, t))),
        Sphere(Vec3(r, 0, -1, t), t(r), Lambertian(Vec3(1, 0, 0, t))),
    ])


def _random_grid(spheres: HittableList, lo: int, hi: int, t) -> None:
    """the `for a in lo:hi, b in lo:hi` body of scene_random_spheres, src/scenes.jl:56-76"""
    for a in range(lo, hi + 1):
        for b in range(lo, hi + 1):
            choose_mat = trand(t)                                            # :57
            cx = t(t(a) + t(t(0.9) * trand(t)))                              # :58 (x drawn before z)
            cz = t(t(b) + t(t(0.9) * trand(t)))
            center = Vec3(cx, t(0.2), cz, t)
            dx, dy, dz = t(center[0] - t(4)), t(center[1] - t(0.2)), t(center[2] - t(0))
            if t(np.sqrt(t(t(dx * dx) + t(dy * dy)) + t(dz * dz))) < t(0.9):  # :61
                continue
            if choose_mat < t(0.8):                                          # :63-66 diffuse
                a3 = [trand(t) for _ in range(3)]
                b3 = [trand(t) for _ in range(3)]
                albedo = np.array([t(x * y) for x, y in zip(a3, b3)], dtype=t)
                spheres.append(Sphere(center, t(0.2), Lambertian(albedo)))
            elif choose_mat < t(0.95):                                       # :67-71 metal, fuzz in [0,5)
                albedo = np.array([random_between(t(0.5), t(1.0)) for _ in range(3)], dtype=t)
                fuzz = random_between(t(0.0), t(5.0))
                spheres.append(Sphere(center, t(0.2), Metal(albedo, fuzz)))
            else:                                                            # :72-74 glass
                spheres.append(Sphere(center, t(0.2), Dielectric(t(1.5))))


def scene_random_spheres(*, elem_type=F32, half_extent: int = 11) -> HittableList:
    """scene_random_spheres(; elem_type), src/scenes.jl:49-84.  Draws from the calling thread's TRNG in
    whatever state it is (scripts call reseed!() first, src/proto/proto.jl:198-199).
    `half_extent` generalises the -11:10 grid (reference value 11) for the synthetic large-N config:
    half_extent=158 gives the ~100k-sphere list of BASELINE.json configs[4]."""
    t = elem_type
    spheres = HittableList()
    spheres.append(Sphere(Vec3(0, -1000, -1, t), t(1000), Lambertian(Vec3(0.5, 0.5, 0.5, t))))  # :53-54
    _random_grid(spheres, -half_extent, half_extent - 1, t)                                      # :56
    spheres.append(Sphere(Vec3(0, 1, 0, t), t(1), Dielectric(t(1.5))))                           # :78
    spheres.append(Sphere(Vec3(-4, 1, 0, t), t(1), Lambertian(Vec3(0.4, 0.2, 0.1, t))))          # :79-80
    spheres.append(Sphere(Vec3(4, 1, 0, t), t(1), Metal(Vec3(0.7, 0.6, 0.5, t), t(0))))          # :81-82
    return spheres


def image_height(image_width: int) -> int:
    """image_width div (16//9), src/render.jl:11-12"""
    return (int(image_width) * 9) // 16


# canonical cameras of the reference's scripts/tests
def t_default_cam(elem_type=F32) -> Camera:
    """default_camera(SA{T}[0,0,0]), test/runtests.jl:190"""
    return default_camera((0, 0, 0), elem_type=elem_type)


def t_cam1(elem_type=F32) -> Camera:
    """t_cam1, src/proto/proto.jl:19"""
    return default_camera([13, 2, 3], [0, 0, 0], [0, 1, 0], 20, 16 / 9, 0.1, 10.0, elem_type=elem_type)


def t_cam2(elem_type=F32) -> Camera:
    """t_cam2, src/proto/proto.jl:21-22"""
    return default_camera([3, 3, 2], [0, 0, -1], [0, 1, 0], 20, 16 / 9, 2.0, math.sqrt(27.0), elem_type=elem_type)
