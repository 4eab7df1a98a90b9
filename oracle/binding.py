"""ctypes binding of the CPU oracle (oracle/librtw_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (raytracingweekend.jl_b200) never imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "librtw_oracle.so"
LIB_PATH_DEFAULT = LIB_PATH

RNG_PHILOX = 0
RNG_XOROSHIRO = 1


class rtwo_camera_f32(C.Structure):
    _fields_ = [(n, C.c_float * 3) for n in
                ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w")] + [("lens_radius", C.c_float)]


class rtwo_camera_f64(C.Structure):
    _fields_ = [(n, C.c_double * 3) for n in
                ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w")] + [("lens_radius", C.c_double)]


class rtwo_stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("ray_segments", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("seconds", C.c_double), ("threads", C.c_int)]


_lib = None


def build() -> None:
    subprocess.run(["make", "-C", os.fspath(_HERE), "librtw_oracle.so"], check=True, capture_output=True)


def use_library(path) -> None:
    """Switch the binding to another build of the same source (bench.py: the -O3 -march=native timing build)."""
    global _lib, LIB_PATH
    LIB_PATH = Path(path)
    _lib = None
    load()


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    try:
        lib = C.CDLL(os.fspath(LIB_PATH))
    except OSError:
        build()  # e.g. built on a different host image
        lib = C.CDLL(os.fspath(LIB_PATH))
    d3, f3 = C.POINTER(C.c_double), C.POINTER(C.c_float)
    u32p = C.POINTER(C.c_uint32)
    lib.rtwo_image_height.restype = C.c_int
    lib.rtwo_image_height.argtypes = [C.c_int]
    lib.rtwo_reflect_f64.argtypes = [d3, d3, d3]
    lib.rtwo_reflect_f32.argtypes = [f3, f3, f3]
    lib.rtwo_refract_f64.argtypes = [d3, d3, C.c_double, d3]
    lib.rtwo_refract_f32.argtypes = [f3, f3, C.c_float, f3]
    lib.rtwo_reflectance_f64.restype = C.c_double
    lib.rtwo_reflectance_f64.argtypes = [C.c_double, C.c_double]
    lib.rtwo_reflectance_f32.restype = C.c_float
    lib.rtwo_reflectance_f32.argtypes = [C.c_float, C.c_float]
    lib.rtwo_near_zero_f64.restype = C.c_int
    lib.rtwo_near_zero_f64.argtypes = [d3]
    lib.rtwo_near_zero_f32.restype = C.c_int
    lib.rtwo_near_zero_f32.argtypes = [f3]
    lib.rtwo_hit_sphere_f64.restype = C.c_int
    lib.rtwo_hit_sphere_f64.argtypes = [d3, C.c_double, d3, d3, C.c_double, C.c_double, d3, d3, d3, C.POINTER(C.c_int)]
    lib.rtwo_hit_sphere_f32.restype = C.c_int
    lib.rtwo_hit_sphere_f32.argtypes = [f3, C.c_float, f3, f3, C.c_float, C.c_float, f3, f3, f3, C.POINTER(C.c_int)]
    lib.rtwo_skycolor_f32.argtypes = [f3, d3]
    lib.rtwo_skycolor_f64.argtypes = [d3, d3]
    lib.rtwo_philox4x32_10.argtypes = [u32p, u32p, u32p]
    lib.rtwo_philox4x32_7.argtypes = [u32p, u32p, u32p]
    lib.rtwo_path_stream_f32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, f3]
    lib.rtwo_xoroshiro_u64.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    lib.rtwo_xoroshiro_f32.argtypes = [C.c_uint64, C.c_int, f3]
    lib.rtwo_path_f32.argtypes = [f3, f3, u32p, C.c_uint32, C.POINTER(rtwo_camera_f32), C.c_int, C.c_int, C.c_uint64,
                                  C.c_int, C.c_int, C.c_int, d3, u32p]
    lib.rtwo_render_f32.restype = C.c_int
    lib.rtwo_render_f32.argtypes = [f3, f3, u32p, C.c_uint32, C.POINTER(rtwo_camera_f32), C.c_int, C.c_int, C.c_int,
                                    C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, f3, d3, C.POINTER(rtwo_stats)]
    lib.rtwo_render_f64.restype = C.c_int
    lib.rtwo_render_f64.argtypes = [d3, d3, u32p, C.c_uint32, C.POINTER(rtwo_camera_f64), C.c_int, C.c_int, C.c_int,
                                    C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, d3, d3, C.POINTER(rtwo_stats)]
    _lib = lib
    return lib


def _vec(a, ctype):
    arr = (ctype * 3)(*[float(x) for x in a])
    return arr


def reflect(v, n, dtype=np.float64):
    lib = load()
    ct = C.c_double if dtype == np.float64 else C.c_float
    out = (ct * 3)()
    (lib.rtwo_reflect_f64 if dtype == np.float64 else lib.rtwo_reflect_f32)(_vec(v, ct), _vec(n, ct), out)
    return np.array(list(out), dtype=dtype)


def refract(d, n, ratio, dtype=np.float64):
    lib = load()
    ct = C.c_double if dtype == np.float64 else C.c_float
    out = (ct * 3)()
    (lib.rtwo_refract_f64 if dtype == np.float64 else lib.rtwo_refract_f32)(_vec(d, ct), _vec(n, ct), ratio, out)
    return np.array(list(out), dtype=dtype)


def reflectance(cos_t, ratio, dtype=np.float64):
    lib = load()
    return (lib.rtwo_reflectance_f64 if dtype == np.float64 else lib.rtwo_reflectance_f32)(cos_t, ratio)


def near_zero(v, dtype=np.float64) -> bool:
    lib = load()
    ct = C.c_double if dtype == np.float64 else C.c_float
    return bool((lib.rtwo_near_zero_f64 if dtype == np.float64 else lib.rtwo_near_zero_f32)(_vec(v, ct)))


def hit_sphere(center, radius, o, d, tmin, tmax, dtype=np.float64):
    """returns None on a miss, else (t, p, n, front_face)"""
    lib = load()
    ct = C.c_double if dtype == np.float64 else C.c_float
    t = ct()
    p, n = (ct * 3)(), (ct * 3)()
    ff = C.c_int()
    fn = lib.rtwo_hit_sphere_f64 if dtype == np.float64 else lib.rtwo_hit_sphere_f32
    ok = fn(_vec(center, ct), radius, _vec(o, ct), _vec(d, ct), tmin, tmax, C.cast(C.byref(t), C.POINTER(ct)), p, n,
            C.byref(ff))
    if not ok:
        return None
    return t.value, np.array(list(p), dtype=dtype), np.array(list(n), dtype=dtype), bool(ff.value)


def skycolor(d, dtype=np.float32):
    lib = load()
    ct = C.c_double if dtype == np.float64 else C.c_float
    out = (C.c_double * 3)()
    (lib.rtwo_skycolor_f64 if dtype == np.float64 else lib.rtwo_skycolor_f32)(_vec(d, ct), out)
    return np.array(list(out))


def philox4x32_10(ctr, key):
    lib = load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib.rtwo_philox4x32_10(c, k, out)
    return [int(x) for x in out]


def philox4x32_7(ctr, key):
    """one block of the PRODUCTION stream (Philox4x32-7, RTWO_PHILOX_ROUNDS)"""
    lib = load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib.rtwo_philox4x32_7(c, k, out)
    return [int(x) for x in out]


def path_stream(seed, pixel, sample, event, n):
    """draws 0..n-1 of `event` of path (pixel, sample) in the production (addressed Philox) stream"""
    lib = load()
    out = np.zeros(n, dtype=np.float32)
    lib.rtwo_path_stream_f32(seed, pixel, sample, event, n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def xoroshiro_u64(seed, n):
    lib = load()
    out = np.zeros(n, dtype=np.uint64)
    lib.rtwo_xoroshiro_u64(seed, n, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def xoroshiro_f32(seed, n):
    lib = load()
    out = np.zeros(n, dtype=np.float32)
    lib.rtwo_xoroshiro_f32(seed, n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def _cam_struct(cam_array, f64=False):
    a = np.asarray(cam_array, dtype=np.float64 if f64 else np.float32).reshape(22)
    c = rtwo_camera_f64() if f64 else rtwo_camera_f32()
    names = ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w")
    for i, nme in enumerate(names):
        getattr(c, nme)[:] = [float(x) for x in a[3 * i:3 * i + 3]]
    c.lens_radius = float(a[21])
    return c


def render(geom4, mat4, kind, cam_array, image_width, n_samples, *, max_depth=16, seed=1, rng_mode=RNG_PHILOX,
           n_threads=0, row_start=0, row_stride=1, f64=False, want_linear=False):
    """Oracle render.  Returns (img[H,W,3], linear[H,W,3] or None, stats dict).  img is post-gamma, dtype T."""
    lib = load()
    ft = np.float64 if f64 else np.float32
    cft = C.c_double if f64 else C.c_float
    geom = np.ascontiguousarray(geom4, dtype=ft).reshape(-1, 4)
    mat = np.ascontiguousarray(mat4, dtype=ft).reshape(-1, 4)
    knd = np.ascontiguousarray(kind, dtype=np.uint32).reshape(-1)
    W = int(image_width)
    H = lib.rtwo_image_height(W)
    out = np.zeros((W, H, 3), dtype=ft)  # column-major H x W x RGB
    lin = np.zeros((W, H, 3), dtype=np.float64) if want_linear else None
    st = rtwo_stats()
    cam = _cam_struct(cam_array, f64)
    fn = lib.rtwo_render_f64 if f64 else lib.rtwo_render_f32
    rc = fn(geom.ctypes.data_as(C.POINTER(cft)), mat.ctypes.data_as(C.POINTER(cft)),
            knd.ctypes.data_as(C.POINTER(C.c_uint32)), len(knd), C.byref(cam), W, int(n_samples), int(max_depth),
            int(seed), int(rng_mode), int(n_threads), int(row_start), int(row_stride),
            out.ctypes.data_as(C.POINTER(cft)),
            lin.ctypes.data_as(C.POINTER(C.c_double)) if lin is not None else None, C.byref(st))
    if rc != 0:
        raise ValueError(f"oracle render failed with status {rc}")
    stats = {"paths": st.paths, "ray_segments": st.ray_segments, "sphere_tests": st.sphere_tests,
             "seconds": st.seconds, "threads": st.threads}
    return out.transpose(1, 0, 2), (lin.transpose(1, 0, 2) if lin is not None else None), stats


def path_trace(geom4, mat4, kind, cam_array, image_width, i0, j0, s0, *, max_depth=16, seed=1):
    """Debugging aid: the segments of one path -- rows of (origin xyz, direction xyz, closest sphere or -1, t)."""
    lib = load()
    geom = np.ascontiguousarray(geom4, dtype=np.float32).reshape(-1, 4)
    mat = np.ascontiguousarray(mat4, dtype=np.float32).reshape(-1, 4)
    knd = np.ascontiguousarray(kind, dtype=np.uint32).reshape(-1)
    cam = _cam_struct(cam_array)
    rgb = (C.c_double * 3)()
    trace = np.zeros((max_depth + 1, 8), dtype=np.float64)
    n = C.c_int()
    fp = C.POINTER(C.c_float)
    lib.rtwo_path_trace_f32.argtypes = [fp, fp, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(rtwo_camera_f32), C.c_int, C.c_int,
                                        C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.c_int, C.POINTER(C.c_int)]
    lib.rtwo_path_trace_f32(geom.ctypes.data_as(fp), mat.ctypes.data_as(fp), knd.ctypes.data_as(C.POINTER(C.c_uint32)),
                            len(knd), C.byref(cam), int(image_width), int(max_depth), int(seed), int(i0), int(j0), int(s0),
                            rgb, trace.ctypes.data_as(C.POINTER(C.c_double)), len(trace), C.byref(n))
    return np.array(list(rgb)), trace[:n.value]


def path(geom4, mat4, kind, cam_array, image_width, i0, j0, s0, *, max_depth=16, seed=1):
    lib = load()
    geom = np.ascontiguousarray(geom4, dtype=np.float32).reshape(-1, 4)
    mat = np.ascontiguousarray(mat4, dtype=np.float32).reshape(-1, 4)
    knd = np.ascontiguousarray(kind, dtype=np.uint32).reshape(-1)
    cam = _cam_struct(cam_array)
    rgb = (C.c_double * 3)()
    seg = C.c_uint32()
    fp = C.POINTER(C.c_float)
    lib.rtwo_path_f32(geom.ctypes.data_as(fp), mat.ctypes.data_as(fp), knd.ctypes.data_as(C.POINTER(C.c_uint32)),
                      len(knd), C.byref(cam), int(image_width), int(max_depth), int(seed), int(i0), int(j0), int(s0),
                      rgb, C.byref(seg))
    return np.array(list(rgb)), seg.value
