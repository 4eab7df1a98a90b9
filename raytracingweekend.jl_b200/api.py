"""render() and the Renderer context: the host side of the C-ABI boundary, mirroring
render(scene::HittableList, cam::Camera{T}, image_width=400, n_samples=1), src/render.jl:8-9.

The Julia shim (julia/RayTracingWeekendB200.jl) makes exactly the same calls with `ccall`.
No CPU fallback: every call needs librtw_b200.so and a CUDA device, and raises RtwError otherwise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import RtwError, rtw_camera, rtw_camera_f64, rtw_stats
from .host import Camera, F32, F64, Sphere, flatten_scene, image_height

DEFAULT_MAX_DEPTH = 16  # ray_color's default depth, src/ray_color.jl:14
DEFAULT_SEED = 1        # reseed!() semantics: the same image on every call, src/render.jl:21


_cam_cache = {}


def _camera_struct(cam: Camera) -> rtw_camera:
    hit = _cam_cache.get(id(cam))
    if hit is not None and hit[0] is cam:  # Camera is frozen: the struct of the same object never changes
        return hit[1]
    c = _camera_struct_uncached(cam)
    if len(_cam_cache) > 64:
        _cam_cache.clear()
    _cam_cache[id(cam)] = (cam, c)
    return c


def _camera_struct_uncached(cam: Camera) -> rtw_camera:
    if cam.elem_type != F32:
        raise RtwError(_lib.RTW_E_UNSUPPORTED, "this entry point takes a Camera{Float32} (render() also serves Camera{Float64})")
    c = rtw_camera()
    for name in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w"):
        arr = np.asarray(getattr(cam, name), dtype=F32)
        getattr(c, name)[:] = [float(x) for x in arr]
    c.lens_radius = float(cam.lens_radius)
    return c


def _camera_struct_f64(cam: Camera) -> rtw_camera_f64:
    hit = _cam_cache.get(("f64", id(cam)))
    if hit is not None and hit[0] is cam:
        return hit[1]
    c = _camera_struct_f64_uncached(cam)
    if len(_cam_cache) > 64:
        _cam_cache.clear()
    _cam_cache[("f64", id(cam))] = (cam, c)
    return c


def _camera_struct_f64_uncached(cam: Camera) -> rtw_camera_f64:
    c = rtw_camera_f64()
    for name in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w"):
        getattr(c, name)[:] = [float(x) for x in np.asarray(getattr(cam, name), dtype=F64)]
    c.lens_radius = float(cam.lens_radius)
    return c


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Renderer:
    """Owns an rtw_ctx (streams, device buffers).  `devices`: list of CUDA device ids (default [0])."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        devs = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devs))(*devs)
        status = self._lib.rtw_create(arr, len(devs), C.byref(self._ctx))
        if status != 0:
            self._ctx = C.c_void_p()
            _lib.check(None, status)
        self.devices = devs
        self.n_spheres = 0
        self.last_stats: Optional[dict] = None

    # -- life cycle
    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.rtw_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, status: int) -> None:
        _lib.check(self._ctx, status)

    def set_option(self, option: int, value: int) -> None:
        self._check(self._lib.rtw_set_option(self._ctx, option, int(value)))
        if option == _lib.RTW_OPT_GATHER:
            self._gather = int(value)

    def gather_name(self) -> str:
        """how a multi-device context collects the row tiles (RTW_OPT_GATHER)"""
        if len(self.devices) == 1:
            return "none (1 device)"
        return "NCCL send/recv" if getattr(self, "_gather", 0) == _lib.RTW_GATHER_NCCL else "peer copies"

    # -- scene
    def set_scene(self, scene) -> None:
        """scene: a HittableList (list of Sphere) or an already flattened (geom4, mat4, kind) triple."""
        geom, mat, kind = _as_flat(scene)
        self._check(self._lib.rtw_set_scene(self._ctx, _fp(geom), _fp(mat),
                                            kind.ctypes.data_as(C.POINTER(C.c_uint32)), len(kind)))
        self.n_spheres = len(kind)

    def generate_random_spheres(self, half_extent: int = 11, *, install: bool = True, rng=None):
        """scene_random_spheres(; elem_type=Float32), src/scenes.jl:49-84, built on the device (no host loop): the same
        list, bit for bit, as host.scene_random_spheres(half_extent=...) drawing from `rng` (default: the calling
        thread's TRNG[0]), whose state advances exactly as the host loop would advance it.  Returns the flattened
        (geom4, mat4, kind); with install=True the list also becomes the scene of this context."""
        from .host import TRNG
        g = rng if rng is not None else TRNG[0]
        state = (C.c_uint64 * 2)(g.x, g.y)
        cap = 4 * int(half_extent) * int(half_extent) + 4
        geom = np.zeros((cap, 4), dtype=F32)
        mat = np.zeros((cap, 4), dtype=F32)
        kind = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32()
        self._check(self._lib.rtw_scene_random_spheres(self._ctx, state, int(half_extent), 1 if install else 0, _fp(geom),
                                                       _fp(mat), kind.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(n)))
        g.x, g.y = int(state[0]), int(state[1])
        if install:
            self.n_spheres = n.value
        return geom[:n.value].copy(), mat[:n.value].copy(), kind[:n.value].copy()

    # -- the hot path, host buffers (what the Julia `render` binds to)
    def render(self, cam: Camera, image_width: int = 400, n_samples: int = 1, *, max_depth: int = DEFAULT_MAX_DEPTH,
               seed: int = DEFAULT_SEED, scene=None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Returns the image as an (H, W, 3) float32 array view (row 0 = top), backed by a buffer in the
        memory layout of Julia's Matrix{RGB{Float32}}(H, W).  With `scene`, the scene upload is part of the call
        (rtw_render_scene)."""
        W = int(image_width)
        H = image_height(W)
        if cam.elem_type == F64:
            return self._render_f64(cam, W, H, int(n_samples), int(max_depth), int(seed), scene, out)
        if out is None:
            out = np.empty((W, H, 3), dtype=F32)  # C-order (W,H,3) == column-major H x W of RGB
        elif out.shape != (W, H, 3) or out.dtype != F32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape (W, H, 3)")
        cs = _camera_struct(cam)
        st = rtw_stats()
        if scene is not None:
            geom, mat, kind = _as_flat(scene)
            status = self._lib.rtw_render_scene(self._ctx, _fp(geom), _fp(mat),
                                                kind.ctypes.data_as(C.POINTER(C.c_uint32)), len(kind), C.byref(cs), W,
                                                int(n_samples), int(max_depth), int(seed), _fp(out), C.byref(st))
            if status == 0:
                self.n_spheres = len(kind)
        else:
            status = self._lib.rtw_render(self._ctx, C.byref(cs), W, int(n_samples), int(max_depth), int(seed),
                                          _fp(out), C.byref(st))
        self._check(status)
        self.last_stats = st.as_dict()
        return out.transpose(1, 0, 2)

    # -- Camera{Float64}: the Float64 instantiation (scene arrays and image are doubles)
    def set_scene_f64(self, scene) -> None:
        geom, mat, kind = _as_flat(scene, F64)
        self._check(self._lib.rtw_set_scene_f64(self._ctx, _dp(geom), _dp(mat), kind.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                len(kind)))

    def _render_f64(self, cam, W, H, n_samples, max_depth, seed, scene, out):
        if out is None:
            out = np.empty((W, H, 3), dtype=F64)
        elif out.shape != (W, H, 3) or out.dtype != F64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (W, H, 3)")
        cs = _camera_struct_f64(cam)
        st = rtw_stats()
        if scene is not None:
            geom, mat, kind = _as_flat(scene, F64)
            status = self._lib.rtw_render_scene_f64(self._ctx, _dp(geom), _dp(mat), kind.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                    len(kind), C.byref(cs), W, n_samples, max_depth, seed, _dp(out), C.byref(st))
        else:
            status = self._lib.rtw_render_f64(self._ctx, C.byref(cs), W, n_samples, max_depth, seed, _dp(out), C.byref(st))
        self._check(status)
        self.last_stats = st.as_dict()
        return out.transpose(1, 0, 2)

    # -- progressive rendering: render() split into passes over the samples (bit-identical to one render)
    def accumulate(self, cam: Camera, image_width: int, sample_first: int, sample_count: int, n_samples_total: int, *,
                   max_depth: int = DEFAULT_MAX_DEPTH, seed: int = DEFAULT_SEED) -> dict:
        """Adds samples [sample_first, sample_first + sample_count) of every pixel to the accumulators held by the
        context; sample_first == 0 starts a new image.  Returns the stats of the pass."""
        cs = _camera_struct(cam)
        st = rtw_stats()
        self._check(self._lib.rtw_accumulate(self._ctx, C.byref(cs), int(image_width), int(sample_first), int(sample_count),
                                             int(n_samples_total), int(max_depth), int(seed), C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    def progress(self):
        """(image_width, samples_done, samples_total) of the progressive image; zeros when there is none."""
        w, d, t = C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.rtw_progress(self._ctx, C.byref(w), C.byref(d), C.byref(t)))
        return w.value, d.value, t.value

    def resolve(self) -> np.ndarray:
        """sqrt(sum / samples so far) of the progressive image: (H, W, 3) float32 view, as render()."""
        W, done, _ = self.progress()
        if done < 1:
            raise RtwError(_lib.RTW_E_INVALID_ARG, "no progressive image: call accumulate() first")
        out = np.empty((W, image_height(W), 3), dtype=F32)
        self._check(self._lib.rtw_resolve(self._ctx, _fp(out)))
        return out.transpose(1, 0, 2)

    def resolve_rgb8(self) -> np.ndarray:
        """The progressive image as (H, W, 3) uint8, row-major (clamp01nan + N0f8 rounding, as Images.jl saves)."""
        W, done, _ = self.progress()
        if done < 1:
            raise RtwError(_lib.RTW_E_INVALID_ARG, "no progressive image: call accumulate() first")
        out = np.empty((image_height(W), W, 3), dtype=np.uint8)
        self._check(self._lib.rtw_resolve_rgb8(self._ctx, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def accumulator_read(self) -> np.ndarray:
        """Checkpoint: the raw fixed-point sums, (H, W, 4) int64 (r, g, b, unused)."""
        W, _, _ = self.progress()
        out = np.empty((image_height(W), W, 4), dtype=np.int64)
        self._check(self._lib.rtw_accumulator_read(self._ctx, out.ctypes.data_as(C.POINTER(C.c_int64)), out.size))
        return out

    def accumulator_write(self, acc: np.ndarray, image_width: int, samples_done: int, samples_total: int) -> None:
        """Resume: install saved sums so that accumulate() continues at sample `samples_done`."""
        acc = np.ascontiguousarray(acc, dtype=np.int64)
        self._check(self._lib.rtw_accumulator_write(self._ctx, acc.ctypes.data_as(C.POINTER(C.c_int64)), acc.size,
                                                    int(image_width), int(samples_done), int(samples_total)))

    def checkpoint_save(self, path) -> None:
        """The progressive image as a self-describing file (sums + what they were traced with + CRC-32)."""
        self._check(self._lib.rtw_checkpoint_save(self._ctx, str(path).encode()))

    def checkpoint_load(self, path) -> None:
        """Resume from a checkpoint file; the scene it was rendered from must be set already."""
        self._check(self._lib.rtw_checkpoint_load(self._ctx, str(path).encode()))

    # -- device-resident variants (plain device pointers; used with torch tensors as plumbing)
    def render_rows_device(self, cam: Camera, image_width: int, n_samples: int, d_tile_ptr: int, *, max_depth: int =
                           DEFAULT_MAX_DEPTH, seed: int = DEFAULT_SEED, row_start: int = 0, row_stride: int = 1,
                           column_major: bool = False, stream: int = 0, device_slot: int = 0) -> None:
        cs = _camera_struct(cam)
        self._check(self._lib.rtw_render_rows_device(self._ctx, device_slot, C.byref(cs), int(image_width),
                                                     int(n_samples), int(max_depth), int(seed), int(row_start),
                                                     int(row_stride), 1 if column_major else 0,
                                                     C.c_void_p(d_tile_ptr), C.c_void_p(stream)))

    def stats(self, device_slot: int = 0) -> dict:
        st = rtw_stats()
        self._check(self._lib.rtw_last_stats(self._ctx, device_slot, C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    def assemble_tiles_device(self, d_tiles_ptr: int, n_tiles: int, image_width: int, d_out_ptr: int, *,
                              stream: int = 0, device_slot: int = 0) -> None:
        self._check(self._lib.rtw_assemble_tiles_device(self._ctx, device_slot, C.c_void_p(d_tiles_ptr), int(n_tiles),
                                                        int(image_width), C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def measure_fp32_peak(self, variant: int = 0, device_slot: int = 0):
        rate = C.c_double()
        ms = C.c_float()
        self._check(self._lib.rtw_measure_fp32_peak(self._ctx, device_slot, int(variant), C.byref(rate), C.byref(ms)))
        return rate.value, ms.value


def _as_flat(scene, elem_type=F32):
    if isinstance(scene, tuple) and len(scene) == 3:
        geom, mat, kind = scene
    else:
        geom, mat, kind = flatten_scene(scene, elem_type)
    geom = np.ascontiguousarray(geom, dtype=elem_type).reshape(-1, 4)
    mat = np.ascontiguousarray(mat, dtype=elem_type).reshape(-1, 4)
    kind = np.ascontiguousarray(kind, dtype=np.uint32).reshape(-1)
    if not (len(geom) == len(mat) == len(kind)):
        raise ValueError("geom4, mat4 and kind must have the same length")
    return geom, mat, kind


# -- image and scene files (host side of the library; no device needed)
def write_ppm(path, rgb8: np.ndarray) -> None:
    """Binary PPM (P6) of an (H, W, 3) uint8 image."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = rgb8.shape
    _lib.check(None, _lib.load().rtw_write_ppm(str(path).encode(), rgb8.ctypes.data_as(C.POINTER(C.c_uint8)), w, h))


def write_png(path, rgb8: np.ndarray) -> None:
    """PNG (8-bit RGB) of an (H, W, 3) uint8 image."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = rgb8.shape
    _lib.check(None, _lib.load().rtw_write_png(str(path).encode(), rgb8.ctypes.data_as(C.POINTER(C.c_uint8)), w, h))


def scene_save(path, scene) -> None:
    """Writes a HittableList (or a flattened triple) as a .rtwscene file."""
    geom, mat, kind = _as_flat(scene)
    _lib.check(None, _lib.load().rtw_scene_save(str(path).encode(), _fp(geom), _fp(mat),
                                                kind.ctypes.data_as(C.POINTER(C.c_uint32)), len(kind)))


def scene_load(path):
    """Reads a .rtwscene file; returns the flattened (geom4, mat4, kind) triple."""
    lib = _lib.load()
    n = C.c_uint32()
    _lib.check(None, lib.rtw_scene_load(str(path).encode(), None, None, None, 0, C.byref(n)))
    geom = np.zeros((n.value, 4), dtype=F32)
    mat = np.zeros((n.value, 4), dtype=F32)
    kind = np.zeros(n.value, dtype=np.uint32)
    if n.value:
        _lib.check(None, lib.rtw_scene_load(str(path).encode(), _fp(geom), _fp(mat),
                                            kind.ctypes.data_as(C.POINTER(C.c_uint32)), n.value, C.byref(n)))
    return geom, mat, kind


def has_variants() -> bool:
    """True when librtw_b200.so was built with RTW_BUILD_VARIANTS=1 (the kernel families kept as measured comparisons)."""
    return bool(_lib.load().rtw_has_variants())


_default_renderer: Optional[Renderer] = None


def render(scene, cam: Camera, image_width: int = 400, n_samples: int = 1, *, max_depth: int = DEFAULT_MAX_DEPTH,
           seed: int = DEFAULT_SEED, devices: Optional[Sequence[int]] = None) -> np.ndarray:
    """render(scene, cam, image_width=400, n_samples=1), src/render.jl:8-44, on the GPU(s).

    Same positional arguments and defaults as the reference.  Keywords the reference hard-codes:
    max_depth (ray_color's depth=16, src/ray_color.jl:14) and seed (reseed!() => constant, src/render.jl:21).
    Returns an (H, W, 3) float32 array: img[i, j] is the reference's img[i+1, j+1] (gamma-2, unclamped)."""
    global _default_renderer
    if devices is not None:
        with Renderer(devices) as r:
            return np.array(r.render(cam, image_width, n_samples, max_depth=max_depth, seed=seed, scene=scene))
    if _default_renderer is None:
        _default_renderer = Renderer()
    return _default_renderer.render(cam, image_width, n_samples, max_depth=max_depth, seed=seed, scene=scene)
