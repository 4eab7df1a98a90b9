"""ctypes binding of librtw_b200.so (the C-ABI declared in include/rtw_b200.h).

There is no CPU fallback: if the CUDA library is missing, or no CUDA device is present, every
compute entry point raises.  Nothing in this module (or package) imports or calls oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# RTW_B200_LIB selects another build of the same library, e.g. csrc/librtw_b200_variants.so (RTW_BUILD_VARIANTS=1)
LIB_PATH = Path(os.environ["RTW_B200_LIB"]) if os.environ.get("RTW_B200_LIB") else _HERE / "csrc" / "librtw_b200.so"

# every symbol include/rtw_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = (
    "rtw_abi_version",
    "rtw_has_variants",
    "rtw_device_count",
    "rtw_image_height",
    "rtw_create",
    "rtw_destroy",
    "rtw_last_error",
    "rtw_set_option",
    "rtw_set_scene",
    "rtw_scene_random_spheres",
    "rtw_render",
    "rtw_render_scene",
    "rtw_render_rows_device",
    "rtw_last_stats",
    "rtw_assemble_tiles_device",
    "rtw_measure_fp32_peak",
    "rtw_set_scene_f64",
    "rtw_render_f64",
    "rtw_render_scene_f64",
    "rtw_accumulate",
    "rtw_resolve",
    "rtw_resolve_rgb8",
    "rtw_progress",
    "rtw_accumulator_read",
    "rtw_accumulator_write",
    "rtw_checkpoint_save",
    "rtw_checkpoint_load",
    "rtw_write_ppm",
    "rtw_write_png",
    "rtw_scene_save",
    "rtw_scene_load",
)

RTW_ABI_VERSION = 3

RTW_OK = 0
RTW_E_INVALID_ARG = -1
RTW_E_NO_DEVICE = -2
RTW_E_NO_SCENE = -3
RTW_E_UNSUPPORTED = -4
RTW_E_INTERNAL = -5
RTW_E_IO = -6
RTW_E_FORMAT = -7

RTW_OPT_MODE = 1
RTW_OPT_STRIP = 2
RTW_OPT_BLOCKS_PER_SM = 3
RTW_OPT_COLLECT_TIMING = 4
RTW_OPT_RAYS_PER_LANE = 5
RTW_OPT_SWEEP = 6
RTW_OPT_COOP = 7
RTW_OPT_TAIL = 8

RTW_OPT_WALK = 9
RTW_OPT_GATHER = 10
RTW_OPT_SMALL_RENDER = 11
RTW_GATHER_PEER = 0
RTW_GATHER_NCCL = 1
RTW_WALK_DEFAULT = 0
RTW_WALK_SLOTS = 1
RTW_WALK_OWN_RAY = 2
RTW_TAIL_DEFAULT = 0
RTW_TAIL_SPLIT = 1
RTW_TAIL_UNIFIED = 2

RTW_SWEEP_DEFAULT = 0
RTW_SWEEP_BRANCH = 1
RTW_SWEEP_MASK = 2
RTW_SWEEP_PACKED = 3

RTW_MODE_FUSED = 0
RTW_MODE_WAVEFRONT = 1
RTW_MODE_CTA_WAVEFRONT = 2
RTW_MODE_GRID = 3


class RtwError(RuntimeError):
    """Raised for every non-zero status of the C-ABI (mirrors the Julia shim's error())."""

    def __init__(self, code: int, message: str):
        super().__init__(f"rtw_b200 error {code}: {message}")
        self.code = code


class rtw_camera(C.Structure):
    """Camera{Float32}, src/camera.jl:1-10 -- 22 x f32, same field order."""

    _fields_ = [
        ("origin", C.c_float * 3),
        ("lower_left_corner", C.c_float * 3),
        ("horizontal", C.c_float * 3),
        ("vertical", C.c_float * 3),
        ("u", C.c_float * 3),
        ("v", C.c_float * 3),
        ("w", C.c_float * 3),
        ("lens_radius", C.c_float),
    ]


class rtw_camera_f64(C.Structure):
    """Camera{Float64}: the same 22 fields as doubles."""

    _fields_ = [
        ("origin", C.c_double * 3),
        ("lower_left_corner", C.c_double * 3),
        ("horizontal", C.c_double * 3),
        ("vertical", C.c_double * 3),
        ("u", C.c_double * 3),
        ("v", C.c_double * 3),
        ("w", C.c_double * 3),
        ("lens_radius", C.c_double),
    ]


class rtw_stats(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64),
        ("ray_segments", C.c_uint64),
        ("sphere_tests", C.c_uint64),
        ("n_spheres", C.c_uint32),
        ("image_width", C.c_int32),
        ("image_height", C.c_int32),
        ("rows_rendered", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("ms_total", C.c_float),
        ("ms_trace", C.c_float),
        ("ms_resolve", C.c_float),
        ("ms_h2d", C.c_float),
        ("ms_d2h", C.c_float),
        ("n_devices", C.c_int32),
        ("reserved0", C.c_int32),
        ("grid_fallback_rays", C.c_uint64),
        ("grid_loose_cells", C.c_uint64),
        ("grid_cells", C.c_uint64),
        ("grid_tests", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RtwError(
            RTW_E_INTERNAL,
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or raytracingweekend.jl_b200/csrc/build.sh -- there is no CPU fallback for the hot path",
        )
    lib = C.CDLL(os.fspath(LIB_PATH))
    vp, i32, u32, u64, i64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_int64
    fp = C.POINTER(C.c_float)
    u32p = C.POINTER(C.c_uint32)
    lib.rtw_abi_version.restype = i32
    lib.rtw_abi_version.argtypes = []
    lib.rtw_has_variants.restype = i32
    lib.rtw_has_variants.argtypes = []
    lib.rtw_device_count.restype = i32
    lib.rtw_device_count.argtypes = [C.POINTER(i32)]
    lib.rtw_image_height.restype = i32
    lib.rtw_image_height.argtypes = [i32]
    lib.rtw_create.restype = i32
    lib.rtw_create.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
    lib.rtw_destroy.restype = i32
    lib.rtw_destroy.argtypes = [vp]
    lib.rtw_last_error.restype = C.c_char_p
    lib.rtw_last_error.argtypes = [vp]
    lib.rtw_set_option.restype = i32
    lib.rtw_set_option.argtypes = [vp, i32, i64]
    lib.rtw_set_scene.restype = i32
    lib.rtw_set_scene.argtypes = [vp, fp, fp, u32p, u32]
    lib.rtw_scene_random_spheres.restype = i32
    lib.rtw_scene_random_spheres.argtypes = [vp, C.POINTER(u64), i32, i32, fp, fp, u32p, u32, u32p]
    lib.rtw_render.restype = i32
    lib.rtw_render.argtypes = [vp, C.POINTER(rtw_camera), i32, i32, i32, u64, fp, C.POINTER(rtw_stats)]
    lib.rtw_render_scene.restype = i32
    lib.rtw_render_scene.argtypes = [vp, fp, fp, u32p, u32, C.POINTER(rtw_camera), i32, i32, i32, u64, fp,
                                     C.POINTER(rtw_stats)]
    lib.rtw_render_rows_device.restype = i32
    lib.rtw_render_rows_device.argtypes = [vp, i32, C.POINTER(rtw_camera), i32, i32, i32, u64, i32, i32, i32, vp, vp]
    lib.rtw_last_stats.restype = i32
    lib.rtw_last_stats.argtypes = [vp, i32, C.POINTER(rtw_stats)]
    lib.rtw_assemble_tiles_device.restype = i32
    lib.rtw_assemble_tiles_device.argtypes = [vp, i32, vp, i32, i32, vp, vp]
    lib.rtw_measure_fp32_peak.restype = i32
    lib.rtw_measure_fp32_peak.argtypes = [vp, i32, i32, C.POINTER(C.c_double), fp]
    dp = C.POINTER(C.c_double)
    lib.rtw_set_scene_f64.restype = i32
    lib.rtw_set_scene_f64.argtypes = [vp, dp, dp, u32p, u32]
    lib.rtw_render_f64.restype = i32
    lib.rtw_render_f64.argtypes = [vp, C.POINTER(rtw_camera_f64), i32, i32, i32, u64, dp, C.POINTER(rtw_stats)]
    lib.rtw_render_scene_f64.restype = i32
    lib.rtw_render_scene_f64.argtypes = [vp, dp, dp, u32p, u32, C.POINTER(rtw_camera_f64), i32, i32, i32, u64, dp,
                                         C.POINTER(rtw_stats)]
    u8p, i64p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(i32)
    lib.rtw_accumulate.restype = i32
    lib.rtw_accumulate.argtypes = [vp, C.POINTER(rtw_camera), i32, i32, i32, i32, i32, u64, C.POINTER(rtw_stats)]
    lib.rtw_resolve.restype = i32
    lib.rtw_resolve.argtypes = [vp, fp]
    lib.rtw_resolve_rgb8.restype = i32
    lib.rtw_resolve_rgb8.argtypes = [vp, u8p]
    lib.rtw_progress.restype = i32
    lib.rtw_progress.argtypes = [vp, i32p, i32p, i32p]
    lib.rtw_accumulator_read.restype = i32
    lib.rtw_accumulator_read.argtypes = [vp, i64p, u64]
    lib.rtw_accumulator_write.restype = i32
    lib.rtw_accumulator_write.argtypes = [vp, i64p, u64, i32, i32, i32]
    lib.rtw_checkpoint_save.restype = i32
    lib.rtw_checkpoint_save.argtypes = [vp, C.c_char_p]
    lib.rtw_checkpoint_load.restype = i32
    lib.rtw_checkpoint_load.argtypes = [vp, C.c_char_p]
    lib.rtw_write_ppm.restype = i32
    lib.rtw_write_ppm.argtypes = [C.c_char_p, u8p, i32, i32]
    lib.rtw_write_png.restype = i32
    lib.rtw_write_png.argtypes = [C.c_char_p, u8p, i32, i32]
    lib.rtw_scene_save.restype = i32
    lib.rtw_scene_save.argtypes = [C.c_char_p, fp, fp, u32p, u32]
    lib.rtw_scene_load.restype = i32
    lib.rtw_scene_load.argtypes = [C.c_char_p, fp, fp, u32p, u32, u32p]
    _lib = lib
    return lib


def check(ctx, status: int) -> None:
    if status == RTW_OK:
        return
    lib = load()
    msg = ""
    if ctx:
        raw = lib.rtw_last_error(ctx)
        msg = raw.decode("utf-8", "replace") if raw else ""
    if not msg:
        msg = {
            RTW_E_INVALID_ARG: "invalid argument",
            RTW_E_NO_DEVICE: "no CUDA device visible (the hot path has no CPU fallback)",
            RTW_E_NO_SCENE: "no scene set",
            RTW_E_UNSUPPORTED: "unsupported",
            RTW_E_IO: "file could not be opened / read / written",
            RTW_E_FORMAT: "not a .rtwscene file, or its checksum does not match",
            RTW_E_INTERNAL: "internal error",
        }.get(status, "CUDA error" if status > 0 else "error")
    raise RtwError(status, msg)
