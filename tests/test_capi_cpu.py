"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, matches the ctypes/Julia struct layouts, and FAILS LOUDLY without a GPU (no CPU fallback).
No compute is attempted here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "rtw_b200.h"


def _header_prototypes():
    text = HEADER.read_text()
    return re.findall(r"RTW_API\s+(?:const\s+)?[a-z_0-9]+\*?\s+\*?(rtw_[a-z0-9_]+)\s*\(", text)


def test_library_exports_every_declared_symbol(rtw):
    names = _header_prototypes()
    assert len(names) >= 14 and len(set(names)) == len(names)
    assert set(names) == set(rtw.EXPORTED_SYMBOLS)
    lib = rtw._lib.load()
    for n in names:
        assert getattr(lib, n) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", str(rtw.LIB_PATH)], capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if " T " in line}
    assert set(names) <= exported
    # nothing but the C-ABI leaks out of the shared object
    assert all(s.startswith("rtw_") for s in exported if not s.startswith("_")), exported


def test_abi_version_and_image_height(rtw):
    lib = rtw._lib.load()
    assert lib.rtw_abi_version() == 3
    for w, h in [(96, 54), (400, 225), (1920, 1080), (200, 112), (1, 0)]:
        assert lib.rtw_image_height(w) == h


def test_struct_layouts_match_header_and_julia(rtw):
    # Camera{Float32}: 7 x Vec3 + lens_radius = 22 floats = 88 bytes, src/camera.jl:1-10; Camera{Float64}: 22 doubles
    assert C.sizeof(rtw.rtw_camera) == 88
    assert C.sizeof(rtw._lib.rtw_camera_f64) == 176
    assert [n for n, _ in rtw._lib.rtw_camera_f64._fields_] == [n for n, _ in rtw.rtw_camera._fields_]
    assert [n for n, _ in rtw.rtw_camera._fields_] == ["origin", "lower_left_corner", "horizontal", "vertical", "u", "v",
                                                       "w", "lens_radius"]
    text = HEADER.read_text()
    body = text[text.index("typedef struct rtw_stats {"):text.index("} rtw_stats;")]
    fields = re.findall(r"\b([a-z_0-9]+);", body)
    assert fields == [n for n, _ in rtw.rtw_stats._fields_]
    assert C.sizeof(rtw.rtw_stats) == 104
    cam = rtw.t_cam1()
    assert cam.as_array().shape == (22,) and cam.as_array().dtype == np.float32


def test_option_and_error_constants_match_header(rtw):
    text = HEADER.read_text()
    for name in ("RTW_OK", "RTW_E_INVALID_ARG", "RTW_E_NO_DEVICE", "RTW_E_NO_SCENE", "RTW_E_UNSUPPORTED",
                 "RTW_E_INTERNAL", "RTW_OPT_MODE", "RTW_OPT_STRIP", "RTW_OPT_BLOCKS_PER_SM", "RTW_OPT_COLLECT_TIMING",
                 "RTW_OPT_RAYS_PER_LANE", "RTW_OPT_SWEEP", "RTW_OPT_COOP", "RTW_OPT_TAIL", "RTW_OPT_WALK", "RTW_WALK_DEFAULT", "RTW_WALK_SLOTS", "RTW_WALK_OWN_RAY", "RTW_TAIL_DEFAULT", "RTW_TAIL_SPLIT",
                 "RTW_TAIL_UNIFIED", "RTW_MODE_FUSED", "RTW_MODE_WAVEFRONT", "RTW_MODE_CTA_WAVEFRONT", "RTW_MODE_GRID"):
        m = re.search(rf"#define\s+{name}\s+\(?(-?\d+)\)?", text)
        assert m, name
        assert int(m.group(1)) == getattr(rtw._lib, name), name


def test_no_cpu_fallback_without_gpu(rtw):
    lib = rtw._lib.load()
    n = C.c_int(-1)
    status = lib.rtw_device_count(C.byref(n))
    if status == 0 and n.value > 0:
        pytest.skip("a GPU is visible; the no-device error path is exercised on the CPU container")
    assert status == rtw._lib.RTW_E_NO_DEVICE and n.value == 0
    ctx = C.c_void_p()
    assert lib.rtw_create(None, 1, C.byref(ctx)) == rtw._lib.RTW_E_NO_DEVICE and not ctx.value
    with pytest.raises(rtw.RtwError) as e:
        rtw.render(rtw.scene_2_spheres(), rtw.t_default_cam(), 96, 16)
    assert e.value.code == rtw._lib.RTW_E_NO_DEVICE
    with pytest.raises(rtw.RtwError):
        rtw.Renderer([0])


def test_null_arguments_are_rejected_not_crashing(rtw):
    lib = rtw._lib.load()
    assert lib.rtw_device_count(None) == rtw._lib.RTW_E_INVALID_ARG
    assert lib.rtw_create(None, 1, None) == rtw._lib.RTW_E_INVALID_ARG
    assert lib.rtw_destroy(None) == 0
    assert lib.rtw_set_scene(None, None, None, None, 0) == rtw._lib.RTW_E_INVALID_ARG
    assert lib.rtw_render(None, None, 96, 1, 16, 1, None, None) == rtw._lib.RTW_E_INVALID_ARG
    assert lib.rtw_set_option(None, 1, 0) == rtw._lib.RTW_E_INVALID_ARG
    assert b"NULL" in lib.rtw_last_error(None)


def test_product_package_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import, load or name it
    pkg = ROOT / "raytracingweekend.jl_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) \
            + list(pkg.rglob("*.jl")) + list(pkg.rglob("*.sh")):
        text = path.read_text()
        assert "librtw_oracle" not in text and "rtwo_" not in text, path
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), path
