import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        import rtw_b200
        import ctypes as C
        lib = rtw_b200._lib.load()
        n = C.c_int()
        return lib.rtw_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.load()
    return binding


@pytest.fixture(scope="session")
def rtw():
    import rtw_b200
    return rtw_b200


@pytest.fixture(scope="session")
def renderer(rtw):
    """A GPU renderer; GPU tests FAIL (not skip) when the CUDA library cannot run -- no silent fallback."""
    r = rtw.Renderer([0])
    # the parity tests are about the persistent kernel: keep small images on it (tests/test_gpu_small_render.py covers
    # the single-launch latency path that small renders take by default)
    r.set_option(rtw.RTW_OPT_SMALL_RENDER, 0)
    yield r
    r.close()


@pytest.fixture(scope="session")
def scenes(rtw):
    """Flattened fixture scenes, built deterministically from the host mirror (reseed() first)."""
    out = {}
    out["two"] = rtw.flatten_scene(rtw.scene_2_spheres())
    out["four"] = rtw.flatten_scene(rtw.scene_4_spheres())
    out["diel"] = rtw.flatten_scene(rtw.scene_diel_spheres())
    # hollow-glass bubble: outer glass sphere + inner one with NEGATIVE radius (src/scenes.jl:34-36)
    out["bubble"] = rtw.flatten_scene(
        rtw.scene_diel_spheres() + [rtw.Sphere(rtw.Vec3(-1, 0, -1), -0.4, rtw.Dielectric(1.5))])
    out["bluered"] = rtw.flatten_scene(rtw.scene_blue_red_spheres())
    rtw.reseed()
    out["random"] = rtw.flatten_scene(rtw.scene_random_spheres())
    return out
