"""raytracingweekend.jl_b200 -- B200-native (sm_100a) implementation of RayTracingWeekend.jl's
render() -> ray_color() -> hit()/scatter() hot path behind a C-ABI, plus the host-side mirror of the
package's API surface (src/RayTracingWeekend.jl:10-31) used to drive it.

    csrc/          hand-written CUDA kernels + the C-ABI library (librtw_b200.so, include/rtw_b200.h)
    julia/         the `ccall` shim a Julia user loads instead of the reference's render()
    host.py        Camera/default_camera, Sphere/materials, scene builders, TRNG (host code in the reference too)
    api.py         render() / Renderer: thin ctypes calls into the C-ABI -- no CPU fallback

The directory name contains a dot, so import it through the repo-root alias module `rtw_b200`.
"""
from ._lib import (RTW_GATHER_NCCL, RTW_GATHER_PEER, RTW_OPT_GATHER, RTW_OPT_SMALL_RENDER)
from ._lib import (EXPORTED_SYMBOLS, LIB_PATH, RTW_MODE_CTA_WAVEFRONT, RTW_MODE_FUSED, RTW_MODE_GRID, RTW_MODE_WAVEFRONT, RTW_OPT_BLOCKS_PER_SM,
                   RTW_OPT_COLLECT_TIMING, RTW_OPT_COOP, RTW_OPT_MODE, RTW_OPT_RAYS_PER_LANE, RTW_OPT_STRIP, RTW_OPT_SWEEP,
                   RTW_OPT_TAIL, RTW_OPT_WALK, RTW_WALK_DEFAULT, RTW_WALK_OWN_RAY, RTW_WALK_SLOTS, RTW_TAIL_DEFAULT, RTW_TAIL_SPLIT, RTW_TAIL_UNIFIED, RtwError, rtw_camera, rtw_stats)
from . import sharding
from .api import (DEFAULT_MAX_DEPTH, DEFAULT_SEED, Renderer, has_variants, render, scene_load, scene_save, write_png, write_ppm)
from .host import (TRNG, Camera, Dielectric, HittableList, Lambertian, Metal, Sphere, Vec3, Xoroshiro128Plus,
                   default_camera, flatten_scene, image_height, near_zero, random_between, reseed,
                   scene_2_spheres, scene_4_spheres, scene_blue_red_spheres, scene_diel_spheres,
                   scene_random_spheres, squared_length, t_cam1, t_cam2, t_default_cam, trand)

__all__ = [
    "Vec3", "squared_length", "near_zero", "TRNG", "reseed", "trand", "random_between", "Xoroshiro128Plus",
    "Sphere", "HittableList", "Lambertian", "Metal", "Dielectric", "Camera", "default_camera", "render", "Renderer",
    "scene_2_spheres", "scene_4_spheres", "scene_blue_red_spheres", "scene_diel_spheres", "scene_random_spheres",
    "flatten_scene", "image_height", "t_default_cam", "t_cam1", "t_cam2", "RtwError", "rtw_camera", "rtw_stats",
    "EXPORTED_SYMBOLS", "LIB_PATH", "DEFAULT_MAX_DEPTH", "DEFAULT_SEED", "sharding", "scene_load", "scene_save",
    "write_png", "write_ppm",
]
