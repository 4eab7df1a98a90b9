"""Renders scene_random_spheres with the t_cam1 camera and writes a PNG through the library's own writer
(rtw_accumulate -> rtw_resolve_rgb8 -> rtw_write_png).  Usage: python tools/render_png.py out.png [W] [spp] [depth]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rtw_b200 as R  # noqa: E402

out = sys.argv[1]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 480
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 256
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 50
R.reseed()
scene = R.scene_random_spheres()
with R.Renderer([0]) as r:
    r.set_scene(scene)
    passes = 4
    for p in range(passes):  # progressive: four passes == one render
        first = p * spp // passes
        r.accumulate(R.t_cam1(), W, first, (p + 1) * spp // passes - first, spp, max_depth=depth)
    R.write_png(out, r.resolve_rgb8())
    print("wrote", out, r.progress())
