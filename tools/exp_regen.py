#!/usr/bin/env python
"""Render-kernel A/B (development): frame hash, ray count and Mrays/s of one render kernel variant.

    TMPT_RENDER_KERNEL=k python tools/exp_regen.py [--scene sponza] [--width 640 --height 360 --spp 16] [--reps 2]

The variant is read once per process (kernels.cu: launch_render), so each variant is its own run; equal sha256
across variants = byte-identical frames."""
import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import toymeshpathtracer_b200 as tm  # noqa: E402
from bench import scene_obj_path  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="sponza")
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--height", type=int, default=360)
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--stats", action="store_true", help="also run the instrumented pass and print its counters")
a = ap.parse_args()
path = scene_obj_path(a.scene)
tris, mn, mx = tm.load_scene(path)
cam = tm.camera_for_scene(path, mn, mx, a.width, a.height)
sc = tm.Scene(tris)
best = 1e30
for _ in range(a.reps + 1):
    rgba, rays, sec = sc.render(cam, a.width, a.height, a.spp)
    best = min(best, sec)
print("lib %s kernel %s %s %dx%dx%d sha %s rays %d  %.1f Mrays/s (%.2f ms)  build %.2f ms" % (
      os.path.basename(os.environ.get("TMPT_LIB", "libtmpt.so")), os.environ.get("TMPT_RENDER_KERNEL", "0"), a.scene, a.width, a.height,
      a.spp, hashlib.sha256(rgba.tobytes()).hexdigest()[:16], rays, rays / best / 1e6, best * 1e3, sc.info()["build_ms"]))
if a.stats:
    import json
    print(json.dumps(sc.traversal_stats(cam, a.width, a.height, a.spp)))
