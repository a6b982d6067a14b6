"""Helper of tests/test_gpu_parity.py (TEST INFRASTRUCTURE): render one frame of a golden scene through the C ABI in a fresh
process -- launch configuration and library variant are read from the environment once per process (TMPT_RENDER_CFG,
TMPT_RENDER_PATHS, TMPT_RENDER_KERNEL, TMPT_LIB) -- and save frame, ray count and the render kernel that ran.

    python tests/render_probe.py <scene> <w> <h> <spp> <out.npz> [progressive chunk counts ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import toymeshpathtracer_b200 as tm  # noqa: E402
from conftest import load_scene, sponza_scene  # noqa: E402

name, w, h, spp, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
passes = [int(x) for x in sys.argv[6:]]
if name == "sponza":
    t, mn, mx = sponza_scene()
    sc = {"tris": t, "bounds_min": mn, "bounds_max": mx}
else:
    sc = load_scene(name)
cam = tm.camera_for_scene(f"{name}.obj", sc["bounds_min"], sc["bounds_max"], w, h)
with tm.Scene(sc["tris"]) as s:
    if passes:
        s.progressive_begin(w, h)
        rays = 0
        for n in passes:
            img, r, _, _ = s.progressive_pass(cam, n)
            rays += r
    else:
        img, rays, _ = s.render(cam, w, h, spp)
    kernel, escape = s.render_kernel_choice()
np.savez(out, img=img, rays=np.int64(rays), kernel=np.int64(kernel), escape=np.float64(escape))
