#!/usr/bin/env python
"""Development tool: run the HitScene tests of tests/test_zz_gpu_fuzz_regressions.py against a stand-in for tm.Scene that is backed
by the host emulation (tests/emu, slab distances fused as on the device) -- to check the TEST LOGIC (names, shapes, classification
thresholds) in the build container, where there is no GPU.  It says nothing about the CUDA path: only `pytest -m gpu` on a B200 does.

    python tools/check_gpu_tests_on_emu.py

(Those tests were written after round 2's GPU minutes were spent; this is how they were vetted before the driver's run.)"""
import importlib
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import toymeshpathtracer_b200 as tm  # noqa: E402
from emu_binding import Emu  # noqa: E402

emu = Emu(defines=["-DTMPT_EMU_FMA=1"], tag="fma")


class EmulatedScene:
    def __init__(self, tris, device=0, flags=0):
        self.s = emu.scene(tris, builder=1 if flags == tm.BUILD_LBVH else 0)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.s.close()

    def HitScene(self, rays, tMin=0.001, tMax=1.0e7, mode=0, payload=True):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        if mode == tm.HIT_SUN:
            return self.s.sun_occluded(rays[:, :3], tmin=tMin, tmax=tMax)[0], None, None, None
        return self.s.hit(rays, tmin=tMin, tmax=tMax, mode={tm.HIT_CLOSEST: 0, tm.HIT_ANY: 1, tm.HIT_BRUTE: 2}[mode])


if __name__ == "__main__":
    tm.Scene = EmulatedScene  # this process only
    z = importlib.import_module("test_zz_gpu_fuzz_regressions")
    from oracle.pyoracle import Oracle
    spec = importlib.util.spec_from_file_location("fuzz_emu", os.path.join(ROOT, "tools", "fuzz_emu.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    orc = Oracle()
    for flags in (0, tm.BUILD_LBVH):
        z.test_rays_that_start_on_shared_vertices_with_tmin_zero(fz, orc, flags)
    z.test_sun_query_from_distant_origins(fz)
    for seed in (0, 2, 7, 12, 15, 20, 37, 41):
        z.test_fuzz_scenes_tree_and_sun_grid_equal_the_scan(fz, seed)
    print("test logic ok on the emulation (the CUDA path is checked by pytest -m gpu on the GPU box)")
