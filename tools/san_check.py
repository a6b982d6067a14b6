#!/usr/bin/env python
"""Small end-to-end run of every entry point (development): build (both builders), HitScene, render with several chunks,
refit, progressive passes.    python tools/san_check.py
(Written as the workload for compute-sanitizer's memcheck; this pool's GPU boxes do not allow the sanitizer to attach
-- gpurun_out/san_memcheck.log -- so it is run plain, as a smoke test of the whole ABI.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import toymeshpathtracer_b200 as tm  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "scenes", "suzanne.npz"))
tris, mn, mx = z["tris"], z["bounds_min"], z["bounds_max"]
rng = np.random.default_rng(3)
o = rng.uniform(mn - 1, mx + 1, (20000, 3)); d = rng.normal(size=(20000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([o, d], 1).astype(np.float32)
w, h = 64, 36
cam = tm.camera_for_scene("suzanne.obj", mn, mx, w, h)
for flags in (tm.BUILD_DEFAULT, tm.BUILD_LBVH):
    with tm.Scene(tris, flags=flags) as s:
        a, b = s.HitScene(rays), s.HitScene(rays, mode=tm.HIT_BRUTE)
        assert (a[0] == b[0]).all()
        s.HitScene(rays, mode=tm.HIT_ANY)
        img, nr, _ = s.render(cam, w, h, 20)
        s.refit((tris.reshape(-1, 3) * np.float32(1.05)).reshape(-1, 9))
        c, e = s.HitScene(rays), s.HitScene(rays, mode=tm.HIT_BRUTE)
        assert (c[0] == e[0]).all()
        s.progressive_begin(w, h)
        s.progressive_pass(cam, 2); s.progressive_pass(cam, 1)
        print("flags", flags, "ok", nr, int((a[0] >= 0).sum()), int((c[0] >= 0).sum()))
