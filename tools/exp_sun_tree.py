#!/usr/bin/env python
"""Would a SUN-ALIGNED tree make the shadow rays cheaper?  (development experiment)

Every shadow ray of the path tracer has the same direction (kLightDir, main.cpp:36).  This script rotates the whole scene so that
the sun direction becomes the +z axis, builds the ordinary BVH on the rotated triangles and traces the SAME shadow rays in both
frames with the any-hit kernel: node visits / triangle tests per ray and Mrays/s, for waves of shadow rays as the megakernel
sees them (pixel order, bounces 0..5) -- an upper bound on what a second, sun-aligned tree could buy with the general slab code."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import toymeshpathtracer_b200 as tm  # noqa: E402
from bench import scene_obj_path  # noqa: E402

dev = torch.device("cuda", 0)
path = scene_obj_path("sponza")
tris, mn, mx = tm.load_scene(path)
W, H = 1920, 1080
cam = torch.tensor(tm.camera_for_scene(path, mn, mx, W, H), device=dev)
light = np.array([-0.7, 1.0, 0.5]); light /= np.linalg.norm(light)
# rotation with rows (u, v, w = light)
a = np.array([1.0, 0, 0]) if abs(light[0]) < 0.9 else np.array([0, 1.0, 0])
u = np.cross(light, a); u /= np.linalg.norm(u)
v = np.cross(light, u)
R = np.stack([u, v, light])           # x' = R x
rtris = (tris.reshape(-1, 3).astype(np.float64) @ R.T).astype(np.float32).reshape(-1, 9)
sc, rsc = tm.Scene(tris), tm.Scene(rtris)
Rt = torch.tensor(R, device=dev, dtype=torch.float32)
lt = torch.tensor(light, device=dev, dtype=torch.float32)
g = torch.Generator(device=dev); g.manual_seed(1)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
uu = ((xs + torch.rand((H, W), device=dev, generator=g)) / W).reshape(-1, 1)
vv = ((ys + torch.rand((H, W), device=dev, generator=g)) / H).reshape(-1, 1)
d = cam[3:6] + uu * cam[6:9] + vv * cam[9:12] - cam[0:3]
d = d / d.norm(dim=1, keepdim=True)
rays = torch.cat([cam[0:3].expand_as(d), d], 1).contiguous().float()
st = torch.cuda.Stream()


def timed(scene, r6, mode):
    n = r6.shape[0]
    ids = torch.empty(n, dtype=torch.int32, device=dev); t = torch.empty(n, device=dev)
    pos = torch.empty((n, 3), device=dev); nrm = torch.empty((n, 3), device=dev)
    best = 1e30
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            scene.hit_scene_device(r6.data_ptr(), n, ids.data_ptr(), t.data_ptr(), pos.data_ptr(), nrm.data_ptr(), mode=mode, stream=st.cuda_stream)
            e1.record(st); st.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return ids, pos, nrm, best, scene.hit_scene_stats(r6.data_ptr(), n, mode=mode)


for b in range(6):
    ids, pos, nrm, ms, stc = timed(sc, rays, tm.HIT_CLOSEST)
    hit = ids >= 0
    pos, nrm = pos[hit], nrm[hit]
    srays = torch.cat([pos, lt.expand_as(pos)], 1).contiguous()
    ida, _, _, msa, sa = timed(sc, srays, tm.HIT_ANY)
    rs = torch.cat([pos @ Rt.T, torch.tensor([0.0, 0.0, 1.0], device=dev).expand_as(pos)], 1).contiguous()
    idb, _, _, msb, sb = timed(rsc, rs, tm.HIT_ANY)
    n = srays.shape[0]
    agree = float(((ida >= 0) == (idb >= 0)).float().mean())
    print(f"bounce {b}: {n} shadow rays | world tree {n / msa / 1e3:7.1f} Mrays/s nodes {sa['node_visits_per_ray']:.2f} tris {sa['tri_tests_per_ray']:.2f} lanes {sa['lanes_with_a_ray']:.1f}"
          f" | sun tree {n / msb / 1e3:7.1f} Mrays/s nodes {sb['node_visits_per_ray']:.2f} tris {sb['tri_tests_per_ray']:.2f} lanes {sb['lanes_with_a_ray']:.1f} | same visibility {agree:.5f}")
    r = torch.randn(pos.shape, device=dev, generator=g); r = r / r.norm(dim=1, keepdim=True)
    nd = nrm + r; nd = nd / nd.norm(dim=1, keepdim=True).clamp_min(1e-20)
    rays = torch.cat([pos, nd], 1).contiguous()
