#!/usr/bin/env python
"""Per-rank work balance of a partition (--stripe 0: tile interleave, the default; > 0: row stripes), emulated on ONE GPU:
render rank k's share of a `world`-way split for every k and time each (development tool)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import toymeshpathtracer_b200 as tm
from toymeshpathtracer_b200 import multigpu
from bench import scene_obj_path

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--stripe", type=int, default=multigpu.DEFAULT_STRIPE_ROWS)
ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080); ap.add_argument("--spp", type=int, default=64)
a = ap.parse_args()
path = scene_obj_path("sponza")
tris, mn, mx = tm.load_scene(path)
cam = tm.camera_for_scene(path, mn, mx, a.width, a.height)
sc = tm.Scene(tris)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
frame = torch.zeros((a.height, a.width, 4), dtype=torch.uint8, device=dev)
rays = torch.zeros(a.world, dtype=torch.int64, device=dev)
times = []
for rep in range(2):
    times = []
    for r in range(a.world):
        with torch.cuda.stream(st):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            sc.render_stripes(cam, a.width, a.height, a.spp, a.stripe, r, a.world, 0, rays[r:].data_ptr(), peer_frame_ptr=frame.data_ptr(), stream=st.cuda_stream)
            e1.record(st)
        st.synchronize()
        times.append(e0.elapsed_time(e1))
    rays_list = rays.tolist(); rays.zero_()
print("per-rank ms:", " ".join(f"{t:.1f}" for t in times))
print("per-rank Mrays:", " ".join(f"{x/1e6:.0f}" for x in rays_list))
print(f"sum {sum(times):.1f} ms  max {max(times):.1f} ms  ideal {sum(times)/a.world:.1f} ms  -> partition efficiency {sum(times)/a.world/max(times):.3f}")
if a.world > 16:
    import numpy as np
    t = np.array(times); r = np.array(rays_list) / 1e6
    order = np.argsort(-t)[:12]
    print("slowest ranks (stripe index): ", [(int(k), round(float(t[k]), 2), round(float(r[k]), 1)) for k in order])
    print("median ms", float(np.median(t)), "median Mrays", float(np.median(r)))
