#!/usr/bin/env python
"""Traversal-only experiment: throughput of the batched HitScene kernels on the ray
distribution a Sponza path trace produces (primary, then successive diffuse bounces and their
shadow rays), generated on the GPU with torch.  Development tool -- not part of bench.py."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import toymeshpathtracer_b200 as tm  # noqa: E402
from bench import scene_obj_path  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="sponza")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--bounces", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--sort", default="none", choices=["none", "origin", "origin_dir", "dir_origin"],
                    help="reorder every wave of rays before tracing it: Morton code of the origin (10 bits per axis), optionally with the direction octant")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    path = scene_obj_path(args.scene)
    tris, mn, mx = tm.load_scene(path)
    cam = torch.tensor(tm.camera_for_scene(path, mn, mx, args.width, args.height), device=dev)
    sc = tm.Scene(tris, flags=args.flags)
    print(json.dumps({"bvh": sc.info()}))
    W, H = args.width, args.height
    g = torch.Generator(device=dev); g.manual_seed(1)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    u = ((xs + torch.rand((H, W), device=dev, generator=g)) / W).reshape(-1, 1)
    v = ((ys + torch.rand((H, W), device=dev, generator=g)) / H).reshape(-1, 1)
    origin, llc, hor, ver = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    d = llc + u * hor + v * ver - origin
    d = d / d.norm(dim=1, keepdim=True)
    rays = torch.cat([origin.expand_as(d), d], 1).contiguous().float()
    light = torch.tensor([-0.7, 1.0, 0.5], device=dev); light = light / light.norm()
    st = torch.cuda.Stream()

    lo_t = torch.tensor(mn, device=dev, dtype=torch.float32)
    ext_t = torch.tensor(mx - mn, device=dev, dtype=torch.float32).clamp_min(1e-20)

    def part1by2(v):
        v = v & 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v

    def reorder(rays):
        if args.sort == "none":
            return rays, 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        q = (((rays[:, 0:3] - lo_t) / ext_t).clamp(0, 0.999999) * 1024).to(torch.int64)
        key = (part1by2(q[:, 0]) << 2) | (part1by2(q[:, 1]) << 1) | part1by2(q[:, 2])
        octant = ((rays[:, 3] < 0).to(torch.int64) << 2) | ((rays[:, 4] < 0).to(torch.int64) << 1) | (rays[:, 5] < 0).to(torch.int64)
        if args.sort == "origin_dir":
            key = (key << 3) | octant
        elif args.sort == "dir_origin":
            key = (octant << 30) | key
        out = rays[torch.argsort(key)].contiguous()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1)

    def timed(rays, mode):
        rays, sort_ms = reorder(rays)
        n = rays.shape[0]
        ids = torch.empty(n, dtype=torch.int32, device=dev)
        t = torch.empty(n, device=dev); pos = torch.empty((n, 3), device=dev); nrm = torch.empty((n, 3), device=dev)
        best = 1e30
        with torch.cuda.stream(st):
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                sc.hit_scene_device(rays.data_ptr(), n, ids.data_ptr(), t.data_ptr(), pos.data_ptr(), nrm.data_ptr(), mode=mode, stream=st.cuda_stream)
                e1.record(st)
                st.synchronize()
                best = min(best, e0.elapsed_time(e1))
        stats = sc.hit_scene_stats(rays.data_ptr(), n, mode=mode)
        stats["sort_ms"] = sort_ms
        return ids, pos, nrm, best, stats

    tot_rays, tot_ms = 0, 0.0
    for b in range(args.bounces + 1):
        torch.cuda.synchronize()
        ids, pos, nrm, ms, stats = timed(rays, tm.HIT_CLOSEST)
        n = rays.shape[0]
        tot_rays += n; tot_ms += ms
        print(f"bounce {b}: closest {n:8d} rays {ms:8.3f} ms {n / ms / 1e3:8.1f} Mrays/s  nodes/ray {stats['node_visits_per_ray']:.1f} tris/ray {stats['tri_tests_per_ray']:.1f} hit {stats['hit_rate']:.2f}"
              f"  lanes/iter {stats['lanes_with_a_ray']:.1f} lanes/node {stats['lanes_per_node_step']:.1f} lanes/tri {stats['lanes_per_tri_test']:.1f}  (torch sort {stats['sort_ms']:.2f} ms)")
        hit = ids >= 0
        pos, nrm = pos[hit], nrm[hit]
        if pos.shape[0] == 0:
            break
        srays = torch.cat([pos, light.expand_as(pos)], 1).contiguous()
        _, _, _, ms, stats = timed(srays, tm.HIT_ANY)
        tot_rays += srays.shape[0]; tot_ms += ms
        print(f"          shadow  {srays.shape[0]:8d} rays {ms:8.3f} ms {srays.shape[0] / ms / 1e3:8.1f} Mrays/s  nodes/ray {stats['node_visits_per_ray']:.1f} tris/ray {stats['tri_tests_per_ray']:.1f} hit {stats['hit_rate']:.2f}"
              f"  lanes/iter {stats['lanes_with_a_ray']:.1f} lanes/node {stats['lanes_per_node_step']:.1f} lanes/tri {stats['lanes_per_tri_test']:.1f}")
        r = torch.randn(pos.shape, device=dev, generator=g); r = r / r.norm(dim=1, keepdim=True)
        nd = nrm + r
        nd = nd / nd.norm(dim=1, keepdim=True).clamp_min(1e-20)
        rays = torch.cat([pos, nd], 1).contiguous()
    print(f"TOTAL {tot_rays} rays {tot_ms:.3f} ms -> {tot_rays / tot_ms / 1e3:.1f} Mrays/s (traversal only)")


if __name__ == "__main__":
    main()
