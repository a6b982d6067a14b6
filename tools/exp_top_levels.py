#!/usr/bin/env python
"""How much of the walk happens in the top of the tree?  (development; CPU only)

Node visits per node index of the golden Sponza rays (host emulation of the product's traversal), then the share of all node
visits that falls on the first N nodes -- the nodes a shared-memory copy of the tree's top would serve.  Also the depth of each
node, to see whether index order is level order."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_rays, sponza_scene  # noqa: E402
from emu_binding import Emu, _p  # noqa: E402

g = load_rays("sponza")
e = Emu()
s = e.scene(sponza_scene()[0])
info = np.zeros(5, np.uint32)
e.L.emu_scene_info(s.h, _p(info))
n_nodes = int(info[0])
print("scene info", info)
for kind, name, any_hit in ((1, "bounce", 0), (2, "shadow", 1)):
    rays = np.ascontiguousarray(g["rays"][g["kind"] == kind], np.float32)
    hist = np.zeros(max(n_nodes, 1 << 16), np.uint64)
    e.L.emu_node_visits(s.h, _p(rays), C.c_long(len(rays)), C.c_float(0.001), C.c_float(1.0e7), any_hit, _p(hist))
    tot = hist.sum()
    cum = np.cumsum(hist) / tot
    print(f"{name}: {len(rays)} rays, {tot / len(rays):.2f} node visits per ray; share on the first N nodes (index order): " +
          ", ".join(f"{n}: {cum[n - 1]:.3f}" for n in (1, 5, 21, 85, 341, 585, 1170, 2340, 4680)))
    top = np.sort(hist)[::-1]
    cumt = np.cumsum(top) / tot
    print(f"   ... on the N most visited nodes: " + ", ".join(f"{n}: {cumt[n - 1]:.3f}" for n in (85, 341, 585, 1170, 2340)))
