#!/usr/bin/env python
"""How full could the warps be?  (development; CPU only)

Walk lengths (iterations of bvh::walk_step, host emulation of the product's traversal) of the bounce and shadow rays of the golden
Sponza ray set, then a simulation of a warp tracing them:
  lockstep (what k_render does): 32 shadow rays together, then 32 bounce rays: the warp takes max + max iterations;
  pool: P rays per lane in one per-warp pool (P = 2: one bounce of 32 paths; 4 / 8: two / four paths per lane in flight); an idle
        lane takes the next ray, but only when >= G lanes are idle (the gate that keeps the divergent fetch code out of most iterations).
Prints lane utilisation = useful lane iterations / (32 x warp iterations).  Result (profiles/r2_tuning_sweeps.txt): lockstep 0.525 -- the
GPU measures 17.2 / 32 = 0.54 -- pools of 2 / 4 / 8 rays per lane 0.62-0.65 / 0.70-0.79 / 0.76-0.88."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_rays, sponza_scene  # noqa: E402
from emu_binding import Emu, _p  # noqa: E402


def main():
    g = load_rays("sponza")
    e = Emu()
    s = e.scene(sponza_scene()[0])

    def iters(rays, any_hit):
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros(len(rays), np.int32)
        e.L.emu_ray_iters(s.h, _p(rays), C.c_long(len(rays)), C.c_float(0.001), C.c_float(1.0e7), int(any_hit), _p(out))
        return out

    b, sh = iters(g["rays"][g["kind"] == 1], False), iters(g["rays"][g["kind"] == 2], True)
    for name, a in (("bounce", b), ("shadow", sh)):
        print(f"{name}: n {len(a)} mean {a.mean():.2f} p50 {np.percentile(a, 50):.0f} p90 {np.percentile(a, 90):.0f} p99 {np.percentile(a, 99):.0f} max {a.max()}")
    rng = np.random.default_rng(1)
    lt = lw = 0
    for _ in range(4000):
        a, c = rng.choice(sh, 32), rng.choice(b, 32)
        lt += a.max() + c.max()
        lw += a.sum() + c.sum()
    print("lockstep utilisation %.3f" % (lw / (32 * lt)))

    def sim(per_lane, gate, trials=600):
        tt = ww = 0
        for _ in range(trials):
            pool = list(np.concatenate([rng.choice(sh, 16 * per_lane), rng.choice(b, 16 * per_lane)]))
            rng.shuffle(pool)
            ww += sum(pool)
            lane = np.zeros(32, np.int64)
            t = 0
            while pool or lane.any():
                idle = np.flatnonzero(lane == 0)
                if pool and (len(idle) >= gate or not lane.any()):
                    for l in idle:
                        if not pool:
                            break
                        lane[l] = pool.pop()
                if not (lane > 0).any():
                    break
                lane[lane > 0] -= 1
                t += 1
            tt += t
        return ww / (32 * tt)

    for per_lane in (2, 4, 8):
        print("rays per lane", per_lane, " ".join("gate %d: %.3f" % (gt, sim(per_lane, gt)) for gt in (1, 4, 8, 12)))


if __name__ == "__main__":
    main()
