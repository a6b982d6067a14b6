#!/usr/bin/env python
"""Fuzz the product's structures on the CPU -- TEST TOOLING (runs the host emulation, tests/emu: the product's own
__host__ __device__ headers compiled with g++; never part of the product path).

    python tools/fuzz_emu.py [first_seed] [last_seed]             # default 0 400, all cores
    python tools/fuzz_emu.py --render [first_seed] [last_seed]    # small frames of random scenes: emulation == oracle, bytes and ray count
    python tools/fuzz_emu.py --refit [first_seed] [last_seed]     # scenes moved three times and refitted: tree and grid == scan of the moved triangles
    (any of them with --fma: the emulation rounds the slab test's distances with one fused multiply-add, as the device does)

A seed makes one random scene -- 1 .. 2000 triangles of one of nine kinds (blobs, sizes over five decades, coplanar overlapping
pieces, slivers, triangles edge-on to the sun, duplicates, a height-field mesh with shared vertices under three giants, triangles
flat in the sun's depth, long slivers that almost contain the sun direction), scaled by 1e-4 .. 1e5 and sometimes shifted far off the origin -- and 21 000 query origins on, a hair
above / below, at vertices and edge midpoints of, around and far below its triangles.  Per seed, for both builders:

  tree closest hit  == all-triangle scan   (id, t bits; rays along the sun and in random / axis-parallel directions)
  tree any hit      == scan
  sun grid          == scan                 (bvh::sun_query, what k_hit_scene<TMPT_HIT_SUN> runs)

at tMin = 0.001 (the integrator's) and tMin = 0 (a caller's choice: rays that START on a surface hit it at t = +-0).

What it found (round 2, both fixed, regression tests in tests/test_emu_logic.py and tests/test_zz_gpu_fuzz_regressions.py):
  * pop-time cull with the child slot still in the key's low bits (bvh.cuh: walk_step) -- with tMin = 0 a ray starting on a shared
    vertex lost the lower-index triangles of the tie;
  * TMPT_HIT_SUN for origins a million scene sizes away (bvh.cuh: sun_query now takes the scan beyond the far limit).
What it documents: GARBAGE HITS.  When a ray lies (almost) in the plane of a triangle with long edges -- a zero-area
triangle is in every ray's "plane" -- the determinant of the reference's test is rounding noise of size ~2^-24 |e1| |e2|, which
passes `Epsilon` = 1e-5 once the edges are longer than about 13 units; u, v, t are then noise too and the test can accept a "hit"
anywhere along the ray, scene sizes away from the triangle.  The all-triangle scan reports such a hit for every ray; the tree only
for rays that cross the triangle's padded box, the sun grid only for origins whose projection falls into the triangle's cells (the
reference: only for rays that cross an octree leaf holding it).  The tool classifies the scan's winner geometrically
(`garbage_hits`), counts the differences that involve one, and fails on any other difference.  Kinds "duplicates+degenerate" and "grazing-slivers" exist to produce them;
`make_scene(degenerate=False)` leaves the zero-area triangles out.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

_EMU = None
_L = None


def _emu():
    global _EMU
    if _EMU is None:
        from emu_binding import Emu
        # --fma: the slab distances rounded as the device rounds them (one fused multiply-add; bvh.cuh: fmaf_).  The plain build
        # uses a * b + c -- the walk's answers must not depend on which.
        _EMU = Emu(defines=["-DTMPT_EMU_FMA=1"], tag="fma") if FMA else Emu()
    return _EMU


FMA = "--fma" in sys.argv


def light_dir():
    """kLightDir as the reference computes it (main.cpp:36): read from a golden shadow ray."""
    global _L
    if _L is None:
        g = np.load(os.path.join(ROOT, "tests", "golden", "rays", "cube.npz"))
        _L = np.ascontiguousarray(g["rays"][g["kind"] == 2][0, 3:6], np.float32)
    return _L


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


KINDS = ("blobs", "log-sizes", "coplanar", "slivers", "edge-on", "duplicates+degenerate", "mesh+giants", "flat-in-depth", "grazing-slivers")


def make_scene(rng, kind=None, degenerate=True):
    """-> (tris9 float32 [n, 9], scale, kind)"""
    L = light_dir().astype(np.float64)
    n = int(rng.choice([1, 2, 5, 17, 64, 300, 2000]))
    scale = float(10 ** rng.uniform(-4, 5))
    shift = rng.uniform(-1, 1, 3) * scale * float(10 ** rng.uniform(-2, 3)) * (rng.random() < 0.5)
    if kind is None:
        kind = int(rng.integers(0, 9))
    c = rng.uniform(-1, 1, (n, 1, 3))
    if kind == 0:
        t = c + rng.normal(scale=0.1, size=(n, 3, 3))
    elif kind == 1:
        t = c + rng.normal(size=(n, 3, 3)) * (10 ** rng.uniform(-5, 0.5, (n, 1, 1)))
    elif kind == 2:
        t = c + rng.normal(scale=0.3, size=(n, 3, 3))
        ax = rng.integers(0, 3)
        t[:, :, ax] = np.round(c[:, :, ax] * 3) / 3
    elif kind == 3:
        a = c + rng.normal(scale=0.2, size=(n, 1, 3))
        b = a + rng.normal(scale=0.5, size=(n, 1, 3))
        t = np.concatenate([a, b, (a + b) / 2 + rng.normal(size=(n, 1, 3)) * 1e-6], 1)
    elif kind == 4:
        a = c + rng.normal(scale=0.2, size=(n, 1, 3))
        t = np.concatenate([a, a + L[None, None, :] * rng.uniform(0.01, 1, (n, 1, 1)), a + rng.normal(scale=0.2, size=(n, 1, 3))], 1)
    elif kind == 5:
        base = c[: max(1, n // 4)] + rng.normal(scale=0.2, size=(max(1, n // 4), 3, 3))
        t = base[rng.integers(0, len(base), n)]
        z1, z2 = rng.random(n) < 0.2, rng.random(n) < 0.1
        if degenerate:
            t[z1, 2] = t[z1, 1]
            t[z2, 1] = t[z2, 0]
            t[z2, 2] = t[z2, 0]
    elif kind == 6:
        m = int(np.sqrt(n / 2)) + 1
        xs, ys = np.meshgrid(np.linspace(-1, 1, m + 1), np.linspace(-1, 1, m + 1))
        P = np.stack([xs, 0.1 * np.sin(3 * xs) * np.cos(2 * ys), ys], -1)
        q = []
        for i in range(m):
            for j in range(m):
                q.append([P[i, j], P[i + 1, j], P[i, j + 1]])
                q.append([P[i + 1, j], P[i + 1, j + 1], P[i, j + 1]])
        t = np.concatenate([np.array(q), rng.normal(size=(3, 3, 3)) * 30])
    elif kind == 7:
        t = c + rng.normal(scale=0.2, size=(n, 3, 3))
        w = (t * L).sum(-1, keepdims=True)
        t = t - w * L * (1 - 10 ** rng.uniform(-8, 0, (n, 1, 1)))
    else:  # long thin triangles that (almost) contain the sun direction: a shadow ray from below one lies IN its plane
        a = c + rng.normal(scale=0.2, size=(n, 1, 3))
        perp = rng.normal(size=(n, 1, 3))
        perp -= (perp * L).sum(-1, keepdims=True) * L
        t = np.concatenate([a, a + L[None, None, :] * rng.uniform(0.05, 1, (n, 1, 1)),
                            a + L * rng.uniform(-0.5, 1.5, (n, 1, 1)) + perp * 10 ** rng.uniform(-8, -3, (n, 1, 1))], 1)
    return np.ascontiguousarray((t * scale + shift).reshape(-1, 9), np.float32), scale, kind


def make_origins(rng, tris, scale, k=3000):
    L = light_dir().astype(np.float64)
    v = tris.reshape(-1, 3, 3).astype(np.float64)
    pick = rng.integers(0, len(v), k)
    on = (v[pick] * rng.dirichlet([0.4, 0.4, 0.4], k)[:, :, None]).sum(1)
    mn, mx = v.reshape(-1, 3).min(0), v.reshape(-1, 3).max(0)
    ext = np.maximum(mx - mn, 1e-30)
    o = np.concatenate([on, on - L * scale * 1e-3, on + L * scale * 1e-5, v[pick, 0], (v[pick, 0] + v[pick, 1]) / 2,
                        rng.uniform(mn - 0.3 * ext, mx + 0.3 * ext, (k, 3)), on - L * ext.max() * rng.uniform(0, 3, (k, 1))])
    return np.ascontiguousarray(o, np.float32)


def zero_area(tris):
    v = tris.reshape(-1, 3, 3).astype(np.float64)
    return np.linalg.norm(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), axis=1) == 0.0


def garbage_hits(tris, rays, ids, t):
    """Which accepted hits are rounding noise as far as the TREE is concerned: the hit point o + t d (binary64) lies outside the
    triangle's bounding box grown by HALF the padding the tree gives it (build_logic.cuh: pad_for = 1e-3 of the triangle's diagonal
    + 64 ulp of the scene's largest |coordinate|).  A hit that is not flagged lies inside the padded box with half the padding to
    spare for the slab arithmetic: the tree must find it.  (ids < 0 -> False.)"""
    v = tris.reshape(-1, 3, 3).astype(np.float64)
    tri = v[np.maximum(ids, 0)]
    lo, hi = tri.min(1), tri.max(1)
    grow = (0.5 * (1.0e-3 * np.linalg.norm(hi - lo, axis=1) + 64 * 2.0 ** -23 * np.abs(v).max()))[:, None]
    r = rays.astype(np.float64)
    p = r[:, :3] + t.astype(np.float64)[:, None] * r[:, 3:6]
    return (ids >= 0) & ((p < lo - grow) | (p > hi + grow)).any(1)


def _seg_dist2(p, a, b):
    ab, ap = b - a, p - a
    s = np.clip((ap * ab).sum(1) / np.maximum((ab * ab).sum(1), 1e-300), 0.0, 1.0)
    return np.linalg.norm(ap - ab * s[:, None], axis=1)


def off_footprint(tris, origins, ids):
    """The same question for the SUN GRID: does the origin's projection along the sun lie outside the projection of triangle `ids`
    by more than half the grid's pad (sungrid.cuh: setup_view)?  Then the triangle need not be in the origin's cell, and a "hit"
    the exact test reports for it is noise (a real hit along L projects INTO the triangle).  (ids < 0 -> False.)"""
    L = light_dir().astype(np.float64)
    L /= np.linalg.norm(L)
    U = np.cross(L, [1.0, 0.0, 0.0] if abs(L[0]) < 0.9 else [0.0, 1.0, 0.0])
    U /= np.linalg.norm(U)
    V = np.cross(L, U)
    v = tris.reshape(-1, 3, 3).astype(np.float64)
    lo, hi = v.reshape(-1, 3).min(0), v.reshape(-1, 3).max(0)
    corners = np.array([[(hi if c >> k & 1 else lo)[k] for k in range(3)] for c in range(8)])
    cu, cv = corners @ U, corners @ V
    pad = max(max(np.ptp(cu), np.ptp(cv)) * 2.0 ** -14, np.abs(np.concatenate([cu, cv, corners @ L])).max() * 2.0 ** -16)
    tri = v[np.maximum(ids, 0)]
    a, b, c = (np.stack([tri[:, k] @ U, tri[:, k] @ V], 1) for k in range(3))
    q = np.stack([origins.astype(np.float64) @ U, origins.astype(np.float64) @ V], 1)
    cross = lambda x, y: x[:, 0] * y[:, 1] - x[:, 1] * y[:, 0]
    s0, s1, s2 = cross(b - a, q - a), cross(c - b, q - b), cross(a - c, q - c)
    inside = ((s0 >= 0) & (s1 >= 0) & (s2 >= 0)) | ((s0 <= 0) & (s1 <= 0) & (s2 <= 0))
    dist = np.where(inside, 0.0, np.minimum(np.minimum(_seg_dist2(q, a, b), _seg_dist2(q, b, c)), _seg_dist2(q, c, a)))
    return (ids >= 0) & (dist > 0.5 * pad)


def run_seed(seed, degenerate=True, kind=None):
    """-> (seed, kind, triangles, scale, failures, documented): `failures` must be empty; `documented` counts the closest-hit
    differences where the scan's winner is a garbage hit (see the module docstring)."""
    rng = np.random.default_rng(seed)
    tris, scale, kind = make_scene(rng, kind, degenerate)
    o = make_origins(rng, tris, scale)
    L = light_dir()
    bad, documented = [], 0
    for builder in (0, 1):
        s = _emu().scene(tris, builder=builder)
        if s.info()["status"] != 0:
            bad.append(("status", builder, s.info()))
            continue
        d = rng.normal(size=o.shape)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        ax = rng.random(len(d)) < 0.1
        d[ax] = np.eye(3)[rng.integers(0, 3, ax.sum())] * rng.choice([-1, 1], (ax.sum(), 1))
        for label, rays in (("sun", np.concatenate([o, np.broadcast_to(L, o.shape)], 1).astype(np.float32)),
                            ("random", np.concatenate([o, d], 1).astype(np.float32))):
            for tmin in (0.001, 0.0):
                scan = s.hit(rays, tmin=tmin, mode=2)
                tree = s.hit(rays, tmin=tmin, mode=0)
                anyh = s.hit(rays, tmin=tmin, mode=1)[0] >= 0
                hit = scan[0] >= 0
                # a difference is "documented" when the scan's winner is a garbage hit (module docstring): rounding noise that
                # passed the reference's determinant test, at a point off the triangle
                noise = garbage_hits(tris, rays, scan[0], scan[1])
                if label == "sun" and builder == 0:
                    got = s.sun_occluded(o, tmin=tmin)[0] > 0
                    off = off_footprint(tris, o, scan[0])
                    documented += int(((got != hit) & off).sum())
                    if ((got != hit) & ~off).any():
                        bad.append(("sun-grid", tmin, int(((got != hit) & ~off).sum())))
                documented += int(((anyh != hit) & noise).sum())
                if ((anyh != hit) & ~noise).any():
                    bad.append(("any-hit", label, builder, tmin, int(((anyh != hit) & ~noise).sum())))
                m = (tree[0] != scan[0]) | (hit & (bits(tree[1]) != bits(scan[1])))
                documented += int((m & noise).sum())
                if (m & ~noise).any():
                    bad.append(("closest", label, builder, tmin, int((m & ~noise).sum())))
        s.close()
    return seed, KINDS[kind], len(tris), scale, bad, documented


_ORC = None


def render_seed(seed):
    """Integrator level: a small frame of a random scene + the loader's floor (no zero-area triangles, no grazing slivers; at most 400
    triangles: the oracle scans them all; 2.2 rays per camera sample on average)
    through the host emulation of the product's path -- tree, sun grid, per-pixel RNG streams -- against the oracle in the same RNG
    mode: the frame's bytes and the ray count must be equal.  -> (seed, kind, triangles, scale, failures)"""
    global _ORC
    if _ORC is None:
        from oracle.pyoracle import Oracle
        _ORC = Oracle()
    rng = np.random.default_rng(seed)
    tris, scale, kind = make_scene(rng, degenerate=False)
    if KINDS[kind] == "grazing-slivers":  # (made to provoke garbage hits, which the oracle's scan reports and the tree does not)
        tris, scale, kind = make_scene(rng, kind=3)
    tris, mn, mx = _ORC.add_floor(tris[:400])  # the two floor triangles of LoadScene (main.cpp:150-162): paths bounce, shadow rays get shot
    w, h, spp = 24, 16, int(rng.choice([1, 3, 8]))
    cam = _ORC.camera_for_scene(mn, mx, w, h)
    oimg, orays = _ORC.render(tris, cam, w, h, spp, threads=1)
    bad = []
    for builder in (0, 1):
        s = _emu().scene(tris, builder=builder)
        img, rays = s.render(cam, w, h, spp)
        if rays != orays or (img != oimg).any():
            bad.append(("frame", builder, rays, orays, int((img != oimg).any(-1).sum())))
        s.close()
    render_seed.rays_per_sample = orays / (w * h * spp)
    return seed, KINDS[kind], len(tris), scale, bad


def refit_seed(seed):
    """tmpt_scene_refit's logic: a random scene (no zero-area triangles) moved three times -- every vertex on its own, whole
    triangles, an anisotropic scale + shift of everything; amplitudes from 1e-4 to 2 scene sizes -- and refitted in place (tree
    boxes recomputed bottom-up, sun grid rebuilt); after each move tree and grid must equal the scan over the MOVED triangles.
    -> (seed, kind, triangles, scale, failures)"""
    rng = np.random.default_rng(seed)
    tris, scale, kind = make_scene(rng, degenerate=False)
    L = light_dir()
    bad = []
    for builder in (0, 1):
        s = _emu().scene(tris, builder=builder)
        cur = tris
        for step in range(3):
            amp = scale * 10 ** rng.uniform(-4, 0.3)
            mode = int(rng.integers(0, 3))
            v = cur.reshape(-1, 3, 3).astype(np.float64)
            if mode == 0:
                v = v + rng.normal(scale=amp, size=v.shape)
            elif mode == 1:
                v = v + rng.normal(scale=amp, size=(len(v), 1, 3))
            else:
                v = (v - v.mean((0, 1))) * rng.uniform(0.2, 3.0, 3) + v.mean((0, 1)) + rng.normal(scale=amp, size=3)
            cur = np.ascontiguousarray(v.reshape(-1, 9), np.float32)
            s.refit(cur)
            o = make_origins(rng, cur, scale, k=1000)
            d = rng.normal(size=o.shape)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            for label, rays in (("sun", np.concatenate([o, np.broadcast_to(L, o.shape)], 1).astype(np.float32)),
                                ("random", np.concatenate([o, d], 1).astype(np.float32))):
                scan, tree = s.hit(rays, mode=2), s.hit(rays, mode=0)
                anyh, hit = s.hit(rays, mode=1)[0] >= 0, scan[0] >= 0
                noise = garbage_hits(cur, rays, scan[0], scan[1])
                m = (tree[0] != scan[0]) | (hit & (bits(tree[1]) != bits(scan[1])))
                if (m & ~noise).any():
                    bad.append(("closest", builder, step, label, int((m & ~noise).sum())))
                if ((anyh != hit) & ~noise).any():
                    bad.append(("any-hit", builder, step, label))
                if label == "sun" and builder == 0:
                    got = s.sun_occluded(o)[0] > 0
                    wrong = (got != hit) & ~off_footprint(cur, o, scan[0])
                    if wrong.any():
                        bad.append(("sun-grid", step, int(wrong.sum())))
        s.close()
    return seed, KINDS[kind], len(tris), scale, bad


def _run(seed):
    return run_seed(seed)


def _run_refit(seed):
    return refit_seed(seed) + (0,)


def _run_render(seed):
    return render_seed(seed) + (0,)


if __name__ == "__main__":
    from multiprocessing import Pool
    args = [x for x in sys.argv[1:] if not x.startswith("--")]
    a = int(args[0]) if len(args) > 0 else 0
    b = int(args[1]) if len(args) > 1 else 400
    os.environ.setdefault("OMP_NUM_THREADS", "1")  # one process per core already: the emulation's own OpenMP loops would oversubscribe
    _emu()  # build once, before the workers start
    fails = docs = 0
    with Pool(os.cpu_count()) as p:
        for seed, kind, n, scale, bad, documented in p.imap_unordered(_run_render if "--render" in sys.argv else _run_refit if "--refit" in sys.argv else _run, range(a, b)):
            docs += documented
            if bad:
                fails += 1
                print(f"FAIL seed {seed} ({kind}, {n} triangles, scale {scale:.3g}): {bad}", flush=True)
    print(f"seeds {a}..{b - 1}: {fails} failing, {docs} differences on garbage hits (documented)")
    sys.exit(1 if fails else 0)
