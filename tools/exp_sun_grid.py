#!/usr/bin/env python
"""Would a 2-D grid in the sun's projection beat the tree for the shadow rays?  (development; CPU only, numpy)

All shadow rays are parallel to kLightDir.  Project the triangles onto the plane perpendicular to it, bin their (padded) 2-D
bounding boxes into an N x N grid, keep each cell's list sorted by the triangle's far depth along the sun direction, and walk the
list of the cell under a ray's origin: candidates whose depth range ends before the origin are skipped (the list is sorted: stop),
the rest are tested until one hits.  Prints list lengths and tests per ray for the golden Sponza shadow rays, beside the tree's
10.4 node visits + 2.5 triangle tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_rays, sponza_scene  # noqa: E402

tris = sponza_scene()[0].reshape(-1, 3, 3).astype(np.float64)
g = load_rays("sponza")
rays = g["rays"][g["kind"] == 2].astype(np.float64)
L = rays[0, 3:6] / np.linalg.norm(rays[0, 3:6])
assert np.allclose(rays[:, 3:6], rays[0, 3:6])
a = np.array([1.0, 0, 0]) if abs(L[0]) < 0.9 else np.array([0, 1.0, 0])
U = np.cross(L, a); U /= np.linalg.norm(U)
V = np.cross(L, U)
pu, pv, pw = tris @ U, tris @ V, tris @ L  # [n,3]
lo = np.array([pu.min(), pv.min()]); hi = np.array([pu.max(), pv.max()])
print(f"{len(tris)} triangles, {len(rays)} shadow rays, projected extent {hi - lo}")


def mt_hits(o, d, t3):  # plain double Moeller-Trumbore, any t > 1e-3
    e1, e2 = t3[:, 1] - t3[:, 0], t3[:, 2] - t3[:, 0]
    p = np.cross(d, e2); det = (e1 * p).sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / det
        s = o - t3[:, 0]; u = (s * p).sum(1) * inv
        q = np.cross(s, e1); v = (q @ d) * inv; t = (e2 * q).sum(1) * inv
    return (np.abs(det) > 1e-12) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 1e-3)


for N in (256, 512, 1024, 2048):
    cell = (hi - lo).max() / N
    x0 = np.clip(((pu.min(1) - lo[0]) / cell).astype(int), 0, N - 1); x1 = np.clip(((pu.max(1) - lo[0]) / cell).astype(int), 0, N - 1)
    y0 = np.clip(((pv.min(1) - lo[1]) / cell).astype(int), 0, N - 1); y1 = np.clip(((pv.max(1) - lo[1]) / cell).astype(int), 0, N - 1)
    entries = int(((x1 - x0 + 1) * (y1 - y0 + 1)).sum())
    zfar = pw.max(1)
    # per-ray candidate walk (only for the cells rays fall in)
    ro = rays[:, 0:3]
    cx = np.clip((((ro @ U) - lo[0]) / cell).astype(int), 0, N - 1); cy = np.clip((((ro @ V) - lo[1]) / cell).astype(int), 0, N - 1)
    rw = ro @ L
    rng = np.random.default_rng(1)
    pick = rng.choice(len(rays), 4000, replace=False)
    listlen, live, tested, hits = [], [], [], 0
    for i in pick:
        m = np.flatnonzero((x0 <= cx[i]) & (x1 >= cx[i]) & (y0 <= cy[i]) & (y1 >= cy[i]))
        listlen.append(len(m))
        m = m[zfar[m] > rw[i]]              # the sorted list stops here
        m = m[np.argsort(-zfar[m])]         # nearest to the sun first
        live.append(len(m))
        h = mt_hits(ro[i], L, tris[m]) if len(m) else np.zeros(0, bool)
        if h.any():
            hits += 1
            tested.append(int(np.argmax(h)) + 1)
        else:
            tested.append(len(m))
    listlen, live, tested = np.array(listlen), np.array(live), np.array(tested)
    print(f"N {N}: cell {cell:.4f}, {entries / 1e6:.1f} M list entries ({entries * 4 / 2**20:.0f} MiB), list per ray cell: mean {listlen.mean():.1f} p90 {np.percentile(listlen, 90):.0f}; "
          f"beyond the origin: mean {live.mean():.1f}; triangle tests until a hit / the end: mean {tested.mean():.2f} p50 {np.percentile(tested, 50):.0f} "
          f"p90 {np.percentile(tested, 90):.0f} p99 {np.percentile(tested, 99):.0f} max {tested.max()}; occluded {hits / len(pick):.2f}")

# ---- refinements: exact triangle-vs-cell overlap (2-D SAT) and a per-cell depth bound (the triangle's plane over the cell)
def tri_overlaps_cell(tu, tv, cx0, cy0, cx1, cy1):
    # 2-D separating axes: the square's axes (= bbox test, done by the caller) and the triangle's three edge normals
    ok = np.ones(len(tu), bool)
    for k in range(3):
        ax, ay = tu[:, k], tv[:, k]
        bx, by = tu[:, (k + 1) % 3], tv[:, (k + 1) % 3]
        ox, oy = tu[:, (k + 2) % 3], tv[:, (k + 2) % 3]
        nx, ny = -(by - ay), (bx - ax)
        side = nx * (ox - ax) + ny * (oy - ay)          # sign of the triangle's interior
        sgn = np.where(side >= 0, 1.0, -1.0)
        nx, ny = nx * sgn, ny * sgn                      # normal pointing inwards
        # the square's corner furthest along the inward normal must be inside this edge's half-plane
        px = np.where(nx >= 0, cx1, cx0); py = np.where(ny >= 0, cy1, cy0)
        ok &= nx * (px - ax) + ny * (py - ay) >= 0
    return ok


for N in (512, 1024, 2048):
    cell = (hi - lo).max() / N
    x0 = np.clip(((pu.min(1) - lo[0]) / cell).astype(int), 0, N - 1); x1 = np.clip(((pu.max(1) - lo[0]) / cell).astype(int), 0, N - 1)
    y0 = np.clip(((pv.min(1) - lo[1]) / cell).astype(int), 0, N - 1); y1 = np.clip(((pv.max(1) - lo[1]) / cell).astype(int), 0, N - 1)
    ro = rays[:, 0:3]
    cx = np.clip((((ro @ U) - lo[0]) / cell).astype(int), 0, N - 1); cy = np.clip((((ro @ V) - lo[1]) / cell).astype(int), 0, N - 1)
    rw = ro @ L
    zmin, zmax = pw.min(1), pw.max(1)
    rng = np.random.default_rng(1)
    pick = rng.choice(len(rays), 4000, replace=False)
    live, tested, lens = [], [], []
    for i in pick:
        m = np.flatnonzero((x0 <= cx[i]) & (x1 >= cx[i]) & (y0 <= cy[i]) & (y1 >= cy[i]))
        c0x, c0y = lo[0] + cx[i] * cell, lo[1] + cy[i] * cell
        m = m[tri_overlaps_cell(pu[m], pv[m], c0x, c0y, c0x + cell, c0y + cell)]
        lens.append(len(m))
        # depth of the triangle's plane over the cell: solve w = a u + b v + c from the three projected vertices
        A = np.stack([pu[m], pv[m], np.ones_like(pu[m])], -1)                      # [k,3,3]
        zf = zmax[m].copy()
        good = np.abs(np.linalg.det(A)) > 1e-12
        if good.any():
            coef = np.linalg.solve(A[good], pw[m][good][..., None])[..., 0]           # [k,3]
            corners = np.array([[c0x, c0y, 1], [c0x + cell, c0y, 1], [c0x, c0y + cell, 1], [c0x + cell, c0y + cell, 1]])
            zc = (coef @ corners.T).max(1)
            zf[good] = np.minimum(zmax[m][good], np.maximum(zc, zmin[m][good]))
        keep = zf > rw[i]
        m, zf = m[keep], zf[keep]
        order = np.argsort(-zf)
        m = m[order]
        live.append(len(m))
        h = mt_hits(ro[i], L, tris[m]) if len(m) else np.zeros(0, bool)
        tested.append(int(np.argmax(h)) + 1 if h.any() else len(m))
    live, tested, lens = np.array(live), np.array(tested), np.array(lens)
    print(f"N {N} exact overlap + per-cell depth: list mean {lens.mean():.1f}; beyond the origin: mean {live.mean():.1f}; tests: mean {tested.mean():.2f} p50 {np.percentile(tested, 50):.0f} "
          f"p90 {np.percentile(tested, 90):.0f} p99 {np.percentile(tested, 99):.0f} max {tested.max()}")
