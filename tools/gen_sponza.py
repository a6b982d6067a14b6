#!/usr/bin/env python
"""Deterministic stand-in for the reference's missing data/sponza.obj.

/root/reference/.MISSING_LARGE_BLOBS lists data/sponza.obj as absent, so BASELINE.json's
Sponza configurations run on this procedural atrium instead: a closed two-storey arcade hall
with the nave open to the sky, EXACTLY 66 450 triangles (LoadScene adds the two floor triangles ->
66 452, the count readme.md:74 quotes), bounds x in [-18,18], y in [-0.2,15], z in [-8,8], so
that the hard-coded Sponza eye (-5.96, 4.08, -1.22) of main.cpp:300-301 stands in the nave.
The file name must contain "sponza.obj" for that camera to trigger.  No randomness: the
same bytes every time.  Every results line that uses it says "sponza stand-in".

    python tools/gen_sponza.py [out.obj]
"""
from __future__ import annotations

import os
import sys

import numpy as np

TARGET_TRIS = 66450


class Mesh:
    def __init__(self):
        self.v = []   # list of (n,3) arrays
        self.f = []   # list of (m,3) int arrays, 0-based global
        self.nv = 0

    def add(self, verts, faces):
        verts = np.asarray(verts, np.float64).reshape(-1, 3)
        faces = np.asarray(faces, np.int64).reshape(-1, 3)
        self.v.append(verts)
        self.f.append(faces + self.nv)
        self.nv += verts.shape[0]

    @property
    def tri_count(self):
        return sum(f.shape[0] for f in self.f)

    def grid(self, fn, nu, nv, flip=False):
        """Tessellate the parametric patch fn(u, v) -> xyz, u,v in [0,1], into 2*nu*nv triangles."""
        u, v = np.meshgrid(np.linspace(0.0, 1.0, nu + 1), np.linspace(0.0, 1.0, nv + 1), indexing="ij")
        p = fn(u.ravel(), v.ravel())
        idx = np.arange((nu + 1) * (nv + 1)).reshape(nu + 1, nv + 1)
        a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
        tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)
        if flip:
            tris = tris[:, ::-1]
        self.add(p, tris)


def quad(p0, du, dv):
    p0, du, dv = (np.asarray(x, np.float64) for x in (p0, du, dv))
    return lambda u, v: p0[None, :] + u[:, None] * du[None, :] + v[:, None] * dv[None, :]


def build() -> Mesh:
    m = Mesh()
    X0, X1, Y0, Y1, Z0, Z1 = -18.0, 18.0, -0.2, 15.0, -8.0, 8.0
    # floor of the hall (y = 0; the plinth below reaches y = -0.2 so LoadScene's floor hides under it)
    m.grid(quad([X0, 0.0, Z0], [0, 0, Z1 - Z0], [X1 - X0, 0, 0]), 32, 72)                       # 4608
    # outer walls, facing inwards
    m.grid(quad([X0, Y0, Z0], [X1 - X0, 0, 0], [0, Y1 - Y0, 0]), 72, 30)                         # 4320  z = -8
    m.grid(quad([X0, Y0, Z1], [0, Y1 - Y0, 0], [X1 - X0, 0, 0]), 30, 72)                         # 4320  z = +8
    m.grid(quad([X0, Y0, Z0], [0, Y1 - Y0, 0], [0, 0, Z1 - Z0]), 30, 32)                         # 1920  x = -18
    m.grid(quad([X1, Y0, Z0], [0, 0, Z1 - Z0], [0, Y1 - Y0, 0]), 32, 30)                         # 1920  x = +18
    # roof over the galleries only: the nave (z in [-5, 5]) is open to the sky like the real court
    m.grid(quad([X0, Y1, Z0], [X1 - X0, 0, 0], [0, 0, 3.0]), 72, 6)                              # 864
    m.grid(quad([X0, Y1, 5.0], [X1 - X0, 0, 0], [0, 0, 3.0]), 72, 6)                             # 864
    # upper gallery slabs (y = 7 .. 7.3) along both long sides
    for s in (-1.0, 1.0):
        zo, zi = 8.0 * s, 5.0 * s
        m.grid(quad([X0, 7.3, zo], [X1 - X0, 0, 0], [0, 0, zi - zo]), 72, 6, flip=s > 0)         # 864 top
        m.grid(quad([X0, 7.0, zo], [0, 0, zi - zo], [X1 - X0, 0, 0]), 6, 72, flip=s > 0)         # 864 bottom
        m.grid(quad([X0, 7.0, zi], [X1 - X0, 0, 0], [0, 0.3, 0]), 72, 1, flip=s < 0)             # 144 edge
    # columns + capitals, two storeys, both sides
    cols_x = np.arange(-16.5, 16.6, 3.0)                                                          # 12
    for s in (-1.0, 1.0):
        for (yb, yt) in ((0.0, 5.2), (7.3, 12.0)):
            for cx in cols_x:
                cz, r = 5.0 * s, 0.32

                def shaft(u, v, cx=cx, cz=cz, yb=yb, yt=yt, r=r):
                    a = 2.0 * np.pi * u
                    rr = r * (1.0 - 0.12 * v)
                    return np.stack([cx + rr * np.cos(a), yb + (yt - yb) * v, cz + rr * np.sin(a)], 1)

                def capital(u, v, cx=cx, cz=cz, yt=yt, r=r):
                    a = 2.0 * np.pi * u
                    rr = r * (0.88 + 0.9 * v * v)
                    return np.stack([cx + rr * np.cos(a), yt + 0.45 * v, cz + rr * np.sin(a)], 1)

                m.grid(shaft, 16, 6)                                                               # 192
                m.grid(capital, 16, 2)                                                             # 64
    # arches between neighbouring columns: half tori
    for s in (-1.0, 1.0):
        for ytop in (5.65, 12.45):
            for k in range(len(cols_x) - 1):
                xm, cz = 0.5 * (cols_x[k] + cols_x[k + 1]), 5.0 * s

                def arch(u, v, xm=xm, cz=cz, ytop=ytop):
                    a = np.pi * u            # along the arch
                    b = 2.0 * np.pi * v      # around the tube
                    R, t = 1.5, 0.22
                    rad = R + t * np.cos(b)
                    return np.stack([xm - rad * np.cos(a), ytop + 0.9 * rad * np.sin(a) * 0.85, cz + 1.6 * t * np.sin(b)], 1)

                m.grid(arch, 16, 10)                                                               # 320
    # hanging drapes across the nave (wavy cloth)
    for k, x in enumerate((-12.0, -6.5, 1.0, 7.5, 13.0)):

        def drape(u, v, x=x, k=k):
            z = -3.6 + 7.2 * u
            y = 13.5 - 4.0 * v - 1.1 * np.sin(np.pi * u) * (0.4 + 0.6 * v)
            xx = x + 0.35 * np.sin(6.0 * np.pi * u + k) * (0.3 + v) + 0.15 * np.sin(9.0 * v + 2.0 * k)
            return np.stack([xx, y, z], 1)

        m.grid(drape, 40, 30)                                                                      # 2400
    # planters: spheres on the floor of the nave
    for k, (x, z) in enumerate(((-14.0, -2.5), (-9.0, 2.6), (-3.0, -2.8), (3.5, 2.4), (9.5, -2.2), (14.5, 2.9), (-16.0, 3.0), (16.0, -3.0))):

        def ball(u, v, x=x, z=z, k=k):
            a, b = 2.0 * np.pi * u, np.pi * v
            r = 0.7 + 0.05 * k
            return np.stack([x + r * np.sin(b) * np.cos(a), r + r * np.cos(b) * -1.0, z + r * np.sin(b) * np.sin(a)], 1)

        m.grid(ball, 16, 12)                                                                       # 384
    # fill to the exact count with small open pyramids (4 triangles) on the gallery floors, then single flags
    remaining = TARGET_TRIS - m.tri_count
    assert remaining >= 0, remaining
    k = 0
    while remaining >= 4:
        side = -1.0 if (k & 1) else 1.0
        i = k >> 1
        x = -17.0 + 0.21 * (i % 160) + 0.05
        z = side * (5.6 + 0.5 * (i // 160))
        y, h, r = 7.3, 0.18 + 0.02 * (i % 5), 0.08
        base = np.array([[x - r, y, z - r], [x + r, y, z - r], [x + r, y, z + r], [x - r, y, z + r], [x, y + h, z]])
        m.add(base, [[0, 4, 1], [1, 4, 2], [2, 4, 3], [3, 4, 0]])
        remaining -= 4
        k += 1
    for j in range(remaining):
        x = -10.0 + 5.0 * j
        m.add([[x, 9.0, -7.9], [x + 0.8, 9.0, -7.9], [x + 0.4, 9.9, -7.9]], [[0, 1, 2]])
    assert m.tri_count == TARGET_TRIS, m.tri_count
    return m


def write_obj(path: str) -> int:
    m = build()
    v = np.concatenate(m.v, 0).astype(np.float32)  # the floats the parser will reproduce
    f = np.concatenate(m.f, 0) + 1
    lines = ["# procedural stand-in for the Crytek Sponza atrium (tools/gen_sponza.py); %d triangles\n" % f.shape[0]]
    lines += ["v %.9g %.9g %.9g\n" % (float(a), float(b), float(c)) for a, b, c in v]
    lines += ["f %d %d %d\n" % (a, b, c) for a, b, c in f]
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "w") as fh:
        fh.writelines(lines)
    os.replace(tmp, path)
    return f.shape[0]


def triangles() -> np.ndarray:
    """(66450, 9) float32 triangle array WITHOUT the floor triangles (what an OBJ parse yields)."""
    m = build()
    v = np.concatenate(m.v, 0).astype(np.float32)
    f = np.concatenate(m.f, 0)
    return v[f].reshape(-1, 9)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "sponza.obj"
    n = write_obj(out)
    print(f"wrote {out}: {n} triangles")
