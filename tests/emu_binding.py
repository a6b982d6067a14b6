"""ctypes binding of tests/emu/libemu.so -- TEST INFRASTRUCTURE (see tests/emu/emu.cpp).

The product's __host__ __device__ headers compiled for the host, so the CPU suite can check
the BVH build, traversal and integrator logic without a GPU.  Never used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu.cpp")
SO = os.path.join(HERE, "emu", "libemu.so")
CSRC = os.path.join(os.path.dirname(HERE), "toymeshpathtracer_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")


def _stale(so=SO):
    if not os.path.exists(so):
        return True
    t = os.path.getmtime(so)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(defines=(), tag=""):
    """`defines` / `tag`: a second build of the same logic with compile-time switches (e.g. -DTMPT_QNODES=1 -> libemu_q.so)."""
    so = SO if not tag else SO.replace("libemu.so", f"libemu_{tag}.so")
    if _stale(so):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", *defines,
                        "-I" + CUDA_INC, "-o", so, SRC], check=True)
    return so


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emu:
    def __init__(self, defines=(), tag=""):
        self.L = C.CDLL(build(defines, tag))
        self.L.emu_scene_create.restype = C.c_void_p
        self.L.emu_scene_create2.restype = C.c_void_p
        self.L.emu_pixel_seed.restype = C.c_uint32
        self.L.emu_chunk_seed.restype = C.c_uint32

    def scene(self, tris, builder=0, c_inner=1.0, c_tri=1.0, max_leaf=8):
        """builder 0 = binned SAH (the default of the product), 1 = LBVH"""
        return EmuScene(self.L, tris, builder, c_inner, c_tri, max_leaf)


class EmuScene:
    def __init__(self, L, tris, builder=0, c_inner=1.0, c_tri=1.0, max_leaf=8):
        self.L = L
        self.tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        self.h = C.c_void_p(L.emu_scene_create2(_p(self.tris), self.tris.shape[0], builder, C.c_float(c_inner), C.c_float(c_tri), max_leaf))

    def close(self):
        if self.h:
            self.L.emu_scene_destroy(self.h)
            self.h = None

    __del__ = close

    def refit(self, tris):
        self.tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        self.L.emu_scene_refit(self.h, _p(self.tris))

    def info(self):
        out = np.zeros(5, np.uint32)
        self.L.emu_scene_info(self.h, _p(out))
        return dict(zip(("nodes", "slots", "leaves", "max_depth", "status"), out.tolist()))

    def nodes(self):
        n = self.L.emu_nodes(self.h, None)
        out = np.zeros((n, 7, 4), np.float32)
        self.L.emu_nodes(self.h, _p(out))
        return out

    def qnodes(self):
        """Quantised nodes in logical row order: uint32 [n, 4 rows, 4 words]."""
        n = self.L.emu_nodes(self.h, None)
        out = np.zeros((n, 4, 4), np.uint32)
        self.L.emu_qnodes(self.h, _p(out))
        return out

    def slots(self):
        out = np.zeros((self.tris.shape[0], 3, 4), np.float32)
        self.L.emu_slots(self.h, _p(out))
        return out

    def hit(self, rays, tmin=0.001, tmax=1.0e7, mode=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        ids = np.full(n, -1, np.int32)
        t = np.zeros(n, np.float32); pos = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
        self.L.emu_hit_scene(self.h, _p(rays), C.c_long(n), C.c_float(tmin), C.c_float(tmax), mode, _p(ids), _p(t), _p(pos), _p(nrm))
        return ids, t, pos, nrm

    def sun_grid(self, cells=-1):
        """Rebuild the shadow rays' grid (csrc/sungrid.cuh) with `cells` per side (-1: default, 0: none) -> info dict."""
        self.L.emu_sun_grid(self.h, int(cells))
        out = np.zeros(3, np.int64)
        self.L.emu_sun_info(self.h, _p(out))
        return {"n": int(out[0]), "entries": int(out[1]), "longest": int(out[2])}

    def sun_occluded(self, origins, tmin=0.001, tmax=1.0e7):
        """The integrator's shadow query through the grid: (1 / -1 per origin, exact triangle tests run)."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        out = np.zeros(o.shape[0], np.int32)
        tests = C.c_ulonglong(0)
        self.L.emu_sun_occluded(self.h, _p(o), C.c_long(o.shape[0]), C.c_float(tmin), C.c_float(tmax), _p(out), C.byref(tests))
        return out, tests.value

    def render_stats(self, cam22, w, h, spp):
        cam22 = np.ascontiguousarray(cam22, np.float32)
        out = np.zeros(3, np.uint64)
        self.L.emu_render_stats(self.h, _p(cam22), w, h, spp, _p(out))
        rays = max(int(out[0]), 1)
        return {"rays": int(out[0]), "nodes_per_ray": int(out[1]) / rays, "tris_per_ray": int(out[2]) / rays}

    def render(self, cam22, w, h, spp, rows=None):
        cam22 = np.ascontiguousarray(cam22, np.float32)
        img = np.zeros((h, w, 4), np.uint8)
        rc = C.c_longlong(0)
        r0, r1 = rows if rows else (0, h)
        self.L.emu_render(self.h, _p(cam22), w, h, spp, r0, r1, _p(img), None, C.byref(rc))
        return img, rc.value
