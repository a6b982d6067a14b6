"""ctypes wrappers for the CHECKERS -- test infrastructure, never the product path.

``Oracle``  -> oracle/liboracle.so   (plain-C restatement, oracle/oracle.c)
``Ref``     -> oracle/_ref/libref.so (the unmodified reference's own functions)
``build_dropin`` -> oracle/_ref/TrimeshTracer_dropin (the reference program with `struct Scene` implemented over the C ABI)

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``toymeshpathtracer_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")
REF_BIN = os.path.join(HERE, "_ref", "TrimeshTracer")
DROPIN_BIN = os.path.join(HERE, "_ref", "TrimeshTracer_dropin")
REF_ROOT = os.environ.get("TMPT_REF", "/root/reference")

RNG_ROW, RNG_PIXEL = 0, 1
TRIG_LIBM, TRIG_SPEC = 0, 1


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_oracle() -> str:
    """Compile oracle.c (seconds).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return ORACLE_SO


def build_ref() -> str | None:
    """Compile the reference from its own sources if they are present (this container only)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "source")):
        return REF_SO if os.path.exists(REF_SO) else None
    subprocess.run(["make", "-s", "-C", HERE, "ref", f"REF={REF_ROOT}"], check=True)
    return REF_SO


def build_dropin() -> str | None:
    """The reference program with its scene.cpp swapped for oracle/dropin/scene_tmpt.cpp (struct Scene over the C ABI), linked
    against the product's libtmpt.so -- needs the reference sources (this container only) and a built product library."""
    lib = os.path.join(os.path.dirname(HERE), "toymeshpathtracer_b200", "libtmpt.so")
    if not os.path.isdir(os.path.join(REF_ROOT, "source")) or not os.path.exists(lib):
        return DROPIN_BIN if os.path.exists(DROPIN_BIN) else None
    subprocess.run(["make", "-s", "-C", HERE, "dropin", f"REF={REF_ROOT}"], check=True)
    return DROPIN_BIN


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "oracle.c")):
            build_oracle()
        L = self.L = C.CDLL(ORACLE_SO)
        L.orc_random_in_unit_disk.restype = C.c_uint32
        L.orc_random_unit_vectors.restype = C.c_uint32
        L.orc_camera_get_rays.restype = C.c_uint32
        L.orc_pixel_seed.restype = C.c_uint32
        L.orc_stream_seed.restype = C.c_uint32
        L.orc_add_floor.restype = C.c_int
        self.threads = int(os.environ.get("ORACLE_THREADS", os.cpu_count() or 1))

    def rng_states(self, seed, n):
        st = np.zeros(n, np.uint32); fl = np.zeros(n, np.float32)
        self.L.orc_rng_states(C.c_uint32(seed), n, _p(st), _p(fl))
        return st, fl

    def random_in_unit_disk(self, seed, n):
        out = np.zeros((n, 3), np.float32)
        s = self.L.orc_random_in_unit_disk(C.c_uint32(seed), n, _p(out))
        return out, s

    def random_unit_vectors(self, seed, n, trig=TRIG_LIBM):
        out = np.zeros((n, 3), np.float32)
        s = self.L.orc_random_unit_vectors(C.c_uint32(seed), n, _p(out), trig)
        return out, s

    def sincos_spec(self, a):
        a = _f32(np.atleast_1d(a))
        s = np.zeros_like(a); c = np.zeros_like(a)
        fs, fc = C.c_float(), C.c_float()
        for i, v in enumerate(a):
            self.L.orc_sincos_spec(C.c_float(float(v)), C.byref(fs), C.byref(fc))
            s[i], c[i] = fs.value, fc.value
        return s, c

    def pixel_seed(self, idx):
        return self.L.orc_pixel_seed(C.c_uint32(idx))

    def stream_seed(self, idx64):
        return self.L.orc_stream_seed(C.c_uint64(idx64))

    def camera_make(self, frm, at, up, vfov, aspect, aperture, focus):
        out = np.zeros(22, np.float32)
        self.L.orc_camera_make(_p(_f32(frm)), _p(_f32(at)), _p(_f32(up)), C.c_float(vfov), C.c_float(aspect),
                               C.c_float(aperture), C.c_float(focus), _p(out))
        return out

    def camera_for_scene(self, mn, mx, w, h, is_sponza=False):
        out = np.zeros(22, np.float32)
        self.L.orc_camera_for_scene(_p(_f32(mn)), _p(_f32(mx)), int(is_sponza), w, h, _p(out))
        return out

    def camera_get_rays(self, cam22, st2, seed):
        st2 = _f32(st2); n = st2.shape[0]
        out = np.zeros((n, 6), np.float32)
        s = self.L.orc_camera_get_rays(_p(_f32(cam22)), _p(st2), n, C.c_uint32(seed), _p(out))
        return out, s

    def add_floor(self, model9):
        model9 = _f32(model9).reshape(-1, 9)
        n = model9.shape[0]
        out = np.zeros((n + 2, 9), np.float32)
        mn = np.zeros(3, np.float32); mx = np.zeros(3, np.float32)
        self.L.orc_add_floor(_p(model9), n, _p(out), _p(mn), _p(mx))
        return out, mn, mx

    def light_dir(self):
        out = np.zeros(3, np.float32)
        self.L.orc_light_dir(_p(out))
        return out

    def hit_brute(self, tris9, rays6, tmin=0.001, tmax=1.0e7, any_hit=False, threads=None):
        tris9 = _f32(tris9).reshape(-1, 9); rays6 = _f32(rays6).reshape(-1, 6)
        n = rays6.shape[0]
        ids = np.full(n, -1, np.int32)
        t = np.zeros(n, np.float32); pos = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
        self.L.orc_hit_brute(_p(tris9), tris9.shape[0], _p(rays6), C.c_long(n), C.c_float(tmin), C.c_float(tmax),
                             int(any_hit), _p(ids), _p(t), _p(pos), _p(nrm), threads or self.threads)
        return ids, t, pos, nrm

    def render(self, tris9, cam22, w, h, spp, rng=RNG_PIXEL, trig=TRIG_SPEC, rows=None, linear=False, threads=None):
        tris9 = _f32(tris9).reshape(-1, 9)
        rgba = np.zeros((h, w, 4), np.uint8)
        lin = np.zeros((h, w, 3), np.float32) if linear else None
        rc = C.c_int64(0)
        r0, r1 = rows if rows else (0, h)
        self.L.orc_render(_p(tris9), tris9.shape[0], _p(_f32(cam22)), w, h, spp, rng, trig, r0, r1, _p(rgba), _p(lin),
                          C.byref(rc), threads or self.threads)
        return (rgba, rc.value, lin) if linear else (rgba, rc.value)


class Ref:
    """The unmodified reference behind a C ABI (oracle/ref_harness.cpp)."""

    def __init__(self):
        so = build_ref()
        if so is None or not os.path.exists(so):
            raise FileNotFoundError("oracle/_ref/libref.so not built and reference sources absent")
        L = self.L = C.CDLL(so)
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_from_tris.restype = C.c_void_p
        L.ref_scene_triangles.restype = C.c_int
        L.ref_random_in_unit_disk.restype = C.c_uint32
        L.ref_random_unit_vectors.restype = C.c_uint32
        L.ref_camera_get_rays.restype = C.c_uint32
        L.ref_record_path_rays.restype = C.c_long
        L.ref_threads.restype = C.c_int

    def scene_load(self, path):
        mn = np.zeros(3, np.float32); mx = np.zeros(3, np.float32); n = C.c_int(0)
        h = self.L.ref_scene_load(path.encode(), _p(mn), _p(mx), C.byref(n))
        if not h:
            raise IOError(f"reference LoadScene failed for {path}")
        tris = np.zeros((n.value, 9), np.float32)
        self.L.ref_scene_triangles(C.c_void_p(h), _p(tris))
        return C.c_void_p(h), tris, mn, mx

    def scene_from_tris(self, tris9, mn, mx):
        tris9 = _f32(tris9).reshape(-1, 9)
        return C.c_void_p(self.L.ref_scene_from_tris(_p(tris9), tris9.shape[0], _p(_f32(mn)), _p(_f32(mx))))

    def scene_free(self, h):
        self.L.ref_scene_free(h)

    def _hit(self, fn, h, rays6, tmin, tmax):
        rays6 = _f32(rays6).reshape(-1, 6); n = rays6.shape[0]
        ids = np.full(n, -1, np.int32)
        t = np.zeros(n, np.float32); pos = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
        fn(h, _p(rays6), C.c_long(n), C.c_float(tmin), C.c_float(tmax), _p(ids), _p(t), _p(pos), _p(nrm))
        return ids, t, pos, nrm

    def hit_scene(self, h, rays6, tmin=0.001, tmax=1.0e7):
        return self._hit(self.L.ref_hit_scene, h, rays6, tmin, tmax)

    def hit_brute(self, h, rays6, tmin=0.001, tmax=1.0e7):
        return self._hit(self.L.ref_hit_brute, h, rays6, tmin, tmax)

    def rng_states(self, seed, n):
        st = np.zeros(n, np.uint32); fl = np.zeros(n, np.float32)
        self.L.ref_rng_states(C.c_uint32(seed), n, _p(st), _p(fl))
        return st, fl

    def random_in_unit_disk(self, seed, n):
        out = np.zeros((n, 3), np.float32)
        s = self.L.ref_random_in_unit_disk(C.c_uint32(seed), n, _p(out))
        return out, s

    def random_unit_vectors(self, seed, n):
        out = np.zeros((n, 3), np.float32)
        s = self.L.ref_random_unit_vectors(C.c_uint32(seed), n, _p(out))
        return out, s

    def camera_make(self, frm, at, up, vfov, aspect, aperture, focus):
        out = np.zeros(22, np.float32)
        self.L.ref_camera_make(_p(_f32(frm)), _p(_f32(at)), _p(_f32(up)), C.c_float(vfov), C.c_float(aspect),
                               C.c_float(aperture), C.c_float(focus), _p(out))
        return out

    def camera_for_scene(self, h, path, w, hgt):
        out = np.zeros(22, np.float32)
        self.L.ref_camera_for_scene(h, path.encode(), w, hgt, _p(out))
        return out

    def camera_get_rays(self, cam22, st2, seed):
        st2 = _f32(st2); n = st2.shape[0]
        out = np.zeros((n, 6), np.float32)
        s = self.L.ref_camera_get_rays(_p(_f32(cam22)), _p(st2), n, C.c_uint32(seed), _p(out))
        return out, s

    def record_path_rays(self, h, cam22, w, hgt, stride, max_rays):
        rays = np.zeros((max_rays, 6), np.float32); kind = np.zeros(max_rays, np.int32)
        n = self.L.ref_record_path_rays(h, _p(_f32(cam22)), w, hgt, stride, C.c_long(max_rays), _p(rays), _p(kind))
        return rays[:n].copy(), kind[:n].copy()

    def render(self, h, cam22, w, hgt, spp):
        rgba = np.zeros((hgt, w, 4), np.uint8); rc = C.c_longlong(0)
        self.L.ref_render(h, _p(_f32(cam22)), w, hgt, spp, _p(rgba), C.byref(rc))
        return rgba, rc.value

    def light_dir(self):
        out = np.zeros(3, np.float32)
        self.L.ref_light_dir(_p(out))
        return out


def ref_available() -> bool:
    return os.path.exists(REF_SO) or os.path.isdir(os.path.join(REF_ROOT, "source"))
