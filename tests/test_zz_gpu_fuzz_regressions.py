"""The cases tools/fuzz_emu.py found on the CPU (host emulation), through the C ABI on the GPU: the CUDA tree walk, any-hit walk and
sun grid against the GPU's own all-triangle scan (TMPT_HIT_BRUTE: the reference's exact test on every triangle) and the oracle.

(Named to run last: these were written in round 2 after the GPU budget of the round was spent; the same logic is green on the CPU
through tests/test_emu_logic.py.)"""
import importlib.util
import os

import numpy as np
import pytest

import toymeshpathtracer_b200 as tm
from conftest import ROOT, bits, load_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fz():
    spec = importlib.util.spec_from_file_location("fuzz_emu", os.path.join(ROOT, "tools", "fuzz_emu.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _sun_rays(fz, o):
    return np.concatenate([o, np.broadcast_to(fz.light_dir(), o.shape)], 1).astype(np.float32)


@pytest.mark.parametrize("flags", [0, tm.BUILD_LBVH], ids=["sah", "lbvh"])
def test_rays_that_start_on_shared_vertices_with_tmin_zero(fz, oracle, flags):
    """tMin = 0, origins ON the vertices of a height-field mesh: all triangles around a vertex tie at t = +-0, the lowest original
    index wins (bvh.cuh walk_step: the pop-time cull masks the child slot out of the key)."""
    rng = np.random.default_rng(5)
    tris, _, _ = fz.make_scene(rng, kind=6)
    v = tris.reshape(-1, 3, 3)
    o = np.concatenate([v[:, 0], v[:, 1], v[:, 2], (v[:, 0] + v[:, 1]) * np.float32(0.5)]).astype(np.float32)
    d = rng.normal(size=o.shape)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    with tm.Scene(tris, flags=flags) as s:
        for rays in (_sun_rays(fz, o), np.concatenate([o, d], 1).astype(np.float32)):
            tree = s.HitScene(rays, tMin=0.0)
            scan = s.HitScene(rays, tMin=0.0, mode=tm.HIT_BRUTE)
            oid, ot, _, _ = oracle.hit_brute(tris, rays, tmin=0.0)
            hit = oid >= 0
            assert (hit & (ot == 0.0)).sum() > 50
            assert (scan[0] == oid).all() and (bits(scan[1])[hit] == bits(ot)[hit]).all()
            assert (tree[0] == oid).all() and (bits(tree[1])[hit] == bits(ot)[hit]).all()
            assert ((s.HitScene(rays, tMin=0.0, mode=tm.HIT_ANY)[0] == 1) == hit).all()


def test_sun_query_from_distant_origins(fz):
    """TMPT_HIT_SUN from 1 to a million scene sizes below the scene along the sun: beyond the far limit bvh::sun_query answers with
    the scan (the projection of such an origin is off by more than the grid's pads)."""
    l = fz.light_dir()
    for name in ("cube", "suzanne"):
        tris = load_scene(name)["tris"]
        v = tris.reshape(-1, 3, 3)
        rng = np.random.default_rng(3)
        pick = rng.integers(0, len(v), 20000)
        on = (v[pick] * rng.dirichlet([1, 1, 1], 20000).astype(np.float32)[:, :, None]).sum(1).astype(np.float32)
        ext = float(np.abs(v).max())
        with tm.Scene(tris) as s:
            for k in (1.0, 15.0, 17.0, 1e3, 1e6):
                rays = _sun_rays(fz, (on - l * np.float32(k * ext)).astype(np.float32))
                for tmax in (1.0e7, 3.0e38):
                    a = s.HitScene(rays, tMax=tmax, mode=tm.HIT_SUN)[0] >= 0
                    b = s.HitScene(rays, tMax=tmax, mode=tm.HIT_BRUTE)[0] >= 0
                    assert (a == b).all(), (name, k, tmax, int((a != b).sum()))


@pytest.mark.parametrize("seed", [0, 2, 7, 12, 15, 20, 37, 41])
def test_fuzz_scenes_tree_and_sun_grid_equal_the_scan(fz, seed):
    """The fuzzer's random scenes (nine kinds, scales 1e-4 .. 1e5, some far off the origin) on the GPU: tree closest / any hit
    and the sun grid against the all-triangle scan at tMin = 0.001 and 0.  Differences are allowed only where the scan's winner is
    a GARBAGE hit (rounding noise that passed the determinant test, at a point outside the triangle's padded box,
    or -- for the grid -- from an origin outside the triangle's footprint in the sun's projection)."""
    rng = np.random.default_rng(seed)
    tris, scale, kind = fz.make_scene(rng)
    o = fz.make_origins(rng, tris, scale)
    d = rng.normal(size=o.shape)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    for flags in (0, tm.BUILD_LBVH):
        with tm.Scene(tris, flags=flags) as s:
            for label, rays in (("sun", _sun_rays(fz, o)), ("random", np.concatenate([o, d], 1).astype(np.float32))):
                for tmin in (0.001, 0.0):
                    what = (fz.KINDS[kind], label, tmin)
                    scan = s.HitScene(rays, tMin=tmin, mode=tm.HIT_BRUTE)
                    tree = s.HitScene(rays, tMin=tmin)
                    hit = scan[0] >= 0
                    noise = fz.garbage_hits(tris, rays, scan[0], scan[1])
                    anyh = s.HitScene(rays, tMin=tmin, mode=tm.HIT_ANY)[0] == 1
                    assert not ((anyh != hit) & ~noise).any(), what
                    if label == "sun" and flags == 0:
                        sun = s.HitScene(rays, tMin=tmin, mode=tm.HIT_SUN)[0] == 1
                        assert not ((sun != hit) & ~fz.off_footprint(tris, o, scan[0])).any(), what
                    m = (tree[0] != scan[0]) | (hit & (bits(tree[1]) != bits(scan[1])))
                    assert not (m & ~noise).any(), what + (int((m & ~noise).sum()),)


@pytest.mark.parametrize("name,w,h,spp", [("suzanne", 64, 36, 2), ("cube", 96, 54, 2), ("teapot", 48, 27, 1)])
def test_reference_program_with_the_scene_class_swapped(tmp_path, name, w, h, spp):
    """The drop-in boundary at its narrowest (oracle/dropin/scene_tmpt.cpp): the reference's OWN program -- its
    main.cpp, loader, camera, Trace / Scatter, per-row RNG streams, TBB row loop, stb PNG writer, and its unmodified scene.h --
    with only scene.cpp replaced, so that every Scene::HitScene is answered by libtmpt.so on the GPU (one ray per call).  Because
    flag, Hit.pos and Hit.normal come back bit for bit, the reference's integrator walks the same paths: output.png is the
    reference binary's output.png byte for byte, and the ray count is the same.  (Both binaries are built in the build container,
    oracle/Makefile, and travel in oracle/_ref/.)"""
    import subprocess
    import bench
    from oracle.pyoracle import DROPIN_BIN, REF_BIN
    if not (os.path.exists(DROPIN_BIN) and os.path.exists(REF_BIN)):
        pytest.skip("oracle/_ref binaries not built (they are built where the reference sources are)")
    obj = bench.scene_obj_path(name)
    out = {}
    for label, exe in (("reference", REF_BIN), ("dropin", DROPIN_BIN)):
        d = tmp_path / label
        d.mkdir()
        # (two worker threads for the patched program: concurrent callers of one scene take turns inside tmpt_hit_scene(TMPT_HOST);
        #  the image does not depend on the thread count -- every row seeds its own stream, main.cpp:204)
        env = dict(os.environ, TBB_SHIM_THREADS="2") if label == "dropin" else None
        r = subprocess.run([exe, str(w), str(h), str(spp), obj], cwd=d, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, (label, r.stdout, r.stderr)
        lines = r.stdout.strip().split("\n")
        assert lines[1].startswith(f"Rendered scene at {w}x{h},{spp}spp in ")
        out[label] = ((d / "output.png").read_bytes(), lines[2].split()[1])  # the PNG, and "- <n> K Rays"
    assert out["dropin"][1] == out["reference"][1]
    assert out["dropin"][0] == out["reference"][0]
