"""Parity tests proper: the CUDA path, called through the C ABI (libtmpt.so), against the
golden vectors the reference produced and against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star):
  * hit flag, ORIGINAL triangle id, t bits and the Hit payload bits: bit-exact;
  * a frame rendered with the GPU path's RNG layout (one XorShift32 stream per pixel, trig
    spec): bit-exact against the oracle run in the same mode;
  * against the reference's OWN image (one stream per row, libm trig) the comparison is
    statistical at high spp: per-pixel mean absolute error and PSNR, tolerances below.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import toymeshpathtracer_b200 as tm
from conftest import GOLD, ROOT, bits, load_rays, load_scene, sponza_scene

pytestmark = pytest.mark.gpu
SCENES = ["triangle", "cube", "suzanne", "teapot"]


@pytest.fixture(scope="module")
def scenes():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = tm.Scene(sponza_scene()[0] if name == "sponza" else load_scene(name)["tris"])
        return cache[name]
    yield get
    for s in cache.values():
        s.close()


def test_native_library_is_what_runs():
    assert tm.device_count() >= 1
    before = tm.launch_count()
    with tm.Scene(load_scene("cube")["tris"]) as s:
        s.HitScene(load_rays("cube")["rays"][:100])
    assert tm.launch_count() > before


@pytest.mark.parametrize("name", SCENES + ["sponza"])
def test_hit_ids_t_and_payload_bit_exact(scenes, name):
    """Config 2 of BASELINE.json (suzanne), the other reference scenes and the headline scene (the Sponza stand-in, 165 505
    rays recorded from a reference render by oracle/gen_golden.py --sponza): every ray a reference render shoots.  Flag and t
    are the reference octree's own; id and payload come from the ID-carrying scan over the reference's triangle test (on 6
    Sponza rays the octree kept another triangle with bit-equal t: the stated tie rule is the lowest input index)."""
    g = load_rays(name)
    s = scenes(name)
    ids, t, pos, nrm = s.HitScene(g["rays"])
    hit = g["id"] >= 0
    assert ((g["flag"] == 1) == (ids >= 0)).all()
    assert (ids == g["id"]).all()
    assert (bits(t)[hit] == bits(g["t"])[hit]).all()
    assert (bits(pos)[hit] == bits(g["pos"])[hit]).all()
    assert (bits(nrm)[hit] == bits(g["normal"])[hit]).all()
    # outputs of misses stay untouched (scene.cpp:86-97 writes outHit only on a hit)
    assert (t[~hit] == 0).all() and (pos[~hit] == 0).all()
    # any-hit (shadow semantics): same boolean
    aid, *_ = s.HitScene(g["rays"], mode=tm.HIT_ANY)
    assert ((aid == 1) == hit).all() and ((aid == -1) == ~hit).all()
    # brute force on the GPU (upstream's scan) agrees too
    bid, bt, *_ = s.HitScene(g["rays"], mode=tm.HIT_BRUTE)
    assert (bid == g["id"]).all() and (bits(bt)[hit] == bits(g["t"])[hit]).all()


@pytest.mark.parametrize("name", ["suzanne", "teapot", "sponza"])
def test_lbvh_builder_is_exact_too(name):
    """TMPT_BUILD_LBVH (Morton + radix sort + Karras) gives the same answers as the default binned-SAH tree."""
    g = load_rays(name)
    with tm.Scene(sponza_scene()[0] if name == "sponza" else load_scene(name)["tris"], flags=tm.BUILD_LBVH) as s:
        assert s.info()["builder"] == tm.BUILD_LBVH
        ids, t, pos, nrm = s.HitScene(g["rays"])
    hit = g["id"] >= 0
    assert (ids == g["id"]).all() and (bits(t)[hit] == bits(g["t"])[hit]).all()
    assert (bits(pos)[hit] == bits(g["pos"])[hit]).all() and (bits(nrm)[hit] == bits(g["normal"])[hit]).all()


def _random_rays(sc, n, seed, axis_parallel=True):
    rng = np.random.default_rng(seed)
    ext = (sc["bounds_max"] - sc["bounds_min"]) * 0.75
    o = rng.uniform(sc["bounds_min"] - ext, sc["bounds_max"] + ext, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    if axis_parallel:
        d[: n // 16, 0] = 0.0
        d[n // 16: n // 8, 1] = 0.0
        d[n // 8: n // 8 + n // 16, 2] = 0.0
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return np.concatenate([o, d], 1)


@pytest.mark.parametrize("name,n", [("cube", 200000), ("suzanne", 200000), ("teapot", 60000)])
def test_random_rays_vs_oracle(scenes, oracle, name, n):
    sc = load_scene(name)
    rays = _random_rays(sc, n, 5)
    ids, t, pos, nrm = scenes(name).HitScene(rays)
    oid, ot, opos, onrm = oracle.hit_brute(sc["tris"], rays)
    hit = oid >= 0
    assert (ids == oid).all()
    assert (bits(t)[hit] == bits(ot)[hit]).all() and (bits(pos)[hit] == bits(opos)[hit]).all() and (bits(nrm)[hit] == bits(onrm)[hit]).all()


@pytest.mark.parametrize("tmin,tmax", [(0.001, 1.0e7), (0.5, 3.0), (0.0, 1.0), (2.0, 2.0)])
def test_t_range_rules(scenes, oracle, tmin, tmax):
    """tMin <= t <= tMax and t < tMax (maths.cpp:371, scene.cpp:34, 90)."""
    sc = load_scene("suzanne")
    rays = _random_rays(sc, 50000, 9)
    ids, t, *_ = scenes("suzanne").HitScene(rays, tMin=tmin, tMax=tmax)
    oid, ot, *_ = oracle.hit_brute(sc["tris"], rays, tmin=tmin, tmax=tmax)
    assert (ids == oid).all() and (bits(t)[oid >= 0] == bits(ot)[oid >= 0]).all()
    assert ((t[ids >= 0] >= tmin) & (t[ids >= 0] < tmax)).all()


def test_tree_vs_brute_force_at_scale(scenes):
    """Size-independent property: the BVH answer equals the all-triangle scan on the GPU, at a
    ray count the CPU oracle cannot reach (conservativeness of the padded boxes)."""
    sc = load_scene("teapot")
    s = scenes("teapot")
    for seed in (1, 2):
        rays = _random_rays(sc, 2_000_000, seed)
        a = s.HitScene(rays, payload=True)
        b = s.HitScene(rays, mode=tm.HIT_BRUTE, payload=True)
        assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all()
        assert (bits(a[2]) == bits(b[2])).all() and (bits(a[3]) == bits(b[3])).all()


def _aimed_rays(sc, n, seed, dist_lo, dist_hi):
    """Rays that start dist_lo..dist_hi scene sizes away and aim at a random point of the scene's box."""
    rng = np.random.default_rng(seed)
    mn, mx = sc["bounds_min"].astype(np.float64), sc["bounds_max"].astype(np.float64)
    size = float(np.abs(np.concatenate([mn, mx])).max())
    target = rng.uniform(mn, mx, (n, 3))
    away = rng.normal(size=(n, 3)); away /= np.linalg.norm(away, axis=1, keepdims=True)
    o = target + away * size * (10.0 ** rng.uniform(np.log10(dist_lo), np.log10(dist_hi), (n, 1)))
    d = target - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], 1).astype(np.float32)


@pytest.mark.parametrize("name", ["suzanne", "teapot"])
def test_distant_origins(scenes, oracle, name):
    """Rays that start far outside the scene.  Up to 16 x the scene's largest |coordinate| the padded boxes must hold (tree ==
    all-triangle scan on the GPU, at scale, and == the oracle); beyond that limit the library answers with the exact scan
    itself (bvh.cuh: ray_is_far), so origins a million scene sizes away are still the reference's answers bit for bit."""
    sc = load_scene(name)
    s = scenes(name)
    near = _aimed_rays(sc, 1_000_000, 41, 2.0, 15.0)        # inside the limit: the tree is walked
    a, b = s.HitScene(near), s.HitScene(near, mode=tm.HIT_BRUTE)
    hit = b[0] >= 0
    assert hit.mean() > 0.1
    assert (a[0] == b[0]).all() and (bits(a[1])[hit] == bits(b[1])[hit]).all() and (bits(a[2])[hit] == bits(b[2])[hit]).all()
    assert ((s.HitScene(near, mode=tm.HIT_ANY)[0] == 1) == hit).all()
    far = _aimed_rays(sc, 20000, 42, 10.0, 1.0e6)            # across and far beyond the limit
    ids, t, pos, nrm = s.HitScene(far, tMax=3.0e38)
    oid, ot, opos, onrm = oracle.hit_brute(sc["tris"], far, tmax=3.0e38)
    ohit = oid >= 0
    assert (ids == oid).all() and (bits(t)[ohit] == bits(ot)[ohit]).all()
    assert (bits(pos)[ohit] == bits(opos)[ohit]).all() and (bits(nrm)[ohit] == bits(onrm)[ohit]).all()
    assert ((s.HitScene(far, mode=tm.HIT_ANY, tMax=3.0e38)[0] == 1) == ohit).all()


def test_nan_rays_miss_quickly(scenes, oracle):
    """NaN rays (the reference's normalize(0) scatter makes them) are misses for the exact test; the tree must say
    so too, and without walking every node (fmin/fmax drop NaN operands)."""
    sc = load_scene("teapot")
    rays = _random_rays(sc, 4096, 3, axis_parallel=False)
    rays[::7, 3:] = np.nan
    rays[3::11, 0] = np.nan
    rays[5::13, 4] = np.nan
    for mode in (tm.HIT_CLOSEST, tm.HIT_BRUTE):
        ids, *_ = scenes("teapot").HitScene(rays, mode=mode)
        oid, *_ = oracle.hit_brute(sc["tris"], rays)
        assert (ids == oid).all()
    bad = np.isnan(rays).any(1)
    assert (ids[bad] == -1).all()
    aid, *_ = scenes("teapot").HitScene(rays, mode=tm.HIT_ANY)
    assert ((aid == 1) == (oid >= 0)).all()


def test_edge_cases():
    ray = np.array([[0.2, 0.2, 1, 0, 0, -1]], np.float32)
    with tm.Scene(np.zeros((0, 9), np.float32)) as s:  # empty scene: everything misses
        ids, *_ = s.HitScene(ray)
        assert ids[0] == -1
        assert s.HitScene(np.zeros((0, 6), np.float32))[0].shape == (0,)
        assert s.HitScene(ray, mode=tm.HIT_SUN)[0][0] == -1 and s.HitScene(ray, mode=tm.HIT_ANY)[0][0] == -1
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    with tm.Scene(tri) as s:
        ids, t, pos, nrm = s.HitScene(ray)
        assert ids[0] == 0 and t[0] == 1.0 and nrm[0].tolist() == [0, 0, 1]
    with tm.Scene(np.repeat(tri, 100, 0)) as s:  # coincident triangles: lowest index wins the tie
        assert s.HitScene(ray)[0][0] == 0
        assert s.info()["tri_count"] == 100
    degenerate = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 2, 0, 0]], np.float32)  # zero-area
    with tm.Scene(np.concatenate([degenerate, tri])) as s:
        assert s.HitScene(ray)[0][0] == 2


@pytest.mark.parametrize("name,w,h,spp", [("cube", 160, 90, 4), ("suzanne", 128, 72, 3), ("teapot", 64, 36, 2), ("triangle", 33, 17, 5),
                                          ("cube", 70, 41, 20), ("suzanne", 48, 28, 9),  # several chunks per pixel (chunk length = spp/32 clamped to 1..8)
                                          ("cube", 640, 360, 8),   # 57 600 warp items: launch_render picks the 1024-thread instantiation the bench times
                                          ("cube", 640, 360, 4),   # BASELINE config 1 at its full size
                                          ("suzanne", 640, 360, 4),  # BASELINE config 2 at its full size (the oracle scans 970 triangles per ray: ~10 s)
                                          ("triangle", 40, 24, 1024),  # the reference's largest spp (main.cpp:275): 128 chunks of 8 samples per pixel
                                          ("cube", 1, 1, 1), ("cube", 7, 3, 33)])  # the smallest frame; a frame smaller than one tile, ragged last chunk
def test_frame_bit_exact_vs_oracle_pixel_mode(scenes, oracle, name, w, h, spp):
    sc = load_scene(name)
    cam = tm.camera_for_scene(f"{name}.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    oimg, orays = oracle.render(sc["tris"], cam, w, h, spp)  # per-pixel RNG streams + trig spec
    img, rays, sec = scenes(name).render(cam, w, h, spp)
    assert rays == orays
    assert (img == oimg).all()
    assert sec > 0


def test_largest_frame_10000x10000_rows_bit_exact_vs_oracle(scenes, oracle):
    """The reference's largest frame (main.cpp:263-270: width, height <= 10000) at 64 spp: 1e8 pixels, 32 chunk sums each, so the
    frame is rendered in 12 bands of 838 rows (4 GB accumulation budget), 1.7e10 rays (the 64-bit counter).  Rows at the frame's
    edges, in the middle and either side of a band boundary equal the oracle's rows bit for bit; two renders agree."""
    name, w, h, spp = "cube", 10000, 10000, 64
    sc = load_scene(name)
    cam = tm.camera_for_scene(f"{name}.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    img, rays, sec = scenes(name).render(cam, w, h, spp)
    assert rays > 2 ** 32 and img.shape == (h, w, 4)
    print(f"10000x10000x64: {rays} rays in {sec:.2f} s = {rays / sec / 1e6:.0f} Mrays/s")
    for row in (0, 837, 838, 5001, 9999):
        oimg, _ = oracle.render(sc["tris"], cam, w, h, spp, rows=(row, row + 1))
        assert (img[row] == oimg[row]).all(), row
    img2, rays2, _ = scenes(name).render(cam, w, h, spp)
    assert rays2 == rays and (img2 == img).all()


@pytest.mark.parametrize("name", SCENES + ["sponza"])
def test_sun_grid_shadow_query_equals_tree_scan_and_reference(scenes, name):
    """TMPT_HIT_SUN -- the integrator's shadow query through the grid in the sun's projection (csrc/sungrid.cuh) -- against the
    reference's own hit flags on every shadow ray of a reference render, and against the tree's any-hit and the all-triangle scan
    on 1.2 M origins: random points in and around the scene, points on triangles (crowded towards edges and vertices), a hair
    below and above them, vertices, edge midpoints."""
    g = load_rays(name)
    sh = g["kind"] == 2
    rays = np.ascontiguousarray(g["rays"][sh], np.float32)
    l = rays[0, 3:6].copy()
    s = scenes(name)
    got = s.HitScene(rays, mode=tm.HIT_SUN)[0]
    assert ((got >= 0) == (g["id"][sh] >= 0)).all()
    tris = sponza_scene()[0] if name == "sponza" else load_scene(name)["tris"]
    v = tris.reshape(-1, 3, 3)
    mn, mx = v.reshape(-1, 3).min(0), v.reshape(-1, 3).max(0)
    rng = np.random.default_rng(17)
    n = 200000
    pick = rng.integers(0, len(v), n)
    bary = rng.dirichlet([0.3, 0.3, 0.3], n).astype(np.float32)
    on = (v[pick] * bary[:, :, None]).sum(1).astype(np.float32)
    o = np.concatenate([rng.uniform(mn - 0.5, mx + 0.5, (n, 3)).astype(np.float32), on, on - l * np.float32(2e-3), on + l * np.float32(1e-4),
                        v[pick, 0], ((v[pick, 0] + v[pick, 1]) * np.float32(0.5))]).astype(np.float32)
    q = np.concatenate([o, np.broadcast_to(l, o.shape)], 1).astype(np.float32)
    for tmin, tmax in ((tm.K_MIN_T, tm.K_MAX_T), (0.0, 2.5)):
        a = s.HitScene(q, tMin=tmin, tMax=tmax, mode=tm.HIT_SUN)[0] >= 0
        b = s.HitScene(q, tMin=tmin, tMax=tmax, mode=tm.HIT_ANY)[0] >= 0
        assert (a == b).all(), (name, tmin, int((a != b).sum()))
    sub = q[:: 7 if name == "sponza" else 1][:300000]
    assert ((s.HitScene(sub, mode=tm.HIT_SUN)[0] >= 0) == (s.HitScene(sub, mode=tm.HIT_BRUTE)[0] >= 0)).all()
    assert 0.02 < a.mean() < 0.98


def test_sun_grid_with_thousands_of_triangles_over_one_cell():
    """3000 nearly coincident triangles: lists longer than sun::kSortMax are left unsorted and without an early exit; the shadow
    query still equals the tree's any-hit and the scan."""
    rng = np.random.default_rng(21)
    base = np.array([[0, 0, 0], [1, 0, 0.2], [0.3, 0.1, 1]], np.float32)
    stack = (base[None] + rng.normal(scale=1e-3, size=(3000, 3, 3))).astype(np.float32).reshape(-1, 9)
    far = (rng.uniform(-2, 2, (50, 1, 3)) + rng.normal(scale=0.2, size=(50, 3, 3))).astype(np.float32).reshape(-1, 9)
    tris = np.concatenate([stack, far]).astype(np.float32)
    l = load_rays("cube")["rays"][load_rays("cube")["kind"] == 2][0, 3:6].astype(np.float32)
    o = rng.uniform(-2.5, 2.5, (20000, 3)).astype(np.float32)
    q = np.concatenate([o, np.broadcast_to(l, o.shape)], 1).astype(np.float32)
    with tm.Scene(tris) as s:
        a, b, c = (s.HitScene(q, mode=m)[0] >= 0 for m in (tm.HIT_SUN, tm.HIT_ANY, tm.HIT_BRUTE))
        assert (a == b).all() and (a == c).all() and 0.02 < a.mean() < 0.98


def test_far_camera_frame_bit_exact_vs_oracle(scenes, oracle):
    """A camera 40 scene sizes away: its origin is beyond the far limit of the padded boxes (bvh.cuh: ray_is_far), so
    launch_render picks the render instantiation that answers such rays with the exact scan.  Still the oracle's bytes."""
    sc = load_scene("suzanne")
    mn, mx = sc["bounds_min"].astype(np.float64), sc["bounds_max"].astype(np.float64)
    centre, size = (mn + mx) / 2, float(np.abs(np.concatenate([mn, mx])).max())
    w, h, spp = 64, 36, 3
    frm = centre + np.array([0.3, 0.5, 1.0]) / np.linalg.norm([0.3, 0.5, 1.0]) * 40.0 * size
    cam = tm.camera_make(frm, centre, [0, 1, 0], 3.0, w / h, 0.03, float(np.linalg.norm(frm - centre)))
    oimg, orays = oracle.render(sc["tris"], cam, w, h, spp)
    img, rays, _ = scenes("suzanne").render(cam, w, h, spp)
    assert rays == orays and (img == oimg).all()
    assert (img[..., :3].std() > 5)  # the model is in view: not a frame of sky only


@pytest.mark.parametrize("cfg", ["1", "2", "3"])
def test_every_launch_configuration_is_bit_exact_vs_oracle(oracle, tmp_path, cfg):
    """k_render is instantiated as 256 x 4, 512 x 2 and 1024 x 1 threads (TMPT_RENDER_CFG = 1, 2, 3; the headline bench runs
    the last one): each forced in a fresh process, each frame byte-equal to the oracle's, ray count included."""
    import sys
    name, w, h, spp = "suzanne", 200, 120, 6
    sc = load_scene(name)
    cam = tm.camera_for_scene(f"{name}.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    oimg, orays = oracle.render(sc["tris"], cam, w, h, spp)
    out = str(tmp_path / f"cfg{cfg}.npz")
    env = dict(os.environ, TMPT_RENDER_CFG=cfg)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "render_probe.py"), name, str(w), str(h), str(spp), out], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(out)
    assert int(z["rays"]) == orays and (z["img"] == oimg).all()


def _image_metrics(a, b):
    a, b = a[..., :3].astype(np.float64), b[..., :3].astype(np.float64)
    # outlier-robust: the reference poisons a few pixels per frame with NaN -> black
    bad = (a.sum(-1) == 0) | (b.sum(-1) == 0)
    d = (a - b)[~bad]
    mae = np.abs(d).mean()
    psnr = 10 * np.log10(255.0 ** 2 / (d ** 2).mean())
    return mae, psnr, int(bad.sum()), np.abs(a[~bad].mean(0) - b[~bad].mean(0)).max()


@pytest.mark.parametrize("name,spp,mae_tol,psnr_tol", [("cube", 1024, 0.45, 44.0), ("suzanne", 256, 0.9, 38.0)])
def test_converged_image_vs_reference_render(scenes, name, spp, mae_tol, psnr_tol):
    """Statistical gate against the REFERENCE's own render (row RNG, libm trig): two independent
    reference renders differ by MAE 0.21 / 0.41 at 1024 / 256 spp; the bars are ~2x that."""
    from PIL import Image
    w, h = 320, 180
    ref = np.array(Image.open(os.path.join(GOLD, "images", f"{name}_{w}x{h}_{spp}spp.png")).convert("RGB"))
    sc = load_scene(name)
    cam = tm.camera_for_scene(f"{name}.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    img, rays, _ = scenes(name).render(cam, w, h, spp)
    mae, psnr, bad, dmean = _image_metrics(img[::-1], ref)
    print(f"{name}: MAE {mae:.3f} PSNR {psnr:.2f} dB, {bad} masked pixels, max channel-mean diff {dmean:.3f}")
    assert mae <= mae_tol and psnr >= psnr_tol and bad <= 16 and dmean <= 0.25


@pytest.mark.parametrize("stripe,world,w,h", [(4, 3, 100, 50), (0, 3, 100, 50), (0, 8, 203, 61), (0, 2, 64, 36), (7, 4, 90, 45)])
def test_stripes_compose_to_the_single_gpu_frame(scenes, stripe, world, w, h):
    """Multi-GPU partition, emulated on one GPU: each rank's share + unpack == the whole-frame render, for row stripes
    (stripe > 0) and for the default tile interleave (stripe 0; 100 and 203 pixels = 13 and 26 tiles per row, not multiples of
    the world size, so some ranks' last local tile of a row lies outside the frame)."""
    import torch
    sc = load_scene("suzanne")
    spp = 2
    cam = tm.camera_for_scene("suzanne.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    s = scenes("suzanne")
    full, full_rays, _ = s.render(cam, w, h, spp)
    from toymeshpathtracer_b200 import multigpu
    rows, max_rows = multigpu.stripe_plan(h, stripe, world)
    gathered = torch.zeros((world, max_rows, multigpu.local_width(w, stripe, world), 4), dtype=torch.uint8, device="cuda")
    rays = torch.zeros(world, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for r in range(world):
            s.render_stripes(cam, w, h, spp, stripe, r, world, gathered[r].data_ptr(), rays[r:].data_ptr(), stream=st.cuda_stream)
        frame = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
        tm._check(tm.lib().tmpt_unpack_stripes(gathered.data_ptr(), w, h, stripe, world, 0, frame.data_ptr(), st.cuda_stream))
    st.synchronize()
    assert (frame.cpu().numpy() == full).all()
    assert int(rays.sum()) == full_rays
    # peer-frame form: every rank writes straight into one full frame
    frame2 = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(st):
        for r in range(world):
            s.render_stripes(cam, w, h, spp, stripe, r, world, 0, rays[r:].data_ptr(), peer_frame_ptr=frame2.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    assert (frame2.cpu().numpy() == full).all()


def test_cli_end_to_end(tmp_path):
    """The drop-in command line: `<width> <height> <spp> <datafile>` -> output.png + the three report lines."""
    from PIL import Image
    sc = load_scene("cube")
    obj = tmp_path / "cube.obj"
    tris = sc["tris"][:-2]  # the loader adds the floor itself
    with open(obj, "w") as f:
        for v in tris.reshape(-1, 3):
            f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in v))
        for i in range(tris.shape[0]):
            f.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
    r = subprocess.run([tm.CLI_PATH, "160", "90", "4", str(obj)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    assert lines[0].startswith(f"Initialized scene '{obj}' (14 tris) in ")
    assert lines[1].startswith("Rendered scene at 160x90,4spp in ")
    assert lines[2].startswith("- ") and lines[2].endswith(" K Rays/s")
    img = np.array(Image.open(tmp_path / "output.png"))
    with tm.Scene(sc["tris"]) as s:
        cam = tm.camera_for_scene(str(obj), sc["bounds_min"], sc["bounds_max"], 160, 90)
        want, rays, _ = s.render(cam, 160, 90, 4)
    assert img.shape == (90, 160, 4) and (img == want[::-1]).all()
    assert abs(float(lines[2].split()[1]) - rays / 1000.0) < 0.06


def _peer_worker(rank, world, port, w, h, spp, stripe, out):
    import torch
    import torch.distributed as dist
    from toymeshpathtracer_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # two ranks share the one GPU of this box: no NCCL
    try:
        sc = load_scene("suzanne")
        cam = tm.camera_for_scene("suzanne.obj", sc["bounds_min"], sc["bounds_max"], w, h)
        with tm.Scene(sc["tris"], device=0) as s:
            peer = multigpu.PeerFrame(w, h, rank, world, 0)
            st = torch.cuda.Stream()
            ok = True
            full, full_rays, _ = s.render(cam, w, h, spp)
            frames = []
            with torch.cuda.stream(st):  # three frames back to back: the double-buffered peer frame alternates A, B, A
                for _ in range(3):
                    frame, rays = multigpu.render_frame(s, cam, w, h, spp, rank, world, stripe=stripe, peer=peer)
                    if rank == 0:
                        frames.append(frame.clone())  # consumed on the render stream, as PeerFrame asks
                    last_rays = rays
            st.synchronize()
            total = last_rays.cpu()
            dist.all_reduce(total)
            if rank == 0:
                ok = all(bool((f.cpu().numpy() == full).all()) for f in frames) and int(total) == full_rays
                ok = ok and frames[0].data_ptr() != frames[1].data_ptr()
                out.put(ok)
            dist.barrier()
            peer.close()
    finally:
        dist.destroy_process_group()


def test_peer_frame_across_processes():
    """The fused gather: two ranks (processes) write their tiles straight into rank 0's (double-buffered) frame through CUDA IPC."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, 96, 50, 2, 0, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get() is True


# ---- BASELINE.json configurations at their full sizes, through size-independent properties ----------------------

def _sponza_tris():
    from tools.gen_sponza import triangles
    from oracle.pyoracle import Oracle
    return Oracle().add_floor(triangles())  # (tris incl. the two floor triangles, boundsMin, boundsMax) like LoadScene


def test_config3_teapot_720p_16spp_properties(scenes):
    """teapot.obj 1280x720 16 spp, 1 GPU: deterministic, and the 8-way stripe partition composes to the same frame
    (several chunks per pixel, so the accumulate + resolve path is exercised at full size)."""
    import torch
    from toymeshpathtracer_b200 import multigpu
    sc = load_scene("teapot")
    w, h, spp = 1280, 720, 16
    cam = tm.camera_for_scene("teapot.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    s = scenes("teapot")
    a, ra, _ = s.render(cam, w, h, spp)
    b, rb, _ = s.render(cam, w, h, spp)
    assert ra == rb and (a == b).all()
    assert 30e6 < ra < 50e6  # the reference counts 38.2 M rays at this configuration; streams differ
    assert a[..., 3].min() == 255 and a[..., :3].std() > 10
    frame = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    rays = torch.zeros(8, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for r in range(8):
            s.render_stripes(cam, w, h, spp, multigpu.DEFAULT_STRIPE_ROWS, r, 8, 0, rays[r:].data_ptr(), peer_frame_ptr=frame.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    assert int(rays.sum()) == ra and (frame.cpu().numpy() == a).all()


def test_config4_sponza_tree_vs_brute_force():
    """sponza (66 452 triangles): nearest-hit ids / t / payload of the SAH tree == the all-triangle scan, on primary,
    bounce and shadow rays of the Sponza camera (the rays configs 4 and 5 shoot)."""
    tris, mn, mx = _sponza_tris()
    assert tris.shape[0] == 66452
    w, h = 320, 180
    cam = tm.camera_for_scene("x/sponza.obj", mn, mx, w, h)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    o = np.broadcast_to(cam[0:3], (h * w, 3))
    d = cam[3:6] + ((xs.ravel() + 0.5) / w)[:, None] * cam[6:9] + ((ys.ravel() + 0.5) / h)[:, None] * cam[9:12] - cam[0:3]
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], 1).astype(np.float32)
    rng = np.random.default_rng(4)
    light = np.array([-0.7, 1.0, 0.5]); light /= np.linalg.norm(light)
    with tm.Scene(tris) as s:
        assert s.info()["builder"] == tm.BUILD_DEFAULT
        for bounce in range(3):
            a = s.HitScene(rays)
            b = s.HitScene(rays, mode=tm.HIT_BRUTE)
            hit = a[0] >= 0
            assert (a[0] == b[0]).all() and (bits(a[1])[hit] == bits(b[1])[hit]).all()
            assert (bits(a[2])[hit] == bits(b[2])[hit]).all() and (bits(a[3])[hit] == bits(b[3])[hit]).all()
            pos, nrm = a[2][hit], a[3][hit]
            shadow = np.concatenate([pos, np.broadcast_to(light, pos.shape)], 1).astype(np.float32)
            sa = s.HitScene(shadow, mode=tm.HIT_ANY)[0]
            sb = s.HitScene(shadow, mode=tm.HIT_BRUTE)[0]
            assert ((sa == 1) == (sb >= 0)).all()
            r = rng.normal(size=pos.shape); r /= np.linalg.norm(r, axis=1, keepdims=True)
            nd = nrm + r; nd /= np.maximum(np.linalg.norm(nd, axis=1, keepdims=True), 1e-20)
            rays = np.concatenate([pos, nd], 1).astype(np.float32)


def test_config4_sponza_640x360_4spp_partitions():
    """sponza 640x360 4 spp: the frame is the same for 1, 2, 4 and 8 ranks (emulated on this GPU), and so is the ray count."""
    import torch
    from toymeshpathtracer_b200 import multigpu
    tris, mn, mx = _sponza_tris()
    w, h, spp = 640, 360, 4
    cam = tm.camera_for_scene("x/sponza.obj", mn, mx, w, h)
    with tm.Scene(tris) as s:
        full, full_rays, _ = s.render(cam, w, h, spp)
        assert 8e6 < full_rays < 20e6
        st = torch.cuda.Stream()
        for world in (2, 4, 8):
            frame = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
            rays = torch.zeros(world, dtype=torch.int64, device="cuda")
            with torch.cuda.stream(st):
                for r in range(world):
                    s.render_stripes(cam, w, h, spp, multigpu.DEFAULT_STRIPE_ROWS, r, world, 0, rays[r:].data_ptr(), peer_frame_ptr=frame.data_ptr(), stream=st.cuda_stream)
            st.synchronize()
            assert int(rays.sum()) == full_rays and (frame.cpu().numpy() == full).all()
            per_rank = rays.cpu().numpy()
            assert per_rank.max() < 1.03 * per_rank.mean()  # the tile interleave balances the work (row stripes of 4: within 15 %)


def test_config5_sponza_1080p_64spp_rank_of_eight():
    """The headline frame, 8-way partition: one rank's share (what each GPU of the 8xB200 run renders) is deterministic
    and carries 1/8 of the rays within 5 %."""
    import torch
    from toymeshpathtracer_b200 import multigpu
    tris, mn, mx = _sponza_tris()
    w, h, spp, world = 1920, 1080, 64, 8
    cam = tm.camera_for_scene("x/sponza.obj", mn, mx, w, h)
    rows, max_rows = multigpu.stripe_plan(h, multigpu.DEFAULT_STRIPE_ROWS, world)
    lw = multigpu.local_width(w, multigpu.DEFAULT_STRIPE_ROWS, world)
    with tm.Scene(tris) as s:
        outs = []
        st = torch.cuda.Stream()
        for rep in range(2):
            packed = torch.zeros((max_rows, lw, 4), dtype=torch.uint8, device="cuda")
            rays = torch.zeros(1, dtype=torch.int64, device="cuda")
            with torch.cuda.stream(st):
                s.render_stripes(cam, w, h, spp, multigpu.DEFAULT_STRIPE_ROWS, 3, world, packed.data_ptr(), rays.data_ptr(), stream=st.cuda_stream)
            st.synchronize()
            outs.append((packed.cpu().numpy(), int(rays)))
        assert outs[0][1] == outs[1][1] and (outs[0][0] == outs[1][0]).all()
        assert abs(outs[0][1] / (2.3774e9 / 8) - 1) < 0.05
        assert outs[0][0][: rows[3], :, 3].min() == 255


@pytest.mark.parametrize("name", ["suzanne", "teapot", "sponza"])
def test_grazing_rays_tree_vs_brute_force(scenes, name):
    """Stress of box conservativeness where Moller-Trumbore is least accurate: rays that start ON a triangle's plane
    (inside it, on its edges, at its vertices) and leave almost inside that plane, so that det is tiny and the computed
    u, v, t carry their largest errors.  The tree must still return exactly what the all-triangle scan returns."""
    if name == "sponza":
        tris = _sponza_tris()[0]
        s = tm.Scene(tris)
    else:
        tris = load_scene(name)["tris"]
        s = scenes(name)
    rng = np.random.default_rng(17)
    n = 400_000
    T = tris.reshape(-1, 3, 3).astype(np.float64)
    k = rng.integers(0, T.shape[0], n)
    v0, v1, v2 = T[k, 0], T[k, 1], T[k, 2]
    bary = rng.dirichlet([0.6, 0.6, 0.6], n)
    bary[: n // 8] = np.eye(3)[rng.integers(0, 3, n // 8)]                      # vertices
    edge = rng.random(n // 8)
    bary[n // 8: n // 4] = np.stack([edge, 1 - edge, np.zeros_like(edge)], 1)   # on an edge
    p = bary[:, :1] * v0 + bary[:, 1:2] * v1 + bary[:, 2:] * v2
    e1, e2 = v1 - v0, v2 - v0
    nrm = np.cross(e1, e2)
    nlen = np.linalg.norm(nrm, axis=1, keepdims=True)
    ok = nlen[:, 0] > 0
    nrm = nrm / np.maximum(nlen, 1e-300)
    a = rng.random((n, 1)) * 2 * np.pi
    t1 = e1 / np.maximum(np.linalg.norm(e1, axis=1, keepdims=True), 1e-300)
    t2 = np.cross(nrm, t1)
    tilt = (10.0 ** rng.uniform(-8, -1, (n, 1))) * rng.choice([-1.0, 1.0], (n, 1))
    d = np.cos(a) * t1 + np.sin(a) * t2 + tilt * nrm
    with np.errstate(invalid="ignore", divide="ignore"):  # zero-area triangles are dropped by `ok` below
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    back = rng.uniform(0.0, 2.0, (n, 1))                                            # start up to 2 units before the point
    rays = np.concatenate([p - back * d, d], 1)[ok].astype(np.float32)
    a_ = s.HitScene(rays)
    b_ = s.HitScene(rays, mode=tm.HIT_BRUTE)
    hit = b_[0] >= 0
    assert (a_[0] == b_[0]).all(), f"{(a_[0] != b_[0]).sum()} of {rays.shape[0]} grazing rays differ"
    assert (bits(a_[1])[hit] == bits(b_[1])[hit]).all()
    sa = s.HitScene(rays, mode=tm.HIT_ANY)[0]
    assert ((sa == 1) == hit).all()
    if name == "sponza":
        s.close()


def test_sponza_image_vs_reference_binary(tmp_path):
    """Config 4's scene against the REFERENCE PROGRAM itself (oracle/_ref/TrimeshTracer, built from the unmodified
    sources; it travels to the GPU box): same .obj, same argv, both through their command lines.  Statistical gate
    (different RNG layout): two reference renders with different seeds differ by about the same amount."""
    from PIL import Image
    from oracle.pyoracle import REF_BIN
    from tools.gen_sponza import write_obj
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/TrimeshTracer not built")
    obj = str(tmp_path / "sponza.obj")
    write_obj(obj)
    w, h, spp = 160, 90, 128
    (tmp_path / "ref").mkdir(); (tmp_path / "gpu").mkdir()
    r = subprocess.run([REF_BIN, str(w), str(h), str(spp), obj], cwd=tmp_path / "ref", capture_output=True, text=True, check=True)
    g = subprocess.run([tm.CLI_PATH, str(w), str(h), str(spp), obj], cwd=tmp_path / "gpu", capture_output=True, text=True, check=True)
    ref = np.array(Image.open(tmp_path / "ref" / "output.png").convert("RGB"))
    gpu = np.array(Image.open(tmp_path / "gpu" / "output.png").convert("RGB"))
    a, b = gpu.astype(np.float64), ref.astype(np.float64)
    # NaN-poisoned pixels are black in one image and clearly lit in the other; dark pixels are legitimate here
    poisoned = ((a.sum(-1) == 0) & (b.sum(-1) > 120)) | ((b.sum(-1) == 0) & (a.sum(-1) > 120))
    a[poisoned] = b[poisoned] = 0.0
    mae = np.abs(a - b).mean()
    dmean = np.abs(a.mean((0, 1)) - b.mean((0, 1))).max()
    # a 5x5 box filter divides the per-pixel noise by ~5 and leaves any bias intact
    box = lambda x: x[: h // 5 * 5, : w // 5 * 5].reshape(h // 5, 5, w // 5, 5, 3).mean((1, 3))
    mae_box = np.abs(box(a) - box(b)).mean()
    kr = float(re.search(r"- ([0-9.]+) K Rays", r.stdout).group(1))
    kg = float(re.search(r"- ([0-9.]+) K Rays", g.stdout).group(1))
    print(f"sponza {w}x{h}x{spp}: MAE {mae:.3f} (5x5 box: {mae_box:.3f}), {int(poisoned.sum())} poisoned, channel-mean diff {dmean:.3f}, "
          f"K rays ref {kr} gpu {kg}")
    assert "(66452 tris)" in r.stdout and "(66452 tris)" in g.stdout
    assert abs(kg / kr - 1) < 0.01          # same scene, same integrator: ray counts agree to a fraction of a percent
    assert mae <= 6.0 and mae_box <= 1.5 and dmean <= 0.3 and poisoned.sum() <= 16


def test_sponza_converged_image_vs_reference_1024spp(scenes, kat):
    """north_star's second check on the headline scene: the converged image against the reference CPU render, per-pixel MAE
    AND PSNR.  tests/golden/images/sponza_160x90_1024spp_{a,b}.png are two INDEPENDENT 1024-spp renders by the reference's own
    row functor (different row seeds, oracle/gen_golden.py); their distance (MAE 1.44, PSNR 42.3 dB: the noise floor of this
    interior at 1024 spp) is in kat.json and the gates are derived from it: an unbiased GPU frame is as far from `a` as `b` is."""
    from PIL import Image
    w, h, spp = 160, 90, 1024
    ref_a = np.array(Image.open(os.path.join(GOLD, "images", f"sponza_{w}x{h}_{spp}spp_a.png")).convert("RGB"))
    ref_b = np.array(Image.open(os.path.join(GOLD, "images", f"sponza_{w}x{h}_{spp}spp_b.png")).convert("RGB"))
    pair = kat["sponza"]["image_pair_160x90_1024spp"]
    tris, mn, mx = sponza_scene()
    cam = tm.camera_for_scene("x/sponza.obj", mn, mx, w, h)
    img, rays, _ = scenes("sponza").render(cam, w, h, spp)
    gpu = img[::-1, :, :3]

    def dist(a, b):
        a, b = a.astype(np.float64), b.astype(np.float64)
        bad = ((a.sum(-1) == 0) & (b.sum(-1) > 120)) | ((b.sum(-1) == 0) & (a.sum(-1) > 120))  # NaN-poisoned pixels
        d = (a - b)[~bad]
        return np.abs(d).mean(), 10 * np.log10(255.0 ** 2 / (d ** 2).mean()), int(bad.sum()), np.abs(a[~bad].mean(0) - b[~bad].mean(0)).max()

    res = [dist(gpu, ref_a), dist(gpu, ref_b)]
    print(f"sponza {w}x{h}x{spp}: vs a MAE {res[0][0]:.3f} PSNR {res[0][1]:.2f} dB, vs b MAE {res[1][0]:.3f} PSNR {res[1][1]:.2f} dB; "
          f"reference pair MAE {pair['mae']:.3f} PSNR {pair['psnr']:.2f} dB; rays {rays} (reference {pair['ray_count_a']})")
    assert abs(rays / pair["ray_count_a"] - 1) < 0.002
    for mae, psnr, bad, dmean in res:
        assert mae <= 1.05 * pair["mae"] and psnr >= pair["psnr"] - 0.3 and bad <= 8 and dmean <= 0.15  # (host emulation of this frame: 1.007x, -0.04 dB, 0.03)


def test_render_multi_single_process(scenes):
    """tmpt_render_multi: replicas on every visible device (one on the test box), pixels stored into device 0's frame."""
    sc = load_scene("suzanne")
    w, h, spp = 120, 66, 12
    cam = tm.camera_for_scene("suzanne.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    want, want_rays, _ = scenes("suzanne").render(cam, w, h, spp)
    n = min(tm.device_count(), 4)
    reps = [tm.Scene(sc["tris"], device=d) for d in range(n)]
    try:
        img, rays, sec = tm.render_multi(reps, cam, w, h, spp)
    finally:
        for r in reps:
            r.close()
    assert rays == want_rays and (img == want).all() and sec > 0


def test_large_random_scene(oracle):
    """500 k small random triangles (a soup: the worst case for the builder's depth and for the traversal stack):
    the build succeeds, the tree agrees with the all-triangle scan, nothing overflows."""
    rng = np.random.default_rng(23)
    n = 500_000
    c = rng.uniform(-50, 50, (n, 1, 3))
    tris = (c + rng.normal(scale=0.4, size=(n, 3, 3))).astype(np.float32).reshape(n, 9)
    rays = _random_rays({"bounds_min": np.full(3, -50, np.float32), "bounds_max": np.full(3, 50, np.float32)}, 20000, 8)
    for flags in (tm.BUILD_DEFAULT, tm.BUILD_LBVH):
        with tm.Scene(tris, flags=flags) as s:
            info = s.info()
            a = s.HitScene(rays)
            b = s.HitScene(rays, mode=tm.HIT_BRUTE)
            print(f"builder {info['builder']}: {info['node_count']} nodes, depth {info['max_depth']}, build {info['build_ms']:.1f} ms, "
                  f"{(a[0] >= 0).sum()} of {len(rays)} rays hit")
            assert info["tri_count"] == n and info["max_depth"] <= 21
            assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all()
            assert (a[0] >= 0).sum() > 1000


def test_path_regeneration_kernel_is_bit_exact_and_chosen_for_open_frames(tmp_path, oracle):
    """k_render_paths (lanes take their next sample at bounce boundaries) against the oracle and against k_render:
    TMPT_RENDER_PATHS=1 / 0 force either kernel -- same bytes, same ray count, one-shot and over progressive passes, on an open
    scene (suzanne, vs the oracle) and a closed one (sponza).  Left alone, the probe picks k_render_paths for the open frame and
    k_render for the hall."""
    import sys

    def run(name, w, h, spp, paths, extra=()):
        out = str(tmp_path / f"{name}_{paths}_{len(extra)}.npz")
        env = dict(os.environ)
        env.pop("TMPT_RENDER_PATHS", None)
        if paths is not None:
            env["TMPT_RENDER_PATHS"] = paths
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "render_probe.py"), name, str(w), str(h), str(spp), out, *extra],
                           env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        z = np.load(out)
        return z["img"], int(z["rays"]), int(z["kernel"]), float(z["escape"])

    name, w, h, spp = "suzanne", 200, 120, 6
    sc = load_scene(name)
    cam = tm.camera_for_scene(f"{name}.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    oimg, orays = oracle.render(sc["tris"], cam, w, h, spp)
    for paths, want in (("0", 0), ("1", 1), (None, 1)):
        img, rays, kernel, escape = run(name, w, h, spp, paths)
        assert kernel == want and rays == orays and (img == oimg).all(), (paths, kernel, escape)
    assert escape > 0.3  # (an object on a floor under the sky)
    prog = [run(name, 101, 57, 20, paths, ("2", "1", "3")) for paths in ("0", "1")]
    assert prog[0][2] == 0 and prog[1][2] == 1 and prog[0][1] == prog[1][1] and (prog[0][0] == prog[1][0]).all()
    hall = [run("sponza", 240, 136, 4, paths) for paths in ("0", "1", None)]
    assert [r[2] for r in hall] == [0, 1, 0] and hall[2][3] < 0.1
    assert hall[0][1] == hall[1][1] == hall[2][1] and (hall[0][0] == hall[1][0]).all() and (hall[0][0] == hall[2][0]).all()


def test_render_kernel_choice_follows_the_camera_and_never_changes_the_bytes():
    """The probe behind the k_render / k_render_paths choice runs when camera or frame size change and is picked up one frame
    later (no host stall): a scene object rendering a SEQUENCE of cameras -- inside the hall (closed), far outside it looking
    at the building (open), other frame sizes -- produces, frame by frame, the bytes a fresh scene produces for that camera
    alone, whichever kernel the history made it use; and the choice does follow the camera."""
    if os.environ.get("TMPT_RENDER_PATHS") or os.environ.get("TMPT_RENDER_CFG"):
        pytest.skip("render kernel forced by the environment")
    tris, mn, mx = _sponza_tris()
    w, h, spp = 192, 108, 4
    inside = tm.camera_for_scene("x/sponza.obj", mn, mx, w, h)
    c = 0.5 * (mn + mx)
    size = float(np.max(mx - mn))
    outside = tm.camera_make(c + np.array([1.5, 1.2, 1.1], np.float32) * size, c, [0, 1, 0], 40.0, w / h, 0.0, 2.5 * size)
    seq = [(inside, w, h), (outside, w, h), (outside, w, h), (outside, 96, 54), (inside, w, h), (inside, w, h), (inside, 96, 54)]
    kernels = []
    with tm.Scene(tris) as s:
        assert s.render_kernel_choice()[0] == -1
        for cam, fw, fh in seq:
            img, rays, _ = s.render(cam, fw, fh, spp)
            kernels.append(s.render_kernel_choice())
            with tm.Scene(tris) as fresh:
                fimg, frays, _ = fresh.render(cam, fw, fh, spp)
                first = fresh.render_kernel_choice()
            assert rays == frays and (img == fimg).all()
            assert first[0] == (0 if cam is inside else 1), first  # (a scene's first frame waits for its probe)
    ks = [k for k, _ in kernels]
    assert ks[0] == 0 and ks[2] == 1 and ks[5] == 0, kernels  # the second frame of an unchanged camera has the new decision (render() is synchronous)


def test_frames_do_not_depend_on_the_sun_grid(tmp_path):
    """TMPT_SUN_GRID=0 (shadow rays walk the tree, the fallback of a scene without a grid), a coarse grid and the default grid:
    the same bytes and ray counts, through both render kernels (suzanne: k_render_paths, sponza: k_render) and progressive passes."""
    import sys

    def run(name, w, h, spp, grid, extra=()):
        out = str(tmp_path / f"{name}_{grid}_{len(extra)}.npz")
        env = dict(os.environ)
        env.pop("TMPT_SUN_GRID", None)
        if grid is not None:
            env["TMPT_SUN_GRID"] = grid
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "render_probe.py"), name, str(w), str(h), str(spp), out, *extra],
                           env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        z = np.load(out)
        return z["img"].tobytes(), int(z["rays"])

    for name, w, h, spp, extra in (("suzanne", 200, 120, 6, ()), ("sponza", 240, 136, 4, ()), ("suzanne", 101, 57, 20, ("2", "1", "3"))):
        base = run(name, w, h, spp, None, extra)
        for grid in ("0", "5", "300"):
            assert run(name, w, h, spp, grid, extra) == base, (name, grid)


def test_regeneration_render_kernel_gives_the_same_bytes(tmp_path):
    """k_render_regen (per-lane ray regeneration; measured, not shipped: it lives in the -DTMPT_EXPERIMENTS=1 build only)
    schedules the same per-lane arithmetic differently: frame and ray count equal the lockstep kernel's, one-shot and over
    progressive passes (which continue the chunk numbering: p.chunk0)."""
    import sys
    from toymeshpathtracer_b200 import build as tb
    exp_lib = tb.build_variant("exp", ["-DTMPT_EXPERIMENTS=1"])
    outs = []
    for k, lib in (("0", None), ("1", exp_lib), ("3", exp_lib)):
        env = dict(os.environ, TMPT_RENDER_KERNEL=k)
        if lib:
            env["TMPT_LIB"] = lib
        else:
            env.pop("TMPT_RENDER_KERNEL")
        res = []
        for extra in ([], ["2", "1", "3"]):  # one-shot 20 spp; three progressive passes (6 chunks = 48 samples)
            out = str(tmp_path / f"regen{k}_{len(extra)}.npz")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "render_probe.py"), "suzanne", "101", "57", "20", out, *extra], env=env,
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            z = np.load(out)
            res.append((z["img"].tobytes(), int(z["rays"])))
        outs.append(res)
    assert outs[0] == outs[1] == outs[2]
    # the shipped library refuses the experiment switches instead of ignoring them
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "render_probe.py"), "suzanne", "32", "32", "1", str(tmp_path / "x.npz")],
                       env=dict(os.environ, TMPT_RENDER_KERNEL="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "TMPT_EXPERIMENTS" in r.stderr


def test_refit_moved_vertices(scenes):
    """tmpt_scene_refit: same topology, moved vertices.  Hits equal the all-triangle scan of the NEW positions and a scene
    created from them, a frame rendered after the refit equals the fresh scene's frame byte for byte (the tree only culls)."""
    sc = load_scene("teapot")
    size = float(np.max(sc["bounds_max"] - sc["bounds_min"]))
    v = sc["tris"].reshape(-1, 3).astype(np.float64)
    moved = (v + 0.12 * size * np.stack([np.sin(1.7 * v[:, 1] + 0.3), np.cos(2.1 * v[:, 2]), np.sin(1.3 * v[:, 0] + 1.0)], 1)).astype(np.float32).reshape(-1, 9)
    rays = _random_rays(sc, 60000, 31)
    w, h, spp = 96, 54, 8
    with tm.Scene(sc["tris"]) as s, tm.Scene(moved) as fresh:
        before = s.HitScene(rays)
        sec = s.refit(moved)
        assert 0 < sec < 1.0
        a, b, c = s.HitScene(rays), s.HitScene(rays, mode=tm.HIT_BRUTE), fresh.HitScene(rays)
        hit = a[0] >= 0
        assert hit.sum() > 1000 and (a[0] != before[0]).any()
        for other in (b, c):
            assert (a[0] == other[0]).all() and (bits(a[1])[hit] == bits(other[1])[hit]).all()
            assert (bits(a[2])[hit] == bits(other[2])[hit]).all() and (bits(a[3])[hit] == bits(other[3])[hit]).all()
        anyhit = s.HitScene(rays, mode=tm.HIT_ANY)
        assert ((anyhit[0] >= 0) == hit).all()
        mn, mx = moved.reshape(-1, 3).min(0), moved.reshape(-1, 3).max(0)
        cam = tm.camera_for_scene("teapot.obj", mn, mx, w, h)
        f1, r1, _ = s.render(cam, w, h, spp)
        f2, r2, _ = fresh.render(cam, w, h, spp)
        assert r1 == r2 and (f1 == f2).all()
        info = s.info()
        assert np.allclose(info["bounds_min"], mn) and np.allclose(info["bounds_max"], mx)
        with pytest.raises(tm.TmptError):
            s.refit(moved[:-1])
        s.refit(sc["tris"])  # and back
        again = s.HitScene(rays)
        assert (again[0] == before[0]).all() and (bits(again[1])[before[0] >= 0] == bits(before[1])[before[0] >= 0]).all()


def test_progressive_passes_converge_to_the_one_shot_frame(scenes):
    """tmpt_progressive_*: P chunks of 8 samples traced over several passes give, byte for byte, the frame tmpt_render
    produces at spp = 8 P (P >= 32), and the same number of rays; the frame is rendered in one-chunk and multi-chunk passes."""
    sc = load_scene("suzanne")
    w, h = 85, 47
    cam = tm.camera_for_scene("suzanne.obj", sc["bounds_min"], sc["bounds_max"], w, h)
    s = scenes("suzanne")
    want, want_rays, _ = s.render(cam, w, h, 272)
    s.progressive_begin(w, h)
    total, frames = 0, []
    for n in (1, 3, 12, 1, 17):  # 34 chunks = 272 samples
        img, rays, sec, spp = s.progressive_pass(cam, n)
        total += rays
        frames.append(img)
        assert sec > 0 and img[..., 3].min() == 255
    assert spp == 272 and total == want_rays and (frames[-1] == want).all()
    # the first 32 chunks are the 256 spp frame; earlier passes are noisier estimates of the same image
    want256, rays256, _ = s.render(cam, w, h, 256)
    s.progressive_begin(w, h)
    img, rays, _, spp = s.progressive_pass(cam, 32)
    assert spp == 256 and rays == rays256 and (img == want256).all()
    err = [np.abs(f.astype(np.float64) - want).mean() for f in frames]
    assert err[0] > err[2] > err[-1] == 0.0
    with tm.Scene(sc["tris"]) as fresh:  # a pass without tmpt_progressive_begin is an argument error
        fresh._prog = (w, h)
        with pytest.raises(tm.TmptError):
            fresh.progressive_pass(cam, 1)
