"""The restatement against the LIVE reference build (oracle/_ref) -- runs where
/root/reference exists; elsewhere the golden-vector tests carry the same checks."""
import os

import numpy as np
import pytest

from conftest import bits, load_scene
from oracle.pyoracle import REF_ROOT, RNG_ROW, TRIG_LIBM

pytestmark = pytest.mark.ref


def test_rng_long_stream(oracle, ref):
    a, b = oracle.rng_states(2024, 100000), ref.rng_states(2024, 100000)
    assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all()
    a, sa = oracle.random_unit_vectors(5, 50000, TRIG_LIBM)
    b, sb = ref.random_unit_vectors(5, 50000)
    assert sa == sb and (bits(a) == bits(b)).all()


@pytest.mark.parametrize("name,w,h,spp", [("cube", 200, 120, 16), ("suzanne", 96, 54, 4), ("teapot", 48, 27, 2)])
def test_render_equals_reference_functor(oracle, ref, name, w, h, spp):
    sc = load_scene(name)
    hdl = ref.scene_from_tris(sc["tris"], sc["bounds_min"], sc["bounds_max"])
    cam = oracle.camera_for_scene(sc["bounds_min"], sc["bounds_max"], w, h)
    rimg, rrc = ref.render(hdl, cam, w, h, spp)
    img, rc = oracle.render(sc["tris"], cam, w, h, spp, RNG_ROW, TRIG_LIBM)
    ref.scene_free(hdl)
    assert rc == rrc and (img == rimg).all()


def test_random_rays_octree_vs_oracle(oracle, ref):
    """Rays the reference render never shoots (random origins inside the root box)."""
    sc = load_scene("suzanne")
    hdl = ref.scene_from_tris(sc["tris"], sc["bounds_min"], sc["bounds_max"])
    rng = np.random.default_rng(7)
    n = 20000
    lo, hi = sc["bounds_min"] - 0.5, sc["bounds_max"] + 0.5
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.concatenate([o, d], 1)
    flag, t, pos, nrm = ref.hit_scene(hdl, rays)
    ids, ot, opos, onrm = oracle.hit_brute(sc["tris"], rays)
    ref.scene_free(hdl)
    hit = ids >= 0
    assert ((flag == 1) == hit).all()
    assert (bits(t)[hit] == bits(ot)[hit]).all() and (bits(pos)[hit] == bits(opos)[hit]).all()


def test_obj_files_present_match_golden(ref):
    path = os.path.join(REF_ROOT, "data", "suzanne.obj")
    if not os.path.exists(path):
        pytest.skip("reference data absent")
    h, tris, mn, mx = ref.scene_load(path)
    ref.scene_free(h)
    assert (bits(tris) == bits(load_scene("suzanne")["tris"])).all()
