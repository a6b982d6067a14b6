"""pytest configuration: markers, paths and shared fixtures.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol checks (CPU only).
`-m gpu`       : parity tests proper -- the CUDA path through the C ABI vs the oracle.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the reference sources / oracle/_ref (this container)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref, ref_available
    if not ref_available():
        pytest.skip("reference build (oracle/_ref) not available here")
    return Ref()


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLD, "kat.json")) as f:
        return json.load(f)


def load_scene(name):
    z = np.load(os.path.join(GOLD, "scenes", f"{name}.npz"))
    return {k: z[k] for k in z.files}


def load_rays(name):
    z = np.load(os.path.join(GOLD, "rays", f"{name}.npz"))
    return {k: z[k] for k in z.files}


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def from_bits(lst, shape=None):
    a = np.array(lst, dtype=np.uint32).view(np.float32)
    return a.reshape(shape) if shape else a


_SPONZA = None


def sponza_scene(kat=None):
    """The Sponza stand-in exactly as LoadScene builds it from tools/gen_sponza.py's file (model + the two floor triangles):
    (tris, boundsMin, boundsMax).  The triangle bytes are pinned by kat.json["sponza"]["tris_sha256"], recorded when the
    reference itself loaded the same file (oracle/gen_golden.py --sponza)."""
    global _SPONZA
    if _SPONZA is None:
        from oracle.pyoracle import Oracle
        from tools.gen_sponza import triangles
        _SPONZA = Oracle().add_floor(triangles())
    return _SPONZA
