"""Build the native library in-tree: csrc/*.cu + csrc/*.cpp -> libtmpt.so, bin/TrimeshTracer.

    python -m toymeshpathtracer_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The outputs are git-ignored but travel to
the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtmpt.so")
BIN = os.path.join(HERE, "bin", "TrimeshTracer")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall", "-Xptxas", "-v"] + \
    os.environ.get("TMPT_NVCC_EXTRA", "").split()
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(os.path.dirname(HERE), "include", "tmpt.h"))
    return out


def _stale(target: str) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in _sources())


def _run(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))


def build(force: bool = False) -> str:
    """Compile if anything under csrc/ or include/ is newer than the outputs.  Returns the .so path."""
    if not force and not _stale(LIB) and not _stale(BIN):
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC} and {LIB} is missing or stale")
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    k_o, h_o = os.path.join(obj_dir, "kernels.o"), os.path.join(obj_dir, "host.o")
    _run([NVCC, *ARCH, *CUFLAGS, "-c", os.path.join(CSRC, "kernels.cu"), "-o", k_o], log=os.path.join(obj_dir, "ptxas.log"))
    _run(["g++", *CXXFLAGS, "-c", os.path.join(CSRC, "host.cpp"), "-o", h_o])
    _run([NVCC, *ARCH, "-shared", "-cudart", "static", "-o", LIB, k_o, h_o])
    _run(["g++", *CXXFLAGS, os.path.join(CSRC, "main.cpp"), "-o", BIN, "-L" + HERE, "-ltmpt", "-Wl,-rpath,$ORIGIN/.."])
    return LIB


def build_variant(tag: str, defines, force: bool = False) -> str:
    """A second build of the same sources with compile-time switches -> libtmpt_<tag>.so (selected with TMPT_LIB=<path>).
    Used for the experiments build (-DTMPT_EXPERIMENTS=1: kernels that were measured and not adopted) and for A/B runs."""
    lib = os.path.join(HERE, f"libtmpt_{tag}.so")
    if not force and not _stale(lib):
        return lib
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC} and {lib} is missing or stale")
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    k_o, h_o = os.path.join(obj_dir, f"kernels_{tag}.o"), os.path.join(obj_dir, "host.o")
    _run([NVCC, *ARCH, *CUFLAGS, *defines, "-c", os.path.join(CSRC, "kernels.cu"), "-o", k_o], log=os.path.join(obj_dir, f"ptxas_{tag}.log"))
    if not os.path.exists(h_o) or _stale(h_o):
        _run(["g++", *CXXFLAGS, "-c", os.path.join(CSRC, "host.cpp"), "-o", h_o])
    _run([NVCC, *ARCH, "-shared", "-cudart", "static", "-o", lib, k_o, h_o])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
