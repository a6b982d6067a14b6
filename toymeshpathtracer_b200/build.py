"""Build the native library in-tree: csrc/*.cu + csrc/*.cpp -> libtmpt.so, bin/TrimeshTracer.

    python -m toymeshpathtracer_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The outputs are git-ignored but travel to
the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import contextlib
import fcntl
import hashlib
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


def _source_hash(extra=()) -> str:
    """sha256 over the contents of every source the library is built from (plus the variant's defines)."""
    h = hashlib.sha256()
    for path in sorted(_sources()):
        h.update(os.path.relpath(path, HERE).encode() + b"\0")
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(list(CUFLAGS) + list(CXXFLAGS) + list(extra)).encode())
    return h.hexdigest()


def _stamp(target: str) -> str:
    return os.path.join(HERE, "build", os.path.basename(target) + ".srchash")


def _stale(target: str, extra=()) -> bool:
    """Stale = the target is missing or was built from other source CONTENTS.  (Not mtimes: a `git checkout`, or the copy of the
    tree to a GPU box, changes them without changing a byte -- and a spurious rebuild under torchrun is eight ranks at once.)"""
    if not os.path.exists(target):
        return True
    try:
        with open(_stamp(target)) as f:
            return f.read().strip() != _source_hash(extra)
    except OSError:
        return True


def _mark_built(target: str, extra=()):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(_stamp(target), "w") as f:
        f.write(_source_hash(extra))


@contextlib.contextmanager
def _build_lock():
    """One builder at a time per tree (ranks of one torchrun share it): the others wait, then find the target up to date."""
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


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
    if not force and not _stale(LIB) and os.path.exists(BIN):
        return LIB
    with _build_lock():
        if not force and not _stale(LIB) and os.path.exists(BIN):  # another process built it while this one waited
            return LIB
        if not os.path.exists(NVCC):
            raise RuntimeError(f"nvcc not found at {NVCC} and {LIB} is missing or stale")
        obj_dir = os.path.join(HERE, "build")
        os.makedirs(obj_dir, exist_ok=True)
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        k_o, h_o = os.path.join(obj_dir, "kernels.o"), os.path.join(obj_dir, "host.o")
        tmp_lib = LIB + ".tmp%d" % os.getpid()
        _run([NVCC, *ARCH, *CUFLAGS, "-c", os.path.join(CSRC, "kernels.cu"), "-o", k_o], log=os.path.join(obj_dir, "ptxas.log"))
        _run(["g++", *CXXFLAGS, "-c", os.path.join(CSRC, "host.cpp"), "-o", h_o])
        _run([NVCC, *ARCH, "-shared", "-cudart", "static", "-o", tmp_lib, k_o, h_o])
        os.replace(tmp_lib, LIB)  # (atomic: a process that has the old file mapped keeps it)
        _run(["g++", *CXXFLAGS, os.path.join(CSRC, "main.cpp"), "-o", BIN, "-L" + HERE, "-ltmpt", "-Wl,-rpath,$ORIGIN/.."])
        _mark_built(LIB)
    return LIB


def build_variant(tag: str, defines, force: bool = False) -> str:
    """A second build of the same sources with compile-time switches -> libtmpt_<tag>.so (selected with TMPT_LIB=<path>).
    Used for the experiments build (-DTMPT_EXPERIMENTS=1: kernels that were measured and not adopted) and for A/B runs."""
    lib = os.path.join(HERE, f"libtmpt_{tag}.so")
    defines = list(defines)
    if not force and not _stale(lib, defines):
        return lib
    with _build_lock():
        if not force and not _stale(lib, defines):
            return lib
        if not os.path.exists(NVCC):
            raise RuntimeError(f"nvcc not found at {NVCC} and {lib} is missing or stale")
        obj_dir = os.path.join(HERE, "build")
        os.makedirs(obj_dir, exist_ok=True)
        k_o, h_o = os.path.join(obj_dir, f"kernels_{tag}.o"), os.path.join(obj_dir, f"host_{tag}.o")
        _run([NVCC, *ARCH, *CUFLAGS, *defines, "-c", os.path.join(CSRC, "kernels.cu"), "-o", k_o], log=os.path.join(obj_dir, f"ptxas_{tag}.log"))
        _run(["g++", *CXXFLAGS, "-c", os.path.join(CSRC, "host.cpp"), "-o", h_o])
        tmp_lib = lib + ".tmp%d" % os.getpid()
        _run([NVCC, *ARCH, "-shared", "-cudart", "static", "-o", tmp_lib, k_o, h_o])
        os.replace(tmp_lib, lib)
        _mark_built(lib, defines)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
