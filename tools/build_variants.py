#!/usr/bin/env python
"""Build A/B variants of the library (development): tools/build_variants.py tag=-DX=1,-DY=2 ...  -> libtmpt_<tag>.so each.
Run here (nvcc cross-compiles); the .so files travel to the GPU box, where tools/ab_libs.sh compares them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from toymeshpathtracer_b200 import build as tb  # noqa: E402

for spec in sys.argv[1:]:
    tag, _, defs = spec.partition("=")
    lib = tb.build_variant(tag, [d for d in defs.split(",") if d], force=True)
    log = os.path.join(os.path.dirname(lib), "build", f"ptxas_{tag}.log")
    info = [l.strip() for l in open(log) if "k_renderILb0ELi1024" in l or "bytes stack frame" in l or "Used" in l]
    k = next(i for i, l in enumerate(info) if "Compiling entry function '_Z8k_renderILb0ELi1024" in l)
    print(tag, "->", lib, "|", " ".join(info[k + 1:k + 3]))
