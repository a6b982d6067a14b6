#!/bin/bash
# The host emulation (tests/emu: the product's __host__ __device__ headers) built with AddressSanitizer + UndefinedBehaviorSanitizer and
# driven by tools/fuzz_emu.py's three modes, seed by seed.  Usage: tools/fuzz_emu_sanitized.sh [first_seed] [last_seed]   (default 0 8)
# Round 2: seeds 0..45 and the Sponza stand-in (both builders, sun grid at 2048 cells per side, refit, a frame): no report.
set -e
cd "$(dirname "$0")/.."
SO=/tmp/libemu_asan_$$.so
g++ -std=c++17 -O1 -g -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -I${CUDA_INC:-/usr/local/cuda/include} -o $SO tests/emu/emu.cpp
ASAN=$(g++ -print-file-name=libasan.so); UBSAN=$(g++ -print-file-name=libubsan.so)
OMP_NUM_THREADS=2 ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 LD_PRELOAD=$ASAN:$UBSAN python - "$SO" "${1:-0}" "${2:-8}" <<'PY'
import importlib.util, os, sys
sys.path.insert(0, "tests")
import emu_binding
so = sys.argv[1]
emu_binding.build = lambda defines=(), tag="": so   # every Emu() of this process loads the sanitized build
spec = importlib.util.spec_from_file_location("fuzz_emu", os.path.join("tools", "fuzz_emu.py"))
fz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fz)
for seed in range(int(sys.argv[2]), int(sys.argv[3])):
    for mode in (fz.run_seed, fz.refit_seed, fz.render_seed):
        r = mode(seed)
        assert not r[4], (mode.__name__, r)
    print("seed", seed, "clean", flush=True)
PY
rm -f $SO
