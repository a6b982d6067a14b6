#!/bin/bash
# csrc/host.cpp (OBJ loader, camera, PNG writer, command line: no CUDA in it) built with AddressSanitizer + UBSan and fed hostile input.
#   tools/host_fuzz/run.sh [seed] [files]      (default 1 3000)
# Round 2: found two signed-integer overflows the reference's parser has too (an exponent of ten digits; "-2147483648" as an index);
# host.cpp now does that arithmetic unsigned -- same bits, defined behaviour -- and 3000 files pass clean.
set -e
cd "$(dirname "$0")/../.."
SO=/tmp/libhost_asan_$$.so
g++ -std=c++17 -O1 -g -fPIC -shared -ffp-contract=off -fsanitize=address,undefined -fno-sanitize-recover=undefined -o $SO \
    toymeshpathtracer_b200/csrc/host.cpp tools/host_fuzz/stubs.cpp
LD_PRELOAD=$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
    UBSAN_OPTIONS=print_stacktrace=1 python tools/host_fuzz/fuzz_host.py "${1:-1}" "${2:-3000}" $SO
rm -f $SO
