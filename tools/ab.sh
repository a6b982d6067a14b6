#!/bin/bash
# A/B two builds of the library on the traversal experiment and the headline frame (development)
for lib in "$@"; do
  echo "== $lib"; cp toymeshpathtracer_b200/libtmpt.so /tmp/libtmpt_saved.so; [ "$lib" != "base" ] && cp "$lib" toymeshpathtracer_b200/libtmpt.so
  python tools/exp_traverse.py --bounces 2 2>&1 | grep -E "bounce [12]|shadow|TOTAL"
  python bench.py --steps 1 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frame %8.1f Mrays/s  %7.1f ms' % (d['value'], d['ms_per_step']))"
  cp /tmp/libtmpt_saved.so toymeshpathtracer_b200/libtmpt.so
done
