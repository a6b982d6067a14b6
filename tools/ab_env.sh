#!/bin/bash
# A/B of environment-selected variants on the headline frame (development): tools/ab_env.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  echo -n "$var=$v: "
  env $var=$v python bench.py --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frame %8.1f Mrays/s  %7.1f ms' % (d['value'], d['ms_per_step']))"
done
