#!/bin/bash
# tuning sweeps (development): render launch configuration, sponza stand-in 1080p 64spp
run() { python bench.py --steps 1 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%8.1f Mrays/s  %7.1f ms' % (d['value'], d['ms_per_step']))"; }
for c in "$@"; do echo -n "RENDER_CFG=$c: "; TMPT_RENDER_CFG=$c run; done
