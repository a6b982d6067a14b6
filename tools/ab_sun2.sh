#!/bin/bash
# sun grid resolution on the headline frame and scene build times (development)
for n in "$@"; do
  export TMPT_SUN_GRID=$n; [ "$n" = default ] && unset TMPT_SUN_GRID
  echo "== TMPT_SUN_GRID=$n"
  timeout 120 python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 3 2>&1 | tail -1
  timeout 120 python tools/exp_regen.py --scene teapot --width 1280 --height 720 --spp 16 --reps 6 2>&1 | tail -1
  timeout 120 python tools/exp_buildtime.py 2>&1 | tail -6
done
