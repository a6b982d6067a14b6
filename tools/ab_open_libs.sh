#!/bin/bash
# library variants on the open-scene configs (development): tools/ab_open_libs.sh old pk old pk
for tag in "$@"; do
  export TMPT_LIB=$PWD/toymeshpathtracer_b200/libtmpt_$tag.so
  timeout 100 python tools/exp_regen.py --scene cube --width 640 --height 360 --spp 4 --reps 8 | tail -1
  timeout 100 python tools/exp_regen.py --scene suzanne --width 640 --height 360 --spp 4 --reps 8 | tail -1
  timeout 100 python tools/exp_regen.py --scene teapot --width 1280 --height 720 --spp 16 --reps 8 | tail -1
done
