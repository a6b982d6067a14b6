#!/bin/bash
# open scenes: lockstep (0), per-lane ray regeneration (1, 3: k_render_regen), path regeneration (6, 8: k_render_paths); experiments build
export TMPT_LIB=$PWD/toymeshpathtracer_b200/libtmpt_exp.so
for k in "$@"; do
  TMPT_RENDER_KERNEL=$k timeout 100 python tools/exp_regen.py --scene cube --width 640 --height 360 --spp 4 --reps 6 | tail -1
  TMPT_RENDER_KERNEL=$k timeout 100 python tools/exp_regen.py --scene suzanne --width 640 --height 360 --spp 4 --reps 6 | tail -1
  TMPT_RENDER_KERNEL=$k timeout 100 python tools/exp_regen.py --scene teapot --width 1280 --height 720 --spp 16 --reps 6 | tail -1
done
