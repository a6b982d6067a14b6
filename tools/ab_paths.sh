#!/bin/bash
# the path-regeneration render kernel against k_render on BASELINE's scenes (development; experiments build): tools/ab_paths.sh K K ...
# every run under its own timeout: a kernel that does not terminate must not take the GPU box with it
export TMPT_LIB=$PWD/toymeshpathtracer_b200/libtmpt_exp.so
for k in "$@"; do
  TMPT_RENDER_KERNEL=$k timeout 60 python tools/exp_regen.py --scene suzanne --width 64 --height 36 --spp 3 --reps 0 2>&1 | tail -1 || { echo "kernel $k: small frame failed or timed out"; continue; }
  for cfg in "cube 640 360 4" "suzanne 640 360 4" "teapot 1280 720 16" "sponza 640 360 4" "sponza 1920 1080 64"; do
    set -- $cfg
    TMPT_RENDER_KERNEL=$k timeout 120 python tools/exp_regen.py --scene $1 --width $2 --height $3 --spp $4 --reps 3 2>&1 | tail -1
  done
done
