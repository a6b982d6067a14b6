#!/bin/bash
# render-kernel variants on small frames (hash must agree) and on the headline frame (development).  The regeneration kernel lives in
# the experiments build only: python tools/build_variants.py exp=-DTMPT_EXPERIMENTS=1 first.   tools/ab_regen.sh 0 1 2 3 4
export TMPT_LIB=$PWD/toymeshpathtracer_b200/libtmpt_exp.so
for k in "$@"; do
  TMPT_RENDER_KERNEL=$k python tools/exp_regen.py --scene suzanne --spp 4
  TMPT_RENDER_KERNEL=$k python tools/exp_regen.py --scene sponza --width 645 --height 363 --spp 20
  TMPT_RENDER_KERNEL=$k python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 1
done
