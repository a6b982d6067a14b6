#!/bin/bash
# A/B of (library variant, TMPT_RENDER_KERNEL) pairs on the headline frame (development): tools/ab_env_libs.sh tag:k tag:k ...
for spec in "$@"; do
  tag=${spec%%:*}; k=${spec##*:}
  lib=toymeshpathtracer_b200/libtmpt_$tag.so; [ "$tag" = base ] && lib=toymeshpathtracer_b200/libtmpt.so
  if [ "$k" = 0 ]; then TMPT_LIB=$PWD/$lib python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps ${REPS:-2} 2>&1 | tail -1
  else TMPT_RENDER_KERNEL=$k TMPT_LIB=$PWD/$lib python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps ${REPS:-2} 2>&1 | tail -1; fi
done
