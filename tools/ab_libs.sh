#!/bin/bash
# A/B of library variants on the headline frame (development): tools/ab_libs.sh base s12 flat ...   ("base" = libtmpt.so)
# Equal sha across variants = byte-identical frames.
for tag in "$@"; do
  lib=toymeshpathtracer_b200/libtmpt_$tag.so; [ "$tag" = base ] && lib=toymeshpathtracer_b200/libtmpt.so
  TMPT_LIB=$PWD/$lib python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps ${REPS:-3} ${EXTRA:-} 2>&1 | tail -${TAILN:-1}
done
