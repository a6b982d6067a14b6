#!/bin/bash
# A/B of launch configurations of one library on the headline frame (development): tools/ab_cfg.sh <tag> cfg cfg ...
tag=$1; shift
lib=toymeshpathtracer_b200/libtmpt_$tag.so; [ "$tag" = base ] && lib=toymeshpathtracer_b200/libtmpt.so
for c in "$@"; do echo -n "cfg $c: "; TMPT_RENDER_CFG=$c TMPT_LIB=$PWD/$lib python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps ${REPS:-3} 2>&1 | tail -1; done
