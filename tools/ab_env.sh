#!/bin/bash
# A/B of environment-selected variants on the headline frame (development): tools/ab_env.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  echo -n "$var=$v: "
  env $var=$v python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 1 | sed -e 's/.*rays [0-9]* *//'
done
