#!/bin/bash
# sweep one environment variable on the headline frame (development): tools/sweep_env.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do echo -n "$var=$v: "; env $var=$v python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps ${REPS:-3} ${EXTRA:-} 2>&1 | tail -${TAILN:-1} | cut -c1-260; done
