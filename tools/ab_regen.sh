#!/bin/bash
# render-kernel variants on small frames (hash must agree) and on the headline frame (development)
for k in "$@"; do
  TMPT_RENDER_KERNEL=$k python tools/exp_regen.py --scene suzanne --spp 4
  TMPT_RENDER_KERNEL=$k python tools/exp_regen.py --scene sponza --width 645 --height 363 --spp 20
  TMPT_RENDER_KERNEL=$k python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 1
done
