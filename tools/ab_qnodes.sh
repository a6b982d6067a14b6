#!/bin/bash
# Quantised 4-wide nodes (64 B, FFMA2 decode) against the float nodes, headline frame + config 4, then one ncu --set full of the
# quantised kernel (development; libraries from tools/build_variants.py q4=-DTMPT_QNODES=1 q5=...,-DTMPT_QSTRIDE=5 q4t=...,-DTMPT_TUNE_CFG=1).
L=toymeshpathtracer_b200
for tag in base q4 q5; do
  lib=$PWD/$L/libtmpt_$tag.so; [ $tag = base ] && lib=$PWD/$L/libtmpt.so
  TMPT_LIB=$lib timeout 120 python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 3 2>&1 | tail -1
  TMPT_LIB=$lib timeout 120 python tools/exp_regen.py --scene sponza --width 640 --height 360 --spp 4 --reps 5 2>&1 | tail -1
done
for cfg in 4 7; do
  TMPT_RENDER_CFG=$cfg TMPT_LIB=$PWD/$L/libtmpt_q4t.so timeout 120 python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 3 2>&1 | tail -1
done
TMPT_LIB=$PWD/$L/libtmpt_q4.so timeout 120 python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 0 --stats 2>&1 | tail -1
TMPT_LIB=$PWD/$L/libtmpt_q4.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -c 1 -o gpurun_out/prof_render_q4 -f \
  python tools/exp_regen.py --scene sponza --width 1920 --height 1080 --spp 64 --reps 0 > gpurun_out/prof_render_q4.log 2>&1
echo "ncu rc $?"
