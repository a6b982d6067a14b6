#!/bin/bash
# ncu --set full of the render kernel: lockstep (0), regen (1), and lockstep with quantised nodes
ARGS="--scene sponza --width 645 --height 363 --spp 20 --reps 0"
TMPT_RENDER_KERNEL=0 ncu --set full --clock-control none --import-source on -k regex:k_render -c 1 -o gpurun_out/cmp_k0 -f python tools/exp_regen.py $ARGS > gpurun_out/cmp_k0.log 2>&1
TMPT_RENDER_KERNEL=1 ncu --set full --clock-control none --import-source on -k regex:k_render -c 1 -o gpurun_out/cmp_k1 -f python tools/exp_regen.py $ARGS > gpurun_out/cmp_k1.log 2>&1
cp toymeshpathtracer_b200/libtmpt.so /tmp/saved.so; cp toymeshpathtracer_b200/build/ab/q5.so toymeshpathtracer_b200/libtmpt.so
TMPT_RENDER_KERNEL=0 ncu --set full --clock-control none --import-source on -k regex:k_render -c 1 -o gpurun_out/cmp_q5 -f python tools/exp_regen.py $ARGS > gpurun_out/cmp_q5.log 2>&1
cp /tmp/saved.so toymeshpathtracer_b200/libtmpt.so
tail -2 gpurun_out/cmp_k0.log gpurun_out/cmp_k1.log gpurun_out/cmp_q5.log
