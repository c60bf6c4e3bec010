#!/bin/bash
# ncu source-level stall sampling of three conv_tc launches at batch 16 (steady state: 27 / 443 tiles per SM)
O=gpurun_out/${1:-ncu_epi}
mkdir -p $O
timeout 500 ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 -k regex:conv_tc -s 15 -c 2 \
   -o $O/spa python tools/prof_forward.py --batch 16 --iters 1 > $O/spa.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 -k regex:conv_tc -s 56 -c 1 \
   -o $O/hx1 python tools/prof_forward.py --batch 16 --iters 1 > $O/hx1.log 2>&1
ls -la $O; tail -3 $O/spa.log $O/hx1.log
