#!/bin/bash
# ncu source-level stall sampling of the two 64->64 3x3 conv_tc launches (spatial_branch_layers.0.0) at batch 16
O=gpurun_out/${1:-ncu_epi}
mkdir -p $O
timeout 500 ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 -k regex:conv_tc -s 15 -c 1 \
   -o $O/spa python tools/prof_forward.py --batch 16 --iters 1 > $O/spa.log 2>&1
ls -la $O
