#!/bin/bash
# Build led-net_b200/libledb200_tl.so: the product objects with conv_tc.cu recompiled under -DLEDB_TC_TIMELINE
# (clock64 stamps of CTA 0's ramp, read back by tools/probes/tc_timeline.py).  Not part of the product build.
set -e
cd "$(dirname "$0")/../.."
python led-net_b200/build.py > /dev/null
B=led-net_b200/csrc/build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  --expt-relaxed-constexpr -DLEDB_TC_TIMELINE -c led-net_b200/csrc/conv_tc.cu -o $B/conv_tc_tl.o
OBJS=$(ls $B/*.cu.o | grep -v conv_tc.cu.o)
/usr/local/cuda/bin/nvcc -shared -o led-net_b200/libledb200_tl.so $OBJS $B/conv_tc_tl.o \
  -gencode arch=compute_100a,code=sm_100a -lcudart
echo led-net_b200/libledb200_tl.so
