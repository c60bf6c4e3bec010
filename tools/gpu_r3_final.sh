#!/bin/bash
# Last visit of the round: GPU tests, smoke, the bench line (with extras) and one full ncu capture of each training kernel
# that is not a convolution (those are in gpu_r3.sh).
TAG=${1:-r3zf}
O=gpurun_out/$TAG
mkdir -p $O
cd /root/repo
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 ) > $O/pytest_gpu.log; cat $O/pytest_gpu.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -3 ) > $O/smoke.log; cat $O/smoke.log
( timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ); echo "bench rc=$?"
( timeout 300 python tools/bench_train.py --steps 5 --warmup 2 > $O/train_tf32x3.json 2> $O/train.err )
for k in ohem_up_bwd_tiled ohem_up_pixel_tiled bn_bwd_apply4 "chan_reduce4_kernel<1>" bn_apply4 resize_bwd_row wgrad_small_cin stem_fwd; do
  timeout 300 ncu --set full --clock-control none -k regex:"$k" -c 1 -o $O/k_$(echo $k | tr -c 'a-zA-Z0-9_' '_') \
     python tools/bench_train.py --steps 1 --warmup 0 > $O/ncu_k.log 2>&1
done
ls $O | head -40
cut -c1-300 $O/bench.json; echo; cut -c1-250 $O/train_tf32x3.json; echo
