#!/bin/bash
# Full GPU visit (round-end evidence): parity tests, smoke, bench + reference arm, train bench, the ncu launch list of the
# bench command, per-launch DRAM traffic of one forward at the bench workload, full ncu captures of the top kernels.
TAG=${1:-r1c}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > $O/smoke.log
( timeout 400 python bench.py --steps 20 --warmup 5 --profile-ops > $O/bench.json 2> $O/bench_ops.txt )
( timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err )
( timeout 300 python tools/bench_train.py --batch 12 --size 1024 --steps 3 --warmup 1 > $O/bench_train.json 2> $O/bench_train.err )
# launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches.log 2>&1
# DRAM bytes per launch, one forward at the bench workload (batch 16)
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/traffic.csv python tools/prof_forward.py --batch 16 --iters 1 > $O/traffic.log 2>&1
# full captures: a 64->64 3x3 pair (launches 15, 16 of conv_tc), the two head convs, tail / stem / upsample / confusion
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 15 -c 2 \
   -o $O/conv_tc_spa python tools/prof_forward.py --batch 16 --iters 1 > $O/ncu_spa.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 54 -c 3 \
   -o $O/conv_tc_head python tools/prof_forward.py --batch 16 --iters 1 > $O/ncu_head.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:'tail3|tail2|stem_tc|upsample_add16|confusion' -c 6 \
   -o $O/others python tools/prof_forward.py --batch 16 --iters 1 > $O/ncu_others.log 2>&1
ls -la $O
tail -4 $O/pytest_gpu.log; cat $O/smoke.log; cut -c1-600 $O/bench.json; cat $O/bench_ref.json | cut -c1-300; cat $O/bench_train.json | cut -c1-400
for f in launches traffic ncu_spa ncu_head ncu_others; do tail -n 2 $O/$f.log; done
