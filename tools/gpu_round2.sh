#!/bin/bash
# One GPU-box visit: parity tests (all), smoke, bench + reference arm, train bench, ncu launch list + full captures.
TAG=${1:-r1b}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > $O/smoke.log
( timeout 400 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench_ops.txt )
( timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err )
( timeout 300 python tools/bench_train.py --batch 4 --size 512 --steps 3 --warmup 1 > $O/bench_train_small.json 2> $O/bench_train_small.err )
( timeout 400 python tools/bench_train.py --batch 12 --size 1024 --steps 3 --warmup 1 > $O/bench_train.json 2> $O/bench_train.err )
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $O/launches.csv python tools/prof_forward.py --batch 4 --iters 2 > $O/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 12 \
   -o $O/conv_tc python tools/prof_forward.py --batch 4 --iters 2 > $O/ncu_conv_tc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tail|stem_tc|upsample|confusion|avgpool' -s 6 -c 8 \
   -o $O/others python tools/prof_forward.py --batch 4 --iters 2 > $O/ncu_others.log 2>&1
ls -la $O
tail -12 $O/pytest_gpu.log; cat $O/smoke.log; cat $O/bench.json; cat $O/bench_train_small.json $O/bench_train.json; tail -3 $O/bench_train.err
