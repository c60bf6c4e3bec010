#!/bin/bash
# GPU visit for the training path: parity tests + train-step bench.
TAG=${1:-r1t}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -40 ) > $O/pytest_train.log
cat $O/pytest_train.log | tail -30
( timeout 300 python tools/bench_train.py --batch 4 --size 512 --steps 3 --warmup 1 > $O/bench_train_small.json 2> $O/bench_train_small.err )
cat $O/bench_train_small.json; tail -5 $O/bench_train_small.err
( timeout 600 python tools/bench_train.py --batch 12 --size 1024 --steps 3 --warmup 1 > $O/bench_train.json 2> $O/bench_train.err )
cat $O/bench_train.json; tail -5 $O/bench_train.err
