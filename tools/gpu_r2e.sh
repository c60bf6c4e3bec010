#!/bin/bash
# round 2, call e: SyncBN two-rank check on one GPU (gloo), train tests
mkdir -p gpurun_out/r2e
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu -s > gpurun_out/r2e/train.log 2>&1; echo "train rc=$?"
grep -E "syncbn|passed|failed|Error" gpurun_out/r2e/train.log | tail -12
timeout 300 python tools/bench_train.py 2>&1 | tail -5 | tee gpurun_out/r2e/bench_train.txt
