#!/bin/bash
# round 2, call h: deterministic training kernels
mkdir -p gpurun_out/r2h
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -s > gpurun_out/r2h/train.log 2>&1; echo "train rc=$?"
grep -E "step [12]:|passed|failed|Error|assert" gpurun_out/r2h/train.log | tail -12
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -m gpu -k "bf16 and config" > gpurun_out/r2h/fullsize1.log 2>&1; echo "fullsize rc=$?"
grep -E "full-size|passed|failed" gpurun_out/r2h/fullsize1.log | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -m gpu -k "bf16 and config" > gpurun_out/r2h/fullsize2.log 2>&1; echo "fullsize (second process) rc=$?"
grep -E "full-size|passed|failed" gpurun_out/r2h/fullsize2.log | tail -5
timeout 300 python tools/bench_train.py 2>&1 | tail -1 | tee gpurun_out/r2h/bench_train.txt
