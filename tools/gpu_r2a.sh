#!/bin/bash
# round 2, call a: baseline test status + trained-weight agreement + full-size parity on the unchanged engine
mkdir -p gpurun_out/r2a
cd /root/repo
python - > gpurun_out/r2a/train.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, 'tests')
import torch
import trained, test_gpu_fullsize as F
t0 = time.time()
sd = trained.train_product(19, F.SCHED19, steps=300, verbose=True)
print('train19 s', time.time() - t0)
t0 = time.time()
sd = trained.train_product(2, F.SCHED2, steps=300, verbose=True)
print('train2 s', time.time() - t0)
PY
tail -5 gpurun_out/r2a/train.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -s -m gpu > gpurun_out/r2a/fullsize.log 2>&1; echo "fullsize rc=$?"
tail -30 gpurun_out/r2a/fullsize.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fullsize.py > gpurun_out/r2a/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -8 gpurun_out/r2a/gpu_tests.log
