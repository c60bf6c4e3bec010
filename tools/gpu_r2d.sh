#!/bin/bash
# round 2, call d: ladder v2 (origin-0 tiling, split producers) + branch-free confusion kernel
mkdir -p gpurun_out/r2d
cd /root/repo
timeout 300 python tools/diag_ladder5.py 2>&1 | grep -v Warning | tee gpurun_out/r2d/diag5.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "confusion or tail or full_path or iou" > gpurun_out/r2d/parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/r2d/parity.log
timeout 120 python tools/time_confusion.py 2>&1 | tee gpurun_out/r2d/confusion.txt
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -m gpu > gpurun_out/r2d/fullsize.log 2>&1; echo "fullsize rc=$?"
grep -E "full-size|passed|failed|Error|assert" gpurun_out/r2d/fullsize.log | tail -12
timeout 300 python tools/diag_ladder.py 2>&1 | grep -v Warning | tee gpurun_out/r2d/probe.txt
timeout 600 python bench.py --steps 20 --warmup 5 --profile-ops > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench_ops.txt; echo "bench rc=$?"
cat gpurun_out/r2d/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac_of_per_layer_roofline'])"
grep -E "head|tail|final|stem.0 " gpurun_out/r2d/bench_ops.txt
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fullsize.py > gpurun_out/r2d/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2d/gpu_tests.log
