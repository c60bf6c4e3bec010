#!/bin/bash
# the driver's round-end sequence: all GPU tests, smoke, bench
mkdir -p gpurun_out/all
cd /root/repo
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/all/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/all/gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -2 | tee gpurun_out/all/smoke.log
