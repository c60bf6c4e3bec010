#!/bin/bash
# round 2, call g: final-rung uniform-corner shortcut
mkdir -p gpurun_out/r2g
cd /root/repo
timeout 300 python tools/diag_ladder.py 2>&1 | grep -v Warning | tee gpurun_out/r2g/probe.txt
LADDER_BASE_DBG=64 timeout 300 python tools/diag_ladder.py 2>&1 | grep -v Warning | grep "dbg  0" | sed 's/^/no shortcut: /' | tee -a gpurun_out/r2g/probe.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail or full_path or ladder or fuse" > gpurun_out/r2g/parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/r2g/parity.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -m gpu > gpurun_out/r2g/fullsize.log 2>&1; echo "fullsize rc=$?"
grep -E "full-size|passed|failed|Error|assert" gpurun_out/r2g/fullsize.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --profile-ops --no-extras > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench_ops.txt; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2g/bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac_of_per_layer_roofline'])"
grep -E "head|tail|final" gpurun_out/r2g/bench_ops.txt
