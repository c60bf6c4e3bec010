#!/bin/bash
# round 2, call i: fused DAPPM
mkdir -p gpurun_out/r2i
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_dappm.py -x -q -m gpu -s > gpurun_out/r2i/dappm.log 2>&1; echo "dappm rc=$?"
grep -E "DAPPM|passed|failed|Error|assert|err" gpurun_out/r2i/dappm.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --profile-ops --no-extras > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench_ops.txt; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2i/bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac_of_per_layer_roofline'], d['gpu_launches'])"
grep -E "spp|final|2.0.conv3" gpurun_out/r2i/bench_ops.txt
tail -3 gpurun_out/r2i/bench_ops.txt
