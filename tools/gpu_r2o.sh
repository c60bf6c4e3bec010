#!/bin/bash
# round 2, call o: stem conv software pipeline
mkdir -p gpurun_out/r2o
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv_tc.py -x -q -m gpu > gpurun_out/r2o/parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/r2o/parity.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-ops --no-extras > gpurun_out/r2o/bench.json 2> gpurun_out/r2o/bench_ops.txt; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2o/bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac_of_per_layer_roofline'], d['gpu_launches'])"
grep -E "stem.0 |stem.1 " gpurun_out/r2o/bench_ops.txt
