#!/bin/bash
# round 2, call f: boundary tests (stack_pad / slide_merge / binary head), SyncBN, bench with extras, label uniformity
mkdir -p gpurun_out/r2f
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r2f/boundary.log 2>&1; echo "boundary rc=$?"
tail -15 gpurun_out/r2f/boundary.log
timeout 200 python tools/label_uniformity.py 2>&1 | grep -v Warn | tee gpurun_out/r2f/uniform.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2f/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f/bench.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('sustained'), json.dumps(d.get('extra'))[:1500])
PY
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2f/bench_ref.json 2>/dev/null; cat gpurun_out/r2f/bench_ref.json | cut -c1-300
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_boundary.py --deselect tests/test_gpu_train.py > gpurun_out/r2f/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r2f/gpu_tests.log
