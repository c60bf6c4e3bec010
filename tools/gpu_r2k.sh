#!/bin/bash
# round 2, call k: fused DAPPM v2 (decoupled producers, 16 B pooled loads)
mkdir -p gpurun_out/r2k
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_dappm.py -x -q -m gpu -s > gpurun_out/r2k/dappm.log 2>&1; echo "dappm rc=$?"
grep -E "DAPPM|passed|failed|Error|assert|err" gpurun_out/r2k/dappm.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --profile-ops --no-extras > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench_ops.txt; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2k/bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac_of_per_layer_roofline'], d['gpu_launches'])"
grep -E "spp|final" gpurun_out/r2k/bench_ops.txt
LEDB200_NO_GRAPH=1 timeout 600 ncu --kernel-name regex:dappm --metrics gpu__time_duration.sum,smsp__inst_executed.sum --cache-control none --clock-control none -c 6 --csv --log-file gpurun_out/r2k/dappm_ncu.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2k/b.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2k/dappm_ncu.csv')) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r))
    print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
