#!/bin/bash
# round 2, call j: which DAPPM kernel takes the time
mkdir -p gpurun_out/r2j
cd /root/repo
LEDB200_NO_GRAPH=1 timeout 600 ncu --kernel-name regex:dappm --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,lts__t_bytes.sum --clock-control none -c 6 --csv --log-file gpurun_out/r2j/dappm_ncu.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2j/b.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j/dappm_ncu.csv')) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r))
    print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
