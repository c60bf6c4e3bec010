#!/bin/bash
mkdir -p gpurun_out/r2q
cd /root/repo
timeout 600 ncu --kernel-name regex:"getb|conv_tc|mfaf|seam" --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/r2q/led_ncu.csv python tools/time_led_blocks.py > gpurun_out/r2q/b.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2q/led_ncu.csv')) if len(r)>10]
hdr=rows[0]; agg=collections.OrderedDict()
for r in rows[1:]:
    d=dict(zip(hdr,r))
    k=d['Kernel Name'][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(d['Metric Value'].replace(',',''))/1e3
for k,(n,us) in agg.items(): print(f'{k:62s} {n:3d} launches {us:9.1f} us')
PY
