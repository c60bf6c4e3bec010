#!/usr/bin/env python
"""Turn an ncu metrics CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) of ONE
forward at the bench workload into profiles/dominant_kernel_traffic.json (read by bench.py for roofline.traffic).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/r1c/traffic.csv python tools/prof_forward.py --batch 16 --iters 1
    python tools/ncu_traffic.py gpurun_out/r1c/traffic.csv 'conv_tc=conv_tc|stem_tc' 16 1024 2048 profiles/r1c_traffic.csv

(`kind=regex`: the engine's op kind as bench.py names it, and the regex selecting that family's kernels.)
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

src, family, batch, height, width, committed = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
h = rows[0]
ki, mi, vi, ui, idi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('ID')
per = OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[idi], dict(name=r[ki]))
    v = float(r[vi].replace(',', ''))
    u = r[ui].lower()
    if 'byte' in u:
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    if u in ('us', 'usecond'):
        v *= 1e3
    elif u in ('ms', 'msecond'):
        v *= 1e6
    d[r[mi]] = v
kind, _, pattern = family.partition('=')
pattern = re.compile(pattern or kind)
fam = [d for d in per.values() if pattern.search(d['name'])]
tot = sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in fam)
out = dict(kernel=kind, kernel_regex=pattern.pattern, batch=batch, height=height, width=width, launches=len(fam),
           dram_bytes_per_launch=tot / max(len(fam), 1),
           dram_read_bytes_total=sum(d.get('dram__bytes_read.sum', 0) for d in fam),
           dram_write_bytes_total=sum(d.get('dram__bytes_write.sum', 0) for d in fam),
           time_ns_total=sum(d.get('gpu__time_duration.sum', 0) for d in fam),
           source=committed + ' (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one forward, per-launch rows)')
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, 'profiles', 'dominant_kernel_traffic.json'), 'w') as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
