"""Sum an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name:  python tools/summarize_launches.py file.csv [steps]"""
import collections
import csv
import re
import sys

steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'ledb::<unnamed>::|void |ledb::', '', name)
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u in ('nsecond', 'ns') else (v * 1e3 if u in ('msecond', 'ms') else v)
    tot[name][0] += 1
    tot[name][1] += v
S = sum(v[1] for v in tot.values())
print(f'total {S / steps / 1e3:.2f} ms per step, {sum(v[0] for v in tot.values()) / steps:.0f} launches per step')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{v[1] / steps / 1e3:9.3f} ms {v[0] / steps:7.1f}  {100 * v[1] / S:5.1f}%  {k[:120]}')
