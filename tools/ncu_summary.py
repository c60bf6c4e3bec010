#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) or a launch-list CSV into a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/r1a/conv_tc.ncu-rep > profiles/r1a_conv_tc_ncu.txt
    python tools/ncu_summary.py --launches gpurun_out/r1a/launches.csv > profiles/r1a_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

COLS = OrderedDict([
    ('gpu__time_duration.sum', 'us'),
    ('dram__bytes_read.sum', 'rdMB'),
    ('dram__bytes_write.sum', 'wrMB'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('lts__t_sector_hit_rate.pct', 'l2hit%'),
])


def rows_of(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def num(s):
    try:
        return float(s.replace(',', ''))
    except Exception:
        return float('nan')


def main():
    if sys.argv[1] == '--launches':
        rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
        h = rows[0]
        kn, mv = h.index('Kernel Name'), h.index('Metric Value')
        agg = OrderedDict()
        for r in rows[1:]:
            name = r[kn].split('(')[0].replace('void ', '').replace('ledb::<unnamed>::', '')[:70]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += num(r[mv])
        tot = sum(a[1] for a in agg.values())
        unit = rows[1][h.index('Metric Unit')]
        print(f'# ncu launch list (gpu__time_duration.sum, cold-cache, serialised): shares matter, not absolutes\n'
              f'# total {tot:.1f} {unit} over {sum(a[0] for a in agg.values())} launches')
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f'{100 * t / tot:6.2f}%  {t:12.1f} {unit}  x{n:<4d} {name}')
        return
    h, units, rows = rows_of(sys.argv[1])
    kn = h.index('Kernel Name')
    idx = [(h.index(c), lab) for c, lab in COLS.items() if c in h]
    print('# ' + ' '.join(f'{lab:>8s}' for _, lab in idx) + '  kernel')
    for r in rows:
        vals = []
        for i, lab in idx:
            v = num(r[i])
            u = units[i]
            if lab in ('rdMB', 'wrMB'):
                v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
            if lab == 'us':
                v = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0)
            vals.append(f'{v:8.1f}')
        name = r[kn].replace('void ', '').replace('ledb::<unnamed>::', '')[:90]
        print('  ' + ' '.join(vals) + '  ' + name)


if __name__ == '__main__':
    main()
