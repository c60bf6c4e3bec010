#!/usr/bin/env python
"""Opcode histogram of every compiled object of the library (cuobjdump -sass on csrc/build/*.o), per source file:
the tensor-core / TMA / TMEM mnemonics that prove which path a kernel takes (B200_PROFILING.md) plus the FP pipes.

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCOMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'SYNCS',
        'HMMA', 'IMMA', 'FFMA2', 'FFMA', 'HFMA2', 'FMUL2', 'LDG', 'STG', 'LDS', 'STS', 'LDSM', 'ATOM', 'RED', 'MATCH', 'SHFL', 'BAR']


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, 'led-net_b200', 'csrc', 'build', '*.o')))
    print('SASS opcode counts per object (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,')
    print('UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync')
    print(f'{"object":18s} ' + ' '.join(f'{k:>8s}' for k in KEYS) + '   kernels')
    for o in objs:
        out = subprocess.run(['cuobjdump', '-sass', o], capture_output=True, text=True).stdout
        cnt = collections.Counter()
        nk = 0
        for line in out.splitlines():
            if 'Function :' in line:
                nk += 1
                continue
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
            if not m:
                continue
            op = m.group(1)
            for k in KEYS:
                if op == k or (op.startswith(k) and k not in ('FFMA', 'LDS', 'BAR', 'RED', 'LDG', 'STG', 'STS')):
                    cnt[k] += 1
                    break
        print(f'{os.path.basename(o)[:-2]:18s} ' + ' '.join(f'{cnt[k]:8d}' for k in KEYS) + f'   {nk}')


if __name__ == '__main__':
    main()
