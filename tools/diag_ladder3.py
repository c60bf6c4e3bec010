#!/usr/bin/env python
"""Fresh-process, graph-mode: the first forwards at batch 16 1024x2048; which of xc / r2 / labels differ from later runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import oracle
import lednet_b200 as L
from lednet_b200 import synth
from util import build_pair

o, m = build_pair(19, dtype='bf16')
eng = m.engine()
n, h, w = 16, 1024, 2048
x = oracle.preprocess(synth.make_images_u8(n, h, w, seed=3)).cuda()
runs = []
for run in range(6):
    p = eng.forward_infer(x).clone()
    torch.cuda.synchronize()
    runs.append((p.cpu(), eng.debug_fetch('xc'), eng.debug_fetch('hx2'), eng.debug_fetch('x2h')[:, :, -4:, :8].clone()))
ref = runs[-1]
for i, r in enumerate(runs[:-1]):
    out = []
    for name, a, b in zip(('labels', 'xc', 'r2', 'x2h(bottom-left)'), r, ref):
        d = (a != b)
        if d.dim() == 4:
            d = d.any(dim=1)
        s = f'{name}: {int(d.sum())}'
        if d.any():
            idx = d.nonzero()
            s += f' imgs {idx[:, 0].unique().tolist()} rows {int(idx[:, 1].min())}-{int(idx[:, 1].max())} cols {int(idx[:, 2].min())}-{int(idx[:, 2].max())}'
        out.append(s)
    print('run', i, '|', ' | '.join(out))
