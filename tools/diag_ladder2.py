#!/usr/bin/env python
"""Graph-mode determinism of the fused ladder with the r1 dump switched on."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import oracle
import lednet_b200 as L
from lednet_b200 import synth
from util import build_pair

K = 19
o, m = build_pair(K, dtype='bf16')
eng = m.engine()
n, h, w = 16, 1024, 2048
x = oracle.preprocess(synth.make_images_u8(n, h, w, seed=3)).cuda()
pl, lg = eng.forward_infer(x, want_logits=True)
torch.cuda.synchronize()
r1_ref = eng.debug_fetch('hx1')
del lg
os.environ['LEDB200_LADDER_DBG'] = sys.argv[1] if len(sys.argv) > 1 else '8'
bad = 0
for run in range(24):
    p = eng.forward_infer(x)
    torch.cuda.synchronize()
    d = p != pl
    if d.any():
        bad += 1
        r1 = eng.debug_fetch('hx1')
        dr = (r1 != r1_ref).any(dim=1)
        idx = d.nonzero()
        print('run', run, 'label mismatches:', int(d.sum()), 'images', idx[:, 0].unique().tolist(), 'rows', int(idx[:, 1].min()), int(idx[:, 1].max()),
              'cols', int(idx[:, 2].min()), int(idx[:, 2].max()), '| r1 pixels differing:', int(dr.sum()))
        if dr.any():
            j = dr.nonzero()
            print('    r1 diff rows', int(j[:, 1].min()), int(j[:, 1].max()), 'cols', int(j[:, 2].min()), int(j[:, 2].max()))
            i = j[0]
            print('    r1 got', [round(v, 3) for v in r1[i[0], :, i[1], i[2]].tolist()[:10]])
            print('    r1 ref', [round(v, 3) for v in r1_ref[i[0], :, i[1], i[2]].tolist()[:10]])
print('bad runs:', bad, 'of 24')
