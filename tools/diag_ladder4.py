#!/usr/bin/env python
"""Reproducer: back-to-back graph-mode forwards (no sync in between) at batch 16 1024x2048."""
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
nrun = 40
preds = [eng.forward_infer(x).clone() for _ in range(nrun)]
torch.cuda.synchronize()
dbg = os.environ.pop('LEDB200_LADDER_DBG', None)
os.environ['LEDB200_LADDER_DBG'] = '0'
pl, lg = eng.forward_infer(x, want_logits=True)
bad = 0
for i, p in enumerate(preds):
    d = p != pl
    if d.any():
        bad += 1
        idx = d.nonzero()
        if bad <= 6:
            print('  run', i, 'mismatches', int(d.sum()), 'imgs', idx[:, 0].unique().tolist(), 'rows', int(idx[:, 1].min()), int(idx[:, 1].max()),
                  'cols', int(idx[:, 2].min()), int(idx[:, 2].max()))
print(f'dbg={dbg}: bad runs {bad} of {nrun}')
