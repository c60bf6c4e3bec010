#!/usr/bin/env python
"""Diagnostic: run-to-run determinism of the fused ladder, its r1 against the rung kernel's, and probe timings."""
import os
import sys

os.environ['LEDB200_NO_GRAPH'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import oracle
import lednet_b200 as L
from lednet_b200 import synth
from util import build_pair

K = int(os.environ.get('K', 19))
BASE = int(os.environ.get('LADDER_BASE_DBG', 0))      # e.g. 64: final rung without the uniform-corner shortcut
o, m = build_pair(K, dtype='bf16')
eng = m.engine()
n, h, w = 16, 1024, 2048
x = oracle.preprocess(synth.make_images_u8(n, h, w, seed=3)).cuda()
pl, lg = eng.forward_infer(x, want_logits=True)
r1_ref = eng.debug_fetch('hx1')                       # rung-mode r1 (fp16 -> fp32 NCHW, host)
os.environ['LEDB200_LADDER_DBG'] = str(8 | BASE)
for run in range(6):
    p = eng.forward_infer(x)
    torch.cuda.synchronize()
    r1 = eng.debug_fetch('hx1')
    d = p != pl
    dr = (r1 != r1_ref).any(dim=1)
    print('run', run, 'label mismatches vs rung+tail2:', int(d.sum()), ' r1 pixels differing:', int(dr.sum()))
    if d.any():
        print('   labels first', d.nonzero()[:4].tolist())
    if dr.any():
        idx = dr.nonzero()
        print('   r1 first', idx[:6].tolist(), 'count by image', torch.bincount(idx[:, 0], minlength=n).tolist())
        i = idx[0]
        print('   r1 got', r1[i[0], :, i[1], i[2]].tolist()[:8], 'ref', r1_ref[i[0], :, i[1], i[2]].tolist()[:8])
os.environ['LEDB200_LADDER_DBG'] = '0'
# ---- probe timings of head_x1 (final mode): which role bounds the kernel
for dbg, what in [(0, 'everything on'), (1, 'no output phase'), (2, 'no up gather'), (3, 'no output phase, no up gather'),
                  (4, 'no MMAs'), (7, 'all off')]:
    os.environ['LEDB200_LADDER_DBG'] = str(dbg | BASE)
    eng.forward_infer(x)
    prof = dict(eng.profile_ops(iters=5))
    print(f'dbg {dbg:2d} {what:32s} head_x1 {prof["decode_head.head_x1"] * 1e3:7.1f} us   head_x2 {prof["decode_head.head_x2"] * 1e3:7.1f} us')
