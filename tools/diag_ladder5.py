#!/usr/bin/env python
"""Reproducer 2: small shape first, then batch 16 1024x2048 (arena re-allocated); poisoned pred buffers."""
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
junk = [torch.full((1 << 28,), 0x7f, dtype=torch.uint8, device='cuda') for _ in range(8)]     # dirty 2 GB, then free it
del junk
torch.cuda.empty_cache()
for (n, h, w) in [(2, 128, 256), (4, 512, 1024), (16, 1024, 2048)]:
    x = oracle.preprocess(synth.make_images_u8(n, h, w, seed=3)).cuda()
    preds = []
    for _ in range(8):
        p = torch.full((n, h, w), 255, dtype=torch.uint8, device='cuda')
        eng.forward_infer(x, pred=p)
        preds.append(p)
    torch.cuda.synchronize()
    pl, lg = eng.forward_infer(x, want_logits=True)
    bad = 0
    for i, p in enumerate(preds):
        d = p != pl
        if d.any():
            bad += 1
            idx = d.nonzero()
            print('  ', (n, h, w), 'run', i, 'mismatches', int(d.sum()), 'unwritten(255):', int((p == 255).sum()), 'imgs', idx[:, 0].unique().tolist(),
                  'rows', int(idx[:, 1].min()), int(idx[:, 1].max()), 'cols', int(idx[:, 2].min()), int(idx[:, 2].max()), 'vals', p[d][:8].tolist(), 'ref', pl[d][:8].tolist())
    print((n, h, w), f'bad runs {bad} of 8  (poison={os.environ.get("LEDB200_POISON")}, dbg={os.environ.get("LEDB200_LADDER_DBG")})')
    del lg
