#!/usr/bin/env python
"""Per-layer timing of LEDNet(variant='led') on the headline shape (batch 16, 1024x2048) and on config 5's (batch 128,
512x512): CUDA events around every C-ABI call of the composed trunk, grouped by block kind (STDC convs, GETB, MFAF, SEAM,
DAPPM, glue), with each block's algorithmic bytes (input + output once, bf16) against the measured HBM peak.

    python tools/time_led_blocks.py > profiles/r2_led_blocks.txt
"""
import collections
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lednet_b200 as L
from lednet_b200 import synth
from lednet_b200.led_variant import LEDTrunk


def main():
    peak = 6520.5
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        peak = json.load(open(p))['hbm_gbs']
    dev = torch.device('cuda')
    m = L.LEDNet(variant='led').eval()
    sd = synth.make_state_dict(m.state_dict(), seed=2)
    sd['fusion_kernel'] = m.state_dict()['fusion_kernel'].clone()
    m.load_state_dict(sd)
    m.to(dev).set_compute_dtype('bf16')
    records = []

    def timed(kind, fn):
        def wrap(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            ins = [t for t in a if torch.is_tensor(t)]
            nbytes = sum(t.numel() * t.element_size() for t in ins) + (out.numel() * out.element_size() if torch.is_tensor(out) else 0)
            records.append((kind(a, k) if callable(kind) else kind, e0, e1, nbytes, tuple(out.shape) if torch.is_tensor(out) else ()))
            return out
        return wrap

    m._conv = timed(lambda a, k: 'conv ' + a[0].split('.')[0] + ('.dappm' if a[0].startswith('spp') else ''), m._conv)
    LEDTrunk._avgpool = staticmethod(timed('glue avgpool', LEDTrunk._avgpool))
    LEDTrunk._resize = staticmethod(timed('glue resize', LEDTrunk._resize))
    LEDTrunk._add = staticmethod(timed('glue add/relu', LEDTrunk._add))
    orig_block = LEDTrunk._block

    def block(mod, *xs):
        return timed(type(mod).__name__, lambda *t: orig_block(mod, *t))(*xs)
    LEDTrunk._block = staticmethod(block)

    for (n, h, w) in [(16, 1024, 2048), (128, 512, 512)]:
        x = torch.randn(n, 3, h, w, device=dev)
        for _ in range(2):
            records.clear()
            m(x)
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        rows = []
        for kind, e0, e1, nbytes, shape in records:
            ms = e0.elapsed_time(e1)
            a = agg.setdefault(kind, [0.0, 0, 0])
            a[0] += ms; a[1] += 1; a[2] += nbytes
            if kind in ('GETBBlock', 'Muti_AFF', 'SEAM'):
                rows.append((kind, shape, ms, nbytes))
        total = sum(v[0] for v in agg.values())
        print(f'== LEDNet(variant=led) trunk, batch {n} x {h}x{w}, bf16: {total:.2f} ms over {len(records)} calls '
              f'(event-timed per call: includes launch gaps)')
        for kind, (ms, cnt, nbytes) in agg.items():
            print(f'  {kind:28s} {cnt:3d} calls {ms:8.3f} ms  {100 * ms / total:5.1f}%   in+out {nbytes / 1e6:9.1f} MB  '
                  f'-> {nbytes / ms / 1e6:7.1f} GB/s = {100 * nbytes / ms / 1e6 / peak:5.1f}% of {peak:.0f} GB/s')
        for kind, shape, ms, nbytes in rows:
            print(f'    {kind:10s} out {str(shape):22s} {ms:7.3f} ms  in+out {nbytes / 1e6:7.1f} MB  {100 * nbytes / ms / 1e6 / peak:5.1f}% of HBM peak')


if __name__ == '__main__':
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        main()
