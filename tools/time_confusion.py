#!/usr/bin/env python
"""CUDA-event timing of the confusion-matrix kernel at BASELINE config 2 size (16 x 1024 x 2048, K=19)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lednet_b200 as L  # noqa: F401
from lednet_b200 import ops, synth

K, N, H, W = 19, 16, 1024, 2048
dev = 'cuda'
lab = synth.make_labels(N, H, W, K, seed=200).to(torch.uint8).to(dev)
cases = {
    'noisy pred (uniform random)': torch.randint(0, K, (N, H, W), device=dev, dtype=torch.int64).to(torch.uint8),
    'pred == blocky labels': lab.clamp(max=K - 1),
}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, pred in cases.items():
    cm = torch.zeros((K + 1, K), dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.confusion_accumulate(pred, lab, K, 255, cm)
    ts = []
    for _ in range(10):
        flush.zero_()                                    # flush L2 (126 MB) between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.confusion_accumulate(pred, lab, K, 255, cm)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f'confusion {name}: median {med * 1e3:.1f} us, {2 * N * H * W / med / 1e6:.0f} GB/s (67 MB algorithmic)')
