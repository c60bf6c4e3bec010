#!/usr/bin/env python
"""How often could the final ladder rung skip the per-class lerps?  If the four r1 pixels around a 2x2 output block share one
arg-max class, all four outputs have it (bilinear weights are a convex combination).  Prints the fraction of 2x2 output
blocks and of warp-sized regions (4 x 8 r1 pixels = 8 x 16 outputs) whose labels are uniform, for the bench workload
(random-init weights) and for the trained weights of tests/trained.py."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lednet_b200 as L
from lednet_b200 import synth


def frac_uniform(pred, bh, bw):
    n, h, w = pred.shape
    p = pred[:, :h // bh * bh, :w // bw * bw].reshape(n, h // bh, bh, w // bw, bw)
    mx, mn = p.amax(dim=(2, 4)), p.amin(dim=(2, 4))
    return float((mx == mn).float().mean())


def main():
    K = 19
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64, num_classes=K, dropout_ratio=0.),
                             data_preprocessor=dict(type='SegDataPreProcessor', mean=list(L.engine.MEAN), std=list(L.engine.STD),
                                                    bgr_to_rgb=True), compute_dtype='bf16').eval()
    m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=2))
    img = synth.make_images_u8(4, 1024, 2048, seed=100).cuda()
    pred = m.predict_labels(img)
    for (bh, bw) in [(2, 2), (4, 4), (8, 16)]:
        print(f'random-init weights, synthetic images: {bh}x{bw} output blocks uniform: {frac_uniform(pred, bh, bw):.4f}')
    print('label histogram', torch.bincount(pred.flatten().long(), minlength=K).tolist())


if __name__ == '__main__':
    main()
