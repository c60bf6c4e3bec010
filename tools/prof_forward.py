"""Profiling driver (not a test): one whole-path forward at a chosen batch, for use under ncu.

    ncu --set full -k regex:conv_tc -c 8 -o gpurun_out/x python tools/prof_forward.py --batch 4
"""
import argparse
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import lednet_b200 as L  # noqa: E402
from lednet_b200 import synth, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=4)
ap.add_argument('--height', type=int, default=1024)
ap.add_argument('--width', type=int, default=2048)
ap.add_argument('--classes', type=int, default=19)
ap.add_argument('--iters', type=int, default=1)
ap.add_argument('--u8', action='store_true')
args = ap.parse_args()
K = args.classes
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64, num_classes=K,
                                                   dropout_ratio=0.),
                         data_preprocessor=dict(type='SegDataPreProcessor', mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], bgr_to_rgb=True)).eval()
m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=2))
img = synth.make_images_u8(args.batch, args.height, args.width, seed=0).cuda()
x = img if args.u8 else ((img[:, [2, 1, 0]].float() - torch.tensor(L.engine.MEAN, device='cuda').view(1, 3, 1, 1))
                         / torch.tensor(L.engine.STD, device='cuda').view(1, 3, 1, 1)).contiguous()
lab = synth.make_labels(args.batch, args.height, args.width, K, seed=1).to(torch.uint8).cuda()
for _ in range(args.iters):
    pred = m.predict_labels(x)
    cm = ops.confusion_accumulate(pred, lab, K)
torch.cuda.synchronize()
print('ok', int(cm.sum()))
