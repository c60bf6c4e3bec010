"""Bottleneck probe for csrc/conv_tc.cu (not a test, not a bench): times every conv_tc op of the
batch-16 1024x2048 plan with the kernel's roles switched off one at a time (LEDB200_TC_DBG bits:
1 no global stores, 2 no MMAs, 4 no A-operand TMA, 8 no residual loads).  Results are WRONG by
construction when a bit is set; only the times mean anything.

    gpurun -- 'python tools/probe_tc.py > gpurun_out/probe_tc.txt'
"""
import argparse
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import lednet_b200 as L  # noqa: E402
from lednet_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--height', type=int, default=1024)
ap.add_argument('--width', type=int, default=2048)
ap.add_argument('--classes', type=int, default=19)
ap.add_argument('--modes', default='0,1,2,4,8,3,5,6,7,15')
args = ap.parse_args()
K = args.classes
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64, num_classes=K,
                                                   dropout_ratio=0.),
                         data_preprocessor=dict(type='SegDataPreProcessor', mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], bgr_to_rgb=True)).eval()
m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=2))
img = synth.make_images_u8(args.batch, args.height, args.width, seed=0).cuda()
x = ((img[:, [2, 1, 0]].float() - torch.tensor(L.engine.MEAN, device='cuda').view(1, 3, 1, 1))
     / torch.tensor(L.engine.STD, device='cuda').view(1, 3, 1, 1)).contiguous()
m.predict_labels(x)
torch.cuda.synchronize()
eng = m.engine()
info = eng.op_info()
modes = [int(v) for v in args.modes.split(',')]
cols = {}
for md in modes:
    os.environ['LEDB200_TC_DBG'] = str(md)
    cols[md] = eng.profile_ops(iters=3)
os.environ['LEDB200_TC_DBG'] = '0'
bw = 6550.7e9
print('%-48s %8s %8s ' % ('op', 'MB', 'roof_us') + ' '.join('dbg%-5d' % md for md in modes))
tot = {md: 0.0 for md in modes}
for i, (name, kind, fl, by) in enumerate(info):
    if kind != 'conv_tc':
        continue
    roof = max(by / bw, fl / 1407.5e12) * 1e6
    row = '%-48s %8.1f %8.1f ' % (name[-48:], by / 1e6, roof)
    for md in modes:
        t = cols[md][i][1] * 1e3
        tot[md] += t
        row += '%8.1f ' % t
    print(row)
print('%-48s %8s %8s ' % ('TOTAL conv_tc (us)', '', '') + ' '.join('%8.1f' % tot[md] for md in modes))
