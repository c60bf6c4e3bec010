"""Ramp timeline of every conv_tc launch of the batch-16 1024x2048 plan (probe, not a test or a bench).

CTA 0 of each launch stamps clock64 at the points of its ramp (csrc/conv_tc.cu, -DLEDB_TC_TIMELINE build,
tools/probes/build_tc_timeline.sh).  Printed in microseconds at the SM clock measured from the two globaltimer
stamps; `gap` = idle time between the previous launch's exit and this launch's entry on the graph replay.

    gpurun -- 'python tools/probes/tc_timeline.py > gpurun_out/tc_timeline.txt'
"""
import ctypes as C
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import lednet_b200 as L  # noqa: E402
from lednet_b200 import synth  # noqa: E402
from lednet_b200 import lib as LIB  # noqa: E402

LIB.LIB_PATH = os.path.join(os.path.dirname(LIB.LIB_PATH), 'libledb200_tl.so')
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
K = 19
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64, num_classes=K,
                                                   dropout_ratio=0.),
                         data_preprocessor=dict(type='SegDataPreProcessor', mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], bgr_to_rgb=True)).eval()
m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=2))
img = synth.make_images_u8(batch, 1024, 2048, seed=0).cuda()
x = ((img[:, [2, 1, 0]].float() - torch.tensor(L.engine.MEAN, device='cuda').view(1, 3, 1, 1))
     / torch.tensor(L.engine.STD, device='cuda').view(1, 3, 1, 1)).contiguous()
for _ in range(4):          # call 1 captures the graph (launch ordinals 0..), later calls replay it into the same slots
    m.predict_labels(x)
torch.cuda.synchronize()
eng = m.engine()
names = [n for (n, kind, fl, by) in eng.op_info() if kind == 'conv_tc' and n != 'backbone.stem.0']   # stem.0 runs stem_tc_kernel
lib = LIB.get()
buf = (C.c_ulonglong * (128 * 16))()
assert lib.ledb200_probe_tc_timeline(buf) == 0
rows = [[buf[i * 16 + j] for j in range(16)] for i in range(128)]
print('# stamps (us after kernel entry of CTA 0): init = barriers initialised + weight TMA issued; sync = first CTA sync '
      '(TMEM allocated, bias staged); a_tma = first A slab requested; b_ok = resident weights landed; a_ok = first A slab '
      'landed; mma = first tile committed; acc = accumulator visible to the epilogue; epi = first tile stored; '
      'prod_end / epi_end = roles finished; end = after the final sync + TMEM dealloc; wall = globaltimer exit - entry; '
      'gap = entry - the previous launch\'s exit (same graph replay)')
hdr = ['init', 'sync', 'a_tma', 'b_ok', 'a_ok', 'mma', 'acc', 'epi', 'prod_end', 'epi_end', 'end', 'wall', 'gap']
print('%-46s ' % 'op' + ' '.join('%8s' % h for h in hdr))
# an eager first pass (if any) takes ordinals 0..n-1 and the captured graph the next n: show the block that ran last
off = max((o for o in (0, len(names)) if o + len(names) <= 128), key=lambda o: rows[o][12])
rows = rows[off:]
print('# launch ordinals %d..%d' % (off, off + len(names) - 1))
prev_exit = None
tot_wall = tot_gap = 0.0
for i, name in enumerate(names):
    r = rows[i]
    wall_ns = r[13] - r[12]
    clk = r[11] - r[0]
    mhz = clk / max(wall_ns, 1) * 1e3
    us = lambda j: (r[j] - r[0]) / max(mhz, 1.0) if r[j] else float('nan')   # noqa: E731
    gap = (r[12] - prev_exit) / 1e3 if prev_exit else float('nan')
    prev_exit = r[13]
    vals = [us(1), us(2), us(3), us(4), us(5), us(6), us(7), us(8), us(9), us(14), us(11), wall_ns / 1e3, gap]
    tot_wall += wall_ns / 1e3
    if gap == gap and abs(gap) < 1e3:
        tot_gap += gap
    print('%-46s ' % name[-46:] + ' '.join('%8.2f' % v for v in vals) + '  (%.0f MHz)' % mhz)
print('# sum of wall %.1f us, sum of gaps (|gap| < 1 ms) %.1f us over %d launches' % (tot_wall, tot_gap, len(names)))
