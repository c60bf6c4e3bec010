"""Where does tf32 cost gradient accuracy?  Product gradients (fp32 kernels, tf32 tensor-core kernels, and tf32 with one
op class at a time) against the float64 oracle on the train-step test problem."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import oracle
import lednet_b200 as L
from lednet_b200 import synth, train_ops as T
from test_gpu_train import _train_pair

K, hw, N = (int(sys.argv[1]), (int(sys.argv[2]), int(sys.argv[3])), int(sys.argv[4])) if len(sys.argv) > 4 else (19, (128, 256), 2)
if os.environ.get('DIAG_SCENE'):      # structured images (class-coloured blocky regions + noise) instead of white noise
    img, lab = synth.make_scene(N, *hw, K, seed=0, coarse=(hw[0] // 32, hw[1] // 32), ignore_frac=0.05)
    x = oracle.preprocess(img)
else:
    x = oracle.preprocess(synth.make_images_u8(N, *hw, seed=0))
    lab = synth.make_labels(N, *hw, K, seed=1)
o64, _ = _train_pair(K)
o64 = o64.double()
r = o64.loss(x.double(), lab)
(r['loss_context'] + r['loss_spatial']).backward()
g64 = {k: p.grad.detach().clone() for k, p in o64.named_parameters()}
o32, _ = _train_pair(K)
r = o32.loss(x, lab)
(r['loss_context'] + r['loss_spatial']).backward()
g32 = {k: p.grad.detach().clone() for k, p in o32.named_parameters()}


def report(tag, grads, top=8):
    num = sum(float((grads[k].double() - g).pow(2).sum()) for k, g in g64.items())
    den = sum(float(g.pow(2).sum()) for g in g64.values())
    rows = sorted(((float((grads[k].double() - g).abs().max()) / max(float(g.abs().max()), 1e-30), k) for k, g in g64.items()),
                  reverse=True)
    print(f'{tag:34s} whole-vector {(num / den) ** 0.5:9.2e}   n>1e-2: {sum(e > 1e-2 for e, _ in rows):3d}   worst: '
          + ', '.join(f'{k.replace("backbone.", "b.").replace("decode_head.", "h.")} {e:.1e}' for e, k in rows[:top]))


report('fp32 oracle (CPU)', g32)
samples = [dict(gt_sem_seg=dict(data=lab[i:i + 1].cuda())) for i in range(N)]
for tag, tc, fast, ops in (('product fp32 CUDA cores', False, False, 7), ('product tensor cores, 3 passes', True, False, 7),
                           ('  3 passes, forward only', True, False, 1), ('  3 passes, dgrad only', True, False, 2),
                           ('  3 passes, wgrad only', True, False, 4), ('product tensor cores, 1 pass (fast)', True, True, 7)):
    T.set_tensor_cores(tc, fast=fast)
    T.TC_OPS = ops
    _, m = _train_pair(K)
    total, _ = m.parse_losses(m.loss(x.cuda(), samples))
    total.backward()
    torch.cuda.synchronize()
    report(tag, {k: p.grad.cpu() for k, p in m.named_parameters()})
