import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import oracle, lednet_b200 as L
from lednet_b200 import synth
from test_gpu_train import _train_pair
from util import rel_err
K, hw, N = 19, (128, 256), 2
for dbl in (False, True):
    o, m = _train_pair(K)
    x = oracle.preprocess(synth.make_images_u8(N, *hw, seed=0))
    lab = synth.make_labels(N, *hw, K, seed=1)
    if dbl:
        o = o.double(); x = x.double()
    ref = o.loss(x, lab)
    (ref['loss_context'] + ref['loss_spatial']).backward()
    samples = [dict(gt_sem_seg=dict(data=lab[i:i + 1].cuda())) for i in range(N)]
    losses = m.loss(x.float().cuda(), samples)
    total, log = m.parse_losses(losses)
    total.backward()
    print('double oracle' if dbl else 'float oracle', float(total), float(ref['loss_context'] + ref['loss_spatial']))
    rows = []
    po = dict(o.named_parameters())
    for k, p in m.named_parameters():
        g, r = p.grad.cpu().double(), po[k].grad.double()
        rows.append((rel_err(g, r), float((g - r).norm() / r.norm().clamp_min(1e-30)), float(r.abs().max()), k))
    rows.sort(reverse=True)
    for e, e2, mx, k in rows[:25]:
        print(f'{e:9.2e} l2 {e2:9.2e} max|g| {mx:9.2e}  {k}')
    print('n > 1e-2:', sum(r[0] > 1e-2 for r in rows), ' n > 1e-3:', sum(r[0] > 1e-3 for r in rows), 'of', len(rows))
if True:
    # float oracle vs double oracle: how chaotic is the reference itself?
    o, _ = _train_pair(K)
    o2, _ = _train_pair(K)
    o2 = o2.double()
    (lambda r: (r['loss_context'] + r['loss_spatial']).backward())(o.loss(x.float(), lab))
    (lambda r: (r['loss_context'] + r['loss_spatial']).backward())(o2.loss(x.double(), lab))
    p2 = dict(o2.named_parameters())
    rows = sorted(((rel_err(p.grad.double(), p2[k].grad), k) for k, p in o.named_parameters()), reverse=True)
    print('float-oracle vs double-oracle worst:', rows[:8])
