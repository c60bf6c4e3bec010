"""Margin-realistic weights for the bf16 argmax gate (TEST INFRASTRUCTURE).

north_star gates bf16 inference at >= 99.9 % argmax agreement with the fp32 reference.  That figure
presumes a TRAINED network: random-init weights leave the top-2 logits inside bf16 rounding noise on
~0.5 % of pixels (tests/util.py: bf16_storage_agreement).  This module trains the CPU oracle (the
reference's own modules restated, torch autograd, fp32) for a few hundred SGD steps of the reference's
loss (LEDHead.loss: two OHEM cross-entropies, led_head.py:101-146; SGD lr 0.01 momentum 0.9 wd 5e-4,
configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:64-65) on the learnable blocky synthetic scenes of
`synth.make_scene`, seeded and single-threaded-deterministic, and caches the resulting state dict.
Both the oracle and the CUDA engine then load the SAME trained state dict.
"""
import os
import time

import torch

import oracle
from lednet_b200 import synth

CACHE_DIR = os.environ.get('LEDB200_TEST_CACHE', '/tmp/ledb200_test_cache')


SIZES = [((128, 256), (4, 8)), ((192, 384), (6, 12)), ((256, 512), (4, 8)), ((192, 384), (3, 6)), ((256, 256), (8, 8))]


def train_oracle(num_classes, steps=300, batch=2, seed=11, lr=0.01, verbose=False):
    torch.manual_seed(seed)
    o = oracle.OracleSegmentor(num_classes=num_classes)
    o.load_state_dict(synth.make_state_dict(o.state_dict(), seed=2))
    # OHEM as configured by the reference (thres 0.9, loss weights 1.0 / 0.4); min_kept scaled to the crop
    o.train()
    opt = torch.optim.SGD(o.parameters(), lr=lr, momentum=0.9, weight_decay=5e-4)
    t0 = time.time()
    for it in range(steps):
        hw, coarse = SIZES[it % len(SIZES)]               # mixed crop sizes / region sizes: the pooled context must generalise
        # OHEM as configured by the reference (thres 0.9, loss weights 1.0 / 0.4); min_kept scaled to the crop
        o.decode_head.loss_decode = [dict(thres=0.9, min_kept=batch * hw[0] * hw[1] // 8, loss_weight=1.0),
                                     dict(thres=0.9, min_kept=batch * hw[0] * hw[1] // 8, loss_weight=0.4)]
        img, lab = synth.make_scene(batch, hw[0], hw[1], num_classes, seed=1000 * seed + it, coarse=coarse)
        losses = o.loss(oracle.preprocess(img), lab)
        loss = losses['loss_context'] + losses['loss_spatial']
        opt.zero_grad(set_to_none=True)
        loss.backward()
        for g in opt.param_groups:                        # PolyLR power 0.9 (config :67-75)
            g['lr'] = lr * (1 - it / steps) ** 0.9
        opt.step()
        if verbose and (it % 20 == 0 or it == steps - 1):
            print(f'it {it:4d} loss {loss.item():.4f} acc {losses["acc_seg"].item():.2f} t {time.time() - t0:.0f}s', flush=True)
    return {k: v.detach().clone() for k, v in o.eval().state_dict().items()}


def trained_state_dict(num_classes, **kw):
    os.makedirs(CACHE_DIR, exist_ok=True)
    tag = '_'.join(f'{k}{v}' for k, v in sorted(kw.items())).replace(' ', '').replace('(', '').replace(')', '').replace(',', 'x')
    path = os.path.join(CACHE_DIR, f'trained_k{num_classes}_{tag or "default"}.pt')
    if os.path.isfile(path):
        return torch.load(path)
    sd = train_oracle(num_classes, **kw)
    torch.save(sd, path + '.tmp')
    os.replace(path + '.tmp', path)
    return sd


# ------------------------------------------------------------------------------------------------------------------
# GPU variant: the same recipe through the repo's own training step (train-mode LEDNet / LEDHead, OhemCrossEntropy,
# FlatSGD + PolyLR on csrc/train.cu kernels), which affords full-size crops (a 1024x2048 step is ~50 ms on a B200 where
# the CPU oracle needs ~10 s).  Used by the -m gpu full-size parity tests; the weights it returns are loaded into BOTH
# the CPU oracle and the CUDA engine, so the parity statement does not depend on how they were produced.
_mem_cache = {}


def train_product(num_classes, schedule, steps=300, seed=11, lr=0.01, verbose=False):
    """schedule: list of (batch, (h, w), (coarse_h, coarse_w)) cycled over the steps."""
    import warnings
    import lednet_b200 as L
    key = (num_classes, tuple(schedule), steps, seed, lr)
    if key in _mem_cache:
        return _mem_cache[key]
    dev = torch.device('cuda')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(dict(type='LEDNet'),
                             dict(type='LEDHead', in_channels=128, channels=64, num_classes=num_classes, dropout_ratio=0.),
                             data_preprocessor=None, compute_dtype='fp32')
    m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=2))
    m.to(dev).train()
    opt = L.FlatSGD(m.parameters(), lr=lr, momentum=0.9, weight_decay=5e-4)
    sched = L.PolyLR(opt, power=0.9, eta_min=0.0, end=steps)
    t0 = time.time()
    for it in range(steps):
        batch, hw, coarse = schedule[it % len(schedule)]
        img, lab = synth.make_scene(batch, hw[0], hw[1], num_classes, seed=1000 * seed + it, coarse=coarse)
        x = oracle.preprocess(img).to(dev)
        lab = lab.to(dev)
        samples = [dict(gt_sem_seg=dict(data=lab[i:i + 1])) for i in range(batch)]
        total, log = m.parse_losses(m.loss(x, samples))
        opt.zero_grad()
        total.backward()
        opt.step()
        sched.step()
        if verbose and (it % 25 == 0 or it == steps - 1):
            print(f'it {it:4d} loss {float(total):.4f} acc {float(log["decode.acc_seg"]) if "decode.acc_seg" in log else -1:.2f} '
                  f't {time.time() - t0:.0f}s', flush=True)
    sd = {k: v.detach().float().cpu().clone() for k, v in m.eval().state_dict().items()}
    _mem_cache[key] = sd
    return sd
