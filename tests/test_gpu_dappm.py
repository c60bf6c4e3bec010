"""The fused DAPPM (csrc/dappm.cu: pooled branches + one clustered tcgen05 kernel) against the oracle's DAPPM
(ppm.py:57-130 restated, pinned by the golden fixtures) and against the 22-launch path it replaces."""
import os

import pytest
import torch

import oracle
from lednet_b200 import synth
from util import build_pair, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _oracle_taps(o, x):
    feats = {}
    hooks = [o.backbone.spp.register_forward_hook(lambda mod, i, out: feats.__setitem__('spp.out', out.detach().clone())),
             o.backbone.spp.register_forward_pre_hook(lambda mod, i: feats.__setitem__('spp.in', i[0].detach().clone()))]
    with torch.no_grad():
        logits, pred = o.predict(x)
    for h in hooks:
        h.remove()
    return feats, logits, pred


@pytest.mark.parametrize('n,hw', [
    (2, (128, 256)),        # 2 x 4 DAPPM pixels: one mostly empty tile, pooled maps of 1 x 2 .. 1 x 1
    (1, (512, 1024)),       # 8 x 16: two tiles side by side, half a tile high
    (2, (1024, 1024)),      # 16 x 16: the training crop
    (1, (1024, 2048)),      # 16 x 32: the benchmarked shape, cluster of 4
    (1, (544, 800)),        # 9 x 13: ragged tiles in both directions
    (3, (2048, 1024)),      # 32 x 16: two tile rows
])
def test_fused_dappm_vs_oracle(n, hw):
    o, m = build_pair(19, dtype='bf16')
    x = oracle.preprocess(synth.make_images_u8(n, *hw, seed=21))
    feats, ref_logits, _ = _oracle_taps(o, x)
    eng = m.engine()
    os.environ.pop('LEDB200_NO_FUSED_DAPPM', None)
    pred, logits = eng.forward_infer(x.to(DEV), want_logits=True)
    names = [eng.lib.ledb200_op_name(eng.h, i).decode() for i in range(eng.plan_launches())]
    assert any('fused chain' in s for s in names), 'the fused DAPPM did not run'
    got = eng.debug_fetch('spp.out')
    ref = feats['spp.out']
    assert tuple(got.shape) == tuple(ref.shape)
    err = rel_err(got, ref)
    assert err < 2e-2, err                                        # bf16 activation tolerance of north_star
    assert rel_err(logits.cpu(), ref_logits) < 2e-2
    # the path it replaces, same engine weights: both are bf16 evaluations of the same module
    os.environ['LEDB200_NO_FUSED_DAPPM'] = '1'
    try:
        _, m2 = build_pair(19, dtype='bf16')
        eng2 = m2.engine()
        eng2.forward_infer(x.to(DEV))
        names2 = [eng2.lib.ledb200_op_name(eng2.h, i).decode() for i in range(eng2.plan_launches())]
        assert not any('fused chain' in s for s in names2)
        old = eng2.debug_fetch('spp.out')
    finally:
        os.environ.pop('LEDB200_NO_FUSED_DAPPM', None)
    err_old = rel_err(old, ref)
    print(f'DAPPM {n}x{hw}: fused vs oracle {err:.2e}, 22-launch path vs oracle {err_old:.2e}, fused vs old {rel_err(got, old):.2e}; '
          f'{len(names)} launches (was {len(names2)})')
    assert err <= max(1.5 * err_old, 5e-3)                       # no worse than the path it replaces


def test_fused_dappm_is_deterministic_and_graph_safe():
    o, m = build_pair(2, dtype='bf16')
    x = oracle.preprocess(synth.make_images_u8(4, 512, 512, seed=3)).to(DEV)
    eng = m.engine()
    a = eng.forward_infer(x).clone()
    outs = [eng.debug_fetch('spp.out')]
    for _ in range(3):                                            # graph replays
        b = eng.forward_infer(x)
        assert torch.equal(a, b)
        outs.append(eng.debug_fetch('spp.out'))
    for t in outs[1:]:
        assert torch.equal(outs[0], t)
