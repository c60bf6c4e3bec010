"""When /root/reference is mounted (build container only): the restated oracle must equal the
reference's own modules executed verbatim, on fresh seeded inputs (not just the committed goldens)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref_loader
import lednet_b200  # noqa: F401
from lednet_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference tree not mounted')


def test_trunk_head_fusion_equal_verbatim_reference():
    ref = ref_loader.load()
    torch.manual_seed(1)
    ddr = ref.DDRNet(in_channels=3, channels=32, ppm_channels=128)
    bb = oracle.OracleLEDNet(3, 32, 128)
    sd = synth.make_state_dict(ddr.state_dict(), seed=21)
    ddr.load_state_dict(sd), bb.load_state_dict(sd)
    ddr.eval(), bb.eval()
    x = oracle.preprocess(synth.make_images_u8(1, 72, 104, seed=3))   # odd /8 sizes: ceil paths
    with torch.no_grad():
        c5_ref = ddr(x)
        c5, x1, x2 = bb(x)
    assert torch.equal(c5, c5_ref)
    assert x1.shape == (1, 32, 36, 52) and x2.shape == (1, 32, 18, 26) and c5.shape == (1, 128, 9, 13)


def test_ohem_equals_reference_module():
    ref = ref_loader.load()
    g = np.random.default_rng(5)
    score = torch.from_numpy(g.normal(scale=3, size=(2, 19, 33, 47)).astype(np.float32))
    target = synth.make_labels(2, 33, 47, 19, seed=8, block=8)
    for kw in (dict(thres=0.9, min_kept=100), dict(thres=0.2, min_kept=2500),
               dict(thres=0.7, min_kept=100000)):
        a = ref.OhemCrossEntropy(**kw)(score, target)
        b = oracle.ohem_cross_entropy(score, target, **kw)
        assert torch.equal(a, b)
    assert torch.equal(ref.accuracy(score, target, ignore_index=255),
                       oracle.accuracy(score, target, ignore_index=255))


def test_sesp_block_loads():
    ref = ref_loader.load()
    if ref.SESP is None:
        pytest.skip('SESP not loadable')
    m = ref.SESP(64, 64)
    assert sum(p.numel() for p in m.parameters()) == 2864
