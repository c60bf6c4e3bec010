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


def test_mfaf_and_getb_equal_verbatim_reference():
    """fresh seeded inputs (not the committed goldens), including sizes that are not multiples of the pooling
    grids / attention window."""
    from oracle.mfaf import OracleMutiAFF
    from oracle.getb import OracleGETBBlock
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(77)
    for kw, hw in ((dict(channels=64), (23, 41)), (dict(channels=128, r=4), (16, 32))):
        a, b = ref.Muti_AFF(**kw).eval(), OracleMutiAFF(**kw).eval()
        sd = synth.make_state_dict(a.state_dict(), seed=51)
        a.load_state_dict(sd), b.load_state_dict(sd)
        x, r = (torch.randn(2, kw['channels'], *hw, generator=g) for _ in range(2))
        with torch.no_grad():
            assert torch.equal(a(x, r), b(x, r))
    for kw, hw in ((dict(dim=128, num_heads=8, window_size=8), (11, 19)), (dict(dim=64, num_heads=8, window_size=4), (8, 12))):
        a, b = ref.GETBBlock(**kw).eval(), OracleGETBBlock(**kw).eval()
        sd = synth.make_state_dict(a.state_dict(), seed=52)
        a.load_state_dict(sd), b.load_state_dict(sd)
        x = torch.randn(1, kw['dim'], *hw, generator=g)
        with torch.no_grad():
            ya, yb = a(x), b(x)
        assert float((ya - yb).abs().max()) <= 1e-6 * float(ya.abs().max())


def test_seam_equals_prototype_statements():
    """oracle/seam.py == the SEAM statements of tools/speed/ddrnet_speed.py executed verbatim (AST slice)."""
    import block_cases as bc
    from oracle.seam import OracleSEAM
    r = ref_loader.load_seam()().eval()
    o = OracleSEAM(64).eval()
    sd = bc.seam_state_dict(o.state_dict())
    r.load_state_dict(sd), o.load_state_dict(sd)
    assert r.boundary_threshold == o.boundary_threshold
    assert torch.equal(r.laplacian_kernel, o.laplacian_kernel)
    for tag, shape in bc.SEAM_CASES:
        x, xs = bc.seam_inputs(shape)
        with torch.no_grad():
            assert torch.equal(r(x, xs), o(x, xs)), tag          # same process, same ATen kernels: bit-equal
            mr, er = r.edge(x)
            mo, eo = o.edge_mask(x)
            assert torch.equal(mr, mo) and torch.equal(er, eo)
