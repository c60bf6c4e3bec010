"""GPU parity of the fused SESP kernel through the registered `SESP`
module / C ABI: against the golden outputs of the reference's own eesp.py, and against the oracle on
larger seeded inputs (odd sizes, tiles that straddle the border, both memory formats, fp32 and bf16)."""
import os

import numpy as np
import pytest
import torch

import lednet_b200 as L
from lednet_b200 import synth
from oracle.sesp import OracleSESP
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _helpers(golden_dir):
    src = open(os.path.join(golden_dir, 'make_golden.py')).read()
    ns = {}
    exec('import torch\nfrom lednet_b200 import synth\n' + src[src.index('SESP_CASES = ['):src.index('def make_sesp')], ns)
    return ns


def test_sesp_vs_reference_golden(golden_dir):
    ns = _helpers(golden_dir)
    g = np.load(os.path.join(golden_dir, 'sesp.npz'))
    for i, (tag, kw, shape) in enumerate(ns['SESP_CASES']):
        m = L.MODELS.build(dict(type='SESP', **kw)).eval()
        m.load_state_dict(ns['sesp_state_dict'](m.state_dict()), strict=True)
        out = m(ns['sesp_input'](i, kw['nIn'], shape).to(DEV))
        assert out.shape == g[tag].shape
        assert rel_err(out.cpu(), torch.from_numpy(g[tag])) < 1e-5, tag


@pytest.mark.parametrize('kw,shape', [
    (dict(nIn=64, nOut=64, Spatial=True), (2, 37, 50)),
    (dict(nIn=128, nOut=128, Spatial=False, r_lim=9), (2, 19, 33)),
    (dict(nIn=256, nOut=256, Spatial=False, r_lim=9), (1, 16, 16)),
    (dict(nIn=64, nOut=128, Spatial=False, r_lim=7), (1, 5, 3)),
    (dict(nIn=64, nOut=64, Spatial=True, SESPV2=False), (1, 1, 1)),
])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_sesp_vs_oracle(golden_dir, kw, shape, dtype):
    ns = _helpers(golden_dir)
    o = OracleSESP(**kw).eval()
    sd = ns['sesp_state_dict'](o.state_dict(), seed=11)
    o.load_state_dict(sd)
    m = L.SESP(**kw).eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(shape[1] * 7 + shape[2])
    x = torch.randn(shape[0], kw['nIn'], shape[1], shape[2], generator=g).to(dtype).float()
    with torch.no_grad():
        ref = o(x)
    for channels_last in (False, True):
        xd = x.to(DEV, dtype)
        if channels_last:
            xd = xd.contiguous(memory_format=torch.channels_last)
        out = m(xd)
        assert out.dtype == dtype and out.shape == ref.shape
        # bf16: input pre-rounded, so only the output rounding (2^-9) remains
        tol = 2e-5 if dtype == torch.float32 else 6e-3
        assert rel_err(out.float().cpu(), ref) < tol, (channels_last, rel_err(out.float().cpu(), ref))


def test_sesp_errors():
    with pytest.raises(NotImplementedError):
        L.SESP(64, 64, stride=2)
    m = L.SESP(64, 64).eval()
    with pytest.raises(L.LedB200Error):
        m(torch.zeros(1, 64, 4, 4))
