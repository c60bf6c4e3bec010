"""GPU parity of the MFAF gate and the GETB block (row B6) through the registered
modules / C ABI: against the golden outputs of the reference's own modules and against the oracle on larger
seeded inputs (odd sizes, both memory formats, fp32 and bf16)."""
import os

import numpy as np
import pytest
import torch

import lednet_b200 as L
import block_cases as bc
from oracle.mfaf import OracleMutiAFF
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_mfaf_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'mfaf.npz'))
    for i, (tag, kw, shape) in enumerate(bc.MFAF_CASES):
        m = L.MODELS.build(dict(type='Muti_AFF', **kw)).eval()
        m.load_state_dict(bc.block_state_dict(m.state_dict(), seed=31), strict=True)
        assert sum(p.numel() for p in m.parameters()) == int(g[tag + '_nparams'])
        x, r = bc.block_input(i, kw['channels'], shape, 300, n_inputs=2)
        out = m(x.to(DEV), r.to(DEV))
        assert out.shape == g[tag].shape
        assert rel_err(out.cpu(), torch.from_numpy(g[tag])) < 2e-5, tag


@pytest.mark.parametrize('kw,shape', [
    (dict(channels=64), (2, 128, 256)),        # the 1/8-resolution map of a 1024x2048 image
    (dict(channels=64), (1, 45, 83)),
    (dict(channels=128, r=4), (2, 17, 31)),
    (dict(channels=256, r=4), (1, 16, 16)),
    (dict(channels=32, r=4), (1, 3, 5)),
    (dict(channels=64), (1, 1, 1)),
])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_mfaf_vs_oracle(kw, shape, dtype):
    o = OracleMutiAFF(**kw).eval()
    sd = bc.block_state_dict(o.state_dict(), seed=13)
    o.load_state_dict(sd)
    m = L.Muti_AFF(**kw).eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(shape[1] * 7 + shape[2])
    x = torch.randn(shape[0], kw['channels'], shape[1], shape[2], generator=g).to(dtype).float()
    r = torch.randn(shape[0], kw['channels'], shape[1], shape[2], generator=g).to(dtype).float()
    with torch.no_grad():
        ref = o(x, r)
    for channels_last in (False, True):
        xd, rd = x.to(DEV, dtype), r.to(DEV, dtype)
        if channels_last:
            xd, rd = (t.contiguous(memory_format=torch.channels_last) for t in (xd, rd))
        out = m(xd, rd)
        assert out.dtype == dtype and out.shape == ref.shape
        # bf16: inputs pre-rounded, so only the output rounding (2^-9) remains
        tol = 2e-5 if dtype == torch.float32 else 6e-3
        assert rel_err(out.float().cpu(), ref) < tol, (channels_last, rel_err(out.float().cpu(), ref))


def test_mfaf_errors():
    with pytest.raises(NotImplementedError):
        L.Muti_AFF(channels=60)
    m = L.Muti_AFF(64).eval()
    with pytest.raises(L.LedB200Error):
        m(torch.zeros(1, 64, 4, 4), torch.zeros(1, 64, 4, 4))
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 64, 4, 4, device=DEV), torch.zeros(1, 64, 4, 4, device=DEV))
