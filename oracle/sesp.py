"""Oracle SESP block (TEST INFRASTRUCTURE) - restates ``mmseg/models/nn_layers/eesp.py:15-118``
with the helper layers of ``mmseg/models/nn_layers/espnet_utils.py:8-168`` (CBR, BR, CB, CDilated).

REDUCE (grouped 1x1 + BN + PReLU) -> SPLIT/TRANSFORM (k depthwise dilated 3x3 with hierarchical
feature fusion, then the SESPV2 cascade: a second depthwise 3x3 with dilation d+1 per branch) ->
MERGE (concat, BN + PReLU, grouped 1x1 + BN) -> residual -> PReLU.  Module and parameter names equal
the reference's so one state dict feeds the verbatim module, this oracle and the CUDA kernel.
Only the stride-1 / non-SPASPP configuration (what SURVEY section 8a row B5 registers) is restated.
"""
import torch
import torch.nn as nn


class _CBR(nn.Module):            # espnet_utils.py:8-37
    def __init__(self, nin, nout, k, stride=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(nin, nout, k, stride, (k - 1) // 2, bias=False, groups=groups)
        self.bn = nn.BatchNorm2d(nout)
        self.act = nn.PReLU(nout)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class _BR(nn.Module):             # espnet_utils.py:40-62
    def __init__(self, nout):
        super().__init__()
        self.bn = nn.BatchNorm2d(nout)
        self.act = nn.PReLU(nout)

    def forward(self, x):
        return self.act(self.bn(x))


class _CB(nn.Module):             # espnet_utils.py:65-92
    def __init__(self, nin, nout, k, stride=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(nin, nout, k, stride, (k - 1) // 2, bias=False, groups=groups)
        self.bn = nn.BatchNorm2d(nout)

    def forward(self, x):
        return self.bn(self.conv(x))


class _CDilated(nn.Module):       # espnet_utils.py:121-145
    def __init__(self, nin, nout, k, stride=1, d=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(nin, nout, k, stride, ((k - 1) // 2) * d, d, groups, bias=False)

    def forward(self, x):
        return self.conv(x)


def sesp_dilations(k=4, r_lim=7, spatial=True):
    """eesp.py:40-57: dilation of each branch's first depthwise conv."""
    if spatial:
        return [1] * k
    table = {3: 1, 5: 2, 7: 3, 9: 4, 11: 5, 13: 6, 15: 7, 17: 6, 19: 12, 21: 18, 23: 24}
    ks = sorted((3 + 2 * i) if (3 + 2 * i) <= r_lim else 3 for i in range(k))
    return [table[s] for s in ks]


class OracleSESP(nn.Module):
    def __init__(self, nIn, nOut, stride=1, k=4, r_lim=7, down_method='esp', Spatial=True,
                 SPASPP_Flag=False, SESPV2=True):
        super().__init__()
        assert stride == 1 and not SPASPP_Flag, 'only the stride-1 block is on the path'
        n = nOut // k
        assert n * k == nOut
        self.SESPV2 = SESPV2
        d = sesp_dilations(k, r_lim, Spatial)
        self.proj_1x1 = _CBR(nIn, n, 1, 1, groups=k)
        self.spp_dw = nn.ModuleList([_CDilated(n, n, 3, 1, di, groups=n) for di in d])
        if SESPV2:
            self.spp_dw_v2 = nn.ModuleList([_CDilated(n, n, 3, 1, di + 1, groups=n) for di in d])
        self.conv_1x1_exp = _CB(nOut, nOut, 1, 1, groups=k)
        self.br_after_cat = _BR(nOut)
        self.module_act = nn.PReLU(nOut)

    def forward(self, x):
        o1 = self.proj_1x1(x)
        outs = [self.spp_dw[0](o1)]
        for i in range(1, len(self.spp_dw)):
            outs.append(self.spp_dw[i](o1) + outs[i - 1])          # HFF, eesp.py:90-95
        if self.SESPV2:
            outs = [self.spp_dw_v2[i](outs[i]) for i in range(len(outs))]
        e = self.conv_1x1_exp(self.br_after_cat(torch.cat(outs, 1)))
        if e.size() == x.size():
            e = e + x
        return self.module_act(e)
