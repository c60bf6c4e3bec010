"""Oracle MFAF gate (TEST INFRASTRUCTURE) - restates ``Muti_AFF`` of
``mmseg/models/classification/model_utils.py:356-429``.

    xa  = x + residual
    att = local(xa) + global(pool1(xa)) + sum_s nearest_up(ctx_s(pool_s(xa)))   s in {4, 8, 16}
    out = 2 x sigmoid(att) + 2 residual (1 - sigmoid(att))

Every attention path is conv1x1(C -> C/r, bias) -> BN -> ReLU -> conv1x1(C/r -> C, bias) -> BN.
Module and parameter names equal the reference's (``local_att.0.weight`` ...), so one state dict feeds
the verbatim module, this oracle and the CUDA kernels.
"""
import torch.nn as nn
import torch.nn.functional as F


def _att(channels, inter, pool):
    layers = [] if pool is None else [nn.AdaptiveAvgPool2d(pool)]
    layers += [nn.Conv2d(channels, inter, 1), nn.BatchNorm2d(inter), nn.ReLU(inplace=True),
               nn.Conv2d(inter, channels, 1), nn.BatchNorm2d(channels)]
    return nn.Sequential(*layers)


class OracleMutiAFF(nn.Module):
    def __init__(self, channels=64, r=4):
        super().__init__()
        inter = int(channels // r)                       # model_utils.py:362
        self.local_att = _att(channels, inter, None)     # :364-370
        self.context1 = _att(channels, inter, (4, 4))    # :372-379
        self.context2 = _att(channels, inter, (8, 8))    # :381-388
        self.context3 = _att(channels, inter, (16, 16))  # :390-397
        self.global_att = _att(channels, inter, 1)       # :399-406

    def forward(self, x, residual):                      # :410-429
        h, w = x.shape[2], x.shape[3]
        xa = x + residual
        att = self.local_att(xa) + self.global_att(xa)
        for ctx in (self.context1, self.context2, self.context3):
            att = att + F.interpolate(ctx(xa), size=[h, w], mode='nearest')
        wei = att.sigmoid()
        return 2 * x * wei + 2 * residual * (1 - wei)
