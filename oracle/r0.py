"""Oracle trunk R0 for the registered name ``LEDNet`` (TEST INFRASTRUCTURE).

The LED-Net backbone source is withheld (``mmseg/models/backbones/lednet.py:1-8``
is a notice, not Python), so the trunk is R0 = the DDRNet-23-slim body, every op
of which is pinned by shipped source, plus two stem taps:

* trunk wiring .......... ``mmseg/models/backbones/ddrnet.py:35-224``
* BasicBlock/Bottleneck .. ``mmseg/models/utils/basic_block.py:13-75,156-221``
* DAPPM ................. ``mmseg/models/utils/ppm.py:12-130``
* resize ................ ``mmseg/models/utils/wrappers.py:8-27``
* output contract ....... ``mmseg/models/decode_heads/led_head.py:66-81``
  eval ``(c5[N,4C,H/8,W/8], x1[N,C,H/2,W/2], x2[N,C,H/4,W/4])``,
  train ``(c3[N,2C,H/8,W/8], c5, x1, x2)``

Module/attribute names equal the reference's so one state dict feeds the
verbatim reference modules, this oracle and the CUDA engine.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .mmcv_shim import ConvModule, build_norm_layer

BN = dict(type='BN', requires_grad=True)
RELU = dict(type='ReLU', inplace=True)


def resize(x, size, align_corners=False):
    """wrappers.py:8-27 reduced to the one mode the path uses."""
    return F.interpolate(x, size=tuple(int(s) for s in size), mode='bilinear',
                         align_corners=align_corners)


class BasicBlock(nn.Module):
    """basic_block.py:13-75: 3x3-BN-ReLU, 3x3-BN, += residual, optional act."""
    expansion = 1

    def __init__(self, in_channels, channels, stride=1, downsample=None,
                 norm_cfg=dict(type='BN'), act_cfg=RELU, act_cfg_out=RELU):
        super().__init__()
        self.conv1 = ConvModule(in_channels, channels, 3, stride=stride,
                                padding=1, norm_cfg=norm_cfg, act_cfg=act_cfg)
        self.conv2 = ConvModule(channels, channels, 3, padding=1,
                                norm_cfg=norm_cfg, act_cfg=None)
        self.downsample = downsample
        if act_cfg_out:
            self.act = nn.ReLU(inplace=True)

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        y = y + (self.downsample(x) if self.downsample is not None else x)
        return self.act(y) if hasattr(self, 'act') else y


class Bottleneck(nn.Module):
    """basic_block.py:156-221: 1x1, 3x3(stride), 1x1 (x2 channels), += res."""
    expansion = 2

    def __init__(self, in_channels, channels, stride=1, downsample=None,
                 norm_cfg=dict(type='BN'), act_cfg=RELU, act_cfg_out=None):
        super().__init__()
        self.conv1 = ConvModule(in_channels, channels, 1, norm_cfg=norm_cfg,
                                act_cfg=act_cfg)
        self.conv2 = ConvModule(channels, channels, 3, stride, 1,
                                norm_cfg=norm_cfg, act_cfg=act_cfg)
        self.conv3 = ConvModule(channels, channels * 2, 1, norm_cfg=norm_cfg,
                                act_cfg=None)
        if act_cfg_out:
            self.act = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        y = self.conv3(self.conv2(self.conv1(x)))
        y = y + (self.downsample(x) if self.downsample is not None else x)
        return self.act(y) if hasattr(self, 'act') else y


class DAPPM(nn.Module):
    """ppm.py:12-130.  All ConvModules are pre-activation (norm, act, conv)."""

    def __init__(self, in_channels, branch_channels, out_channels,
                 num_scales=5, kernel_sizes=(5, 9, 17), strides=(2, 4, 8),
                 paddings=(2, 4, 8)):
        super().__init__()
        self.num_scales = num_scales
        kw = dict(norm_cfg=dict(type='BN', momentum=0.1), act_cfg=RELU,
                  order=('norm', 'act', 'conv'), bias=False)

        def pre(cin, cout, k, p=0):
            return ConvModule(cin, cout, k, padding=p, **kw)

        scales = [pre(in_channels, branch_channels, 1)]
        for i in range(1, num_scales - 1):
            scales.append(nn.Sequential(
                nn.AvgPool2d(kernel_sizes[i - 1], strides[i - 1],
                             paddings[i - 1]),
                pre(in_channels, branch_channels, 1)))
        scales.append(nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)),
                                    pre(in_channels, branch_channels, 1)))
        self.scales = nn.ModuleList(scales)
        self.processes = nn.ModuleList(
            [pre(branch_channels, branch_channels, 3, 1)
             for _ in range(num_scales - 1)])
        self.compression = pre(branch_channels * num_scales, out_channels, 1)
        self.shortcut = pre(in_channels, out_channels, 1)

    def forward(self, x):
        feats = [self.scales[0](x)]
        for i in range(1, self.num_scales):
            up = F.interpolate(self.scales[i](x), size=x.shape[2:],
                               mode='bilinear')
            feats.append(self.processes[i - 1](up + feats[i - 1]))
        return self.compression(torch.cat(feats, 1)) + self.shortcut(x)


class OracleLEDNet(nn.Module):
    """R0 body; ctor signature from configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:24-30."""

    def __init__(self, in_channels=3, channels=32, ppm_channels=128,
                 align_corners=False, norm_cfg=BN, act_cfg=RELU, init_cfg=None):
        super().__init__()
        C = channels
        self.norm_cfg, self.act_cfg = norm_cfg, act_cfg
        self.align_corners = align_corners
        # ddrnet.py:121-149
        self.stem = nn.Sequential(
            ConvModule(in_channels, C, 3, 2, 1, norm_cfg=norm_cfg, act_cfg=act_cfg),
            ConvModule(C, C, 3, 2, 1, norm_cfg=norm_cfg, act_cfg=act_cfg),
            self._make_layer(BasicBlock, C, C, 2), nn.ReLU(),
            self._make_layer(BasicBlock, C, 2 * C, 2, stride=2), nn.ReLU())
        self.relu = nn.ReLU()
        # ddrnet.py:56-65 (context) and :107-116 (spatial)
        self.context_branch_layers = nn.ModuleList([
            self._make_layer(BasicBlock, 2 * C, 4 * C, 2, stride=2),
            self._make_layer(BasicBlock, 4 * C, 8 * C, 2, stride=2),
            self._make_layer(Bottleneck, 8 * C, 8 * C, 1, stride=2)])
        # ddrnet.py:68-105 bilateral fusion convs
        self.compression_1 = ConvModule(4 * C, 2 * C, 1, norm_cfg=norm_cfg, act_cfg=None)
        self.down_1 = ConvModule(2 * C, 4 * C, 3, 2, 1, norm_cfg=norm_cfg, act_cfg=None)
        self.compression_2 = ConvModule(8 * C, 2 * C, 1, norm_cfg=norm_cfg, act_cfg=None)
        self.down_2 = nn.Sequential(
            ConvModule(2 * C, 4 * C, 3, 2, 1, norm_cfg=norm_cfg, act_cfg=act_cfg),
            ConvModule(4 * C, 8 * C, 3, 2, 1, norm_cfg=norm_cfg, act_cfg=None))
        self.spatial_branch_layers = nn.ModuleList([
            self._make_layer(BasicBlock, 2 * C, 2 * C, 2),
            self._make_layer(BasicBlock, 2 * C, 2 * C, 2),
            self._make_layer(Bottleneck, 2 * C, 2 * C, 1)])
        self.spp = DAPPM(16 * C, ppm_channels, 4 * C, num_scales=5)

    def _make_layer(self, block, inplanes, planes, num_blocks, stride=1):
        """ddrnet.py:151-180: first block default BN + ReLU out, later blocks
        norm_cfg, and no output act on the last one."""
        down = None
        if stride != 1 or inplanes != planes * block.expansion:
            down = nn.Sequential(
                nn.Conv2d(inplanes, planes * block.expansion, 1, stride, bias=False),
                build_norm_layer(self.norm_cfg, planes * block.expansion)[1])
        blocks = [block(inplanes, planes, stride=stride, downsample=down)]
        for i in range(1, num_blocks):
            blocks.append(block(planes * block.expansion, planes, stride=1,
                                norm_cfg=self.norm_cfg,
                                act_cfg_out=None if i == num_blocks - 1 else self.act_cfg))
        return nn.Sequential(*blocks)

    def forward(self, x):
        size8 = (math.ceil(x.shape[-2] / 8), math.ceil(x.shape[-1] / 8))  # ddrnet.py:185
        x1 = self.stem[0](x)          # tap C @ 1/2
        x2 = self.stem[1](x1)         # tap C @ 1/4
        x = x2
        for m in list(self.stem)[2:]:
            x = m(x)
        ac = self.align_corners
        # stage 3 (ddrnet.py:190-201)
        x_c = self.context_branch_layers[0](x)
        x_s = self.spatial_branch_layers[0](x)
        comp = self.compression_1(self.relu(x_c))
        x_c = x_c + self.down_1(self.relu(x_s))
        x_s = x_s + resize(comp, size8, ac)
        c3 = x_s
        # stage 4 (ddrnet.py:203-212)
        x_c = self.context_branch_layers[1](self.relu(x_c))
        x_s = self.spatial_branch_layers[1](self.relu(x_s))
        comp = self.compression_2(self.relu(x_c))
        x_c = x_c + self.down_2(self.relu(x_s))
        x_s = x_s + resize(comp, size8, ac)
        # stage 5 (ddrnet.py:214-224)
        x_s = self.spatial_branch_layers[2](self.relu(x_s))
        x_c = self.spp(self.context_branch_layers[2](self.relu(x_c)))
        c5 = x_s + resize(x_c, size8, ac)
        return (c3, c5, x1, x2) if self.training else (c5, x1, x2)
