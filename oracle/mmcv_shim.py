"""Restatement of the three mmcv helpers on the hot path (TEST INFRASTRUCTURE).

mmcv (``>=2.0.0rc4,<2.2.0``; ``/root/reference/mmseg/__init__.py:10-11``,
``requirements/mminstall.txt:1``) is a third-party dependency that is not under
``/root/reference`` and not installed.  Call sites this follows:
``mmseg/models/decode_heads/led_head.py:87-96``,
``mmseg/models/utils/basic_block.py:43-57,186-201``,
``mmseg/models/backbones/ddrnet.py:68-105,123-138,161``,
``mmseg/models/utils/ppm.py:57-117``.

Published semantics restated:
* order default ('conv','norm','act'); ``bias='auto'`` => conv bias iff no norm;
* norm channels = out_channels when norm follows conv, else in_channels;
* default activation ReLU(inplace=True); ``act_cfg=None`` => none;
* BN eps 1e-5, momentum 0.1; 'SyncBN' -> torch.nn.SyncBatchNorm (here plain BN:
  the reference itself reverts SyncBN for single-process runs,
  ``tools/analysis_tools/benchmark.py:77``);
* sub-module names ``conv`` / ``bn`` / ``activate`` (state-dict keys).
"""
import torch.nn as nn


def build_norm_layer(cfg, num_features):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    cfg.pop('requires_grad', None)
    assert typ in ('BN', 'SyncBN', 'BN2d'), typ
    cfg.setdefault('eps', 1e-5)
    return 'bn', nn.BatchNorm2d(num_features, **cfg)


def build_activation_layer(cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    assert typ == 'ReLU', typ
    return nn.ReLU(**cfg)


class ConvModule(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, dilation=1, groups=1, bias='auto', conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True,
                 order=('conv', 'norm', 'act')):
        super().__init__()
        assert conv_cfg is None
        self.order = tuple(order)
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size,
                              stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            if self.order.index('norm') > self.order.index('conv'):
                norm_channels = out_channels
            else:
                norm_channels = in_channels
            _, self.bn = build_norm_layer(norm_cfg, norm_channels)
        if self.with_activation:
            act_cfg = dict(act_cfg)
            if act_cfg['type'] == 'ReLU':
                act_cfg.setdefault('inplace', inplace)
            self.activate = build_activation_layer(act_cfg)
        # mmcv init: kaiming_normal_(a=0, fan_out, relu), norm weight 1 bias 0
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out',
                                nonlinearity='relu')
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        for layer in self.order:
            if layer == 'conv':
                x = self.conv(x)
            elif layer == 'norm' and self.with_norm:
                x = self.bn(x)
            elif layer == 'act' and self.with_activation:
                x = self.activate(x)
        return x
