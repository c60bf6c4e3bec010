"""Oracle LEDHead and the 3-level logit fusion (TEST INFRASTRUCTURE).

Follows ``mmseg/models/decode_heads/led_head.py:29-146`` (head modules, eval/
train forward, ``loss_by_feat``) and the author's patch in the base class,
``mmseg/models/decode_heads/decode_head.py:362-379`` (``predict_by_feat``).

One generalisation: the reference hard-codes 2 output channels for
``head_x1``/``head_x2`` (``led_head.py:47-48``), which only broadcasts against
``num_classes == 2``.  Here they produce ``num_classes`` channels; at
``num_classes == 2`` the module is identical to the reference (pinned by
``tests/golden``).
"""
import math

import torch
import torch.nn as nn

from .mmcv_shim import ConvModule, build_norm_layer, build_activation_layer
from .r0 import resize
from .losses import ohem_cross_entropy, accuracy


def fuse_logits(x_c, head_x1, head_x2, align_corners=False):
    """decode_head.py:362-379: size = ceil(2*head_x1.HW); r = head_x2 + up(x_c);
    r = head_x1 + up(r); return up(r -> size).  ``img_shape`` is ignored there."""
    size = tuple(math.ceil(s * 2) for s in head_x1.shape[2:])
    r = head_x2 + resize(x_c, tuple(math.ceil(s / 4) for s in size), align_corners)
    r = head_x1 + resize(r, tuple(math.ceil(s / 2) for s in size), align_corners)
    return resize(r, size, align_corners)


def fuse_logits_train(logit, head_x1, head_x2, label_hw, align_corners=False):
    """led_head.py:105-138: same ladder but sizes derive from the label (// not ceil)."""
    h, w = label_hw
    r = head_x2 + resize(logit, (h // 4, w // 4), align_corners)
    r = head_x1 + resize(r, (h // 2, w // 2), align_corners)
    return resize(r, (h, w), align_corners)


class OracleLEDHead(nn.Module):

    def __init__(self, in_channels=128, channels=64, num_classes=2,
                 norm_cfg=dict(type='BN'), act_cfg=dict(type='ReLU', inplace=True),
                 align_corners=False, ignore_index=255, dropout_ratio=0.0,
                 loss_decode=None, tap_channels=32):
        super().__init__()
        assert dropout_ratio == 0.0, 'LED-Net config sets dropout_ratio=0.'
        self.in_channels, self.channels = in_channels, channels
        self.num_classes = self.out_channels = num_classes
        self.norm_cfg, self.act_cfg = norm_cfg, act_cfg
        self.align_corners, self.ignore_index = align_corners, ignore_index
        self.loss_decode = loss_decode or [
            dict(thres=0.9, min_kept=131072, loss_weight=1.0),
            dict(thres=0.9, min_kept=131072, loss_weight=0.4)]
        self.conv_seg = nn.Conv2d(channels, num_classes, 1)       # decode_head.py:158
        self.head = self._base_head(in_channels, channels)        # led_head.py:44
        self.aux_head = self._base_head(in_channels // 2, channels)
        self.head_x1 = self._base_head(tap_channels, num_classes)  # ref: (32, 2)
        self.head_x2 = self._base_head(tap_channels, num_classes)
        self.aux_cls_seg = nn.Conv2d(channels, num_classes, 1)
        self.init_weights()

    def _base_head(self, cin, cout):
        """led_head.py:84-99: BN(in)->ReLU->Conv3x3(no bias) -> BN(out) -> ReLU."""
        return nn.Sequential(
            ConvModule(cin, cout, 3, padding=1, norm_cfg=self.norm_cfg,
                       act_cfg=self.act_cfg, order=('norm', 'act', 'conv')),
            build_norm_layer(self.norm_cfg, cout)[1],
            build_activation_layer(self.act_cfg))

    def init_weights(self):
        """led_head.py:53-60."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, inputs):
        if self.training:                                         # led_head.py:66-75
            c3, c5, x1, x2 = inputs
            return (self.conv_seg(self.head(c5)), self.aux_cls_seg(self.aux_head(c3)),
                    self.head_x1(x1), self.head_x2(x2))
        c5, x1, x2 = inputs                                       # led_head.py:76-81
        h1, h2 = self.head_x1(x1), self.head_x2(x2)
        return self.conv_seg(self.head(c5)), h1, h2

    def predict(self, inputs):
        x_c, h1, h2 = self.forward(inputs)
        return fuse_logits(x_c, h1, h2, self.align_corners)

    def loss(self, inputs, seg_label):
        """led_head.py:101-146.  seg_label: int64 [N,H,W]."""
        ctx, spa, h1, h2 = self.forward(inputs)
        hw = seg_label.shape[-2:]
        ctx = fuse_logits_train(ctx, h1, h2, hw, self.align_corners)
        spa = fuse_logits_train(spa, h1, h2, hw, self.align_corners)
        l0, l1 = self.loss_decode
        return dict(
            loss_context=ohem_cross_entropy(ctx, seg_label, ignore_label=self.ignore_index, **l0),
            loss_spatial=ohem_cross_entropy(spa, seg_label, ignore_label=self.ignore_index, **l1),
            acc_seg=accuracy(ctx, seg_label, ignore_index=self.ignore_index))
