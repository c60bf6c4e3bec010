"""`LEDNet` and `LEDHead` as registered in MODELS - parameter containers whose state-dict keys
equal the reference's module paths, with forward passes that call the CUDA engine.

Reference surface mirrored
* ``LEDNet(in_channels=3, channels=32, ppm_channels=128, align_corners=False, norm_cfg, act_cfg,
  init_cfg)`` - ctor kwargs from ``configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:24-30``;
  the body source is withheld (``mmseg/models/backbones/lednet.py:1-8``), so the trunk is R0 =
  ``mmseg/models/backbones/ddrnet.py:35-224`` plus taps after ``stem[0]`` / ``stem[1]``;
  eval ``forward(x) -> (c5, x1, x2)`` as ``led_head.py:76-81`` consumes it.
* ``LEDHead(in_channels, channels, num_classes, norm_cfg, act_cfg, **BaseDecodeHead kwargs)`` -
  ``mmseg/models/decode_heads/led_head.py:29-146`` + ``decode_head.py:85-162, 241-290, 362-379``
  including its error conventions (ValueError on out_channels mismatch, TypeError on a bad
  loss_decode, warning for binary segmentation).

There is no torch arithmetic in the eval path: `forward` hands device pointers to libledb200.
"""
import math
import warnings

import torch
import torch.nn as nn

from .engine import Engine
from .registry import MODELS

_BN = dict(type='BN', requires_grad=True)
_RELU = dict(type='ReLU', inplace=True)


def _bn(c):
    return nn.BatchNorm2d(c, eps=1e-5, momentum=0.1)


class ConvModule(nn.Module):
    """Parameter holder named like mmcv's ConvModule: `.conv`, `.bn` (the activation has none)."""

    def __init__(self, cin, cout, k, stride=1, pre_act=False, bias=False):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=bias)
        self.bn = _bn(cin if pre_act else cout)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, c, stride=1, downsample=None):
        super().__init__()
        self.conv1 = ConvModule(cin, c, 3, stride)
        self.conv2 = ConvModule(c, c, 3)
        self.downsample = downsample


class Bottleneck(nn.Module):
    expansion = 2

    def __init__(self, cin, c, stride=1, downsample=None):
        super().__init__()
        self.conv1 = ConvModule(cin, c, 1)
        self.conv2 = ConvModule(c, c, 3, stride)
        self.conv3 = ConvModule(c, 2 * c, 1)
        self.downsample = downsample


def _layer(block, cin, c, n, stride=1):
    ds = None
    if stride != 1 or cin != c * block.expansion:
        ds = nn.Sequential(nn.Conv2d(cin, c * block.expansion, 1, stride, bias=False),
                           _bn(c * block.expansion))
    blocks = [block(cin, c, stride, ds)]
    blocks += [block(c * block.expansion, c) for _ in range(1, n)]
    return nn.Sequential(*blocks)


class DAPPM(nn.Module):

    def __init__(self, cin, branch, cout, num_scales=5):
        super().__init__()
        scales = [ConvModule(cin, branch, 1, pre_act=True)]
        for _ in range(1, num_scales):
            scales.append(nn.Sequential(nn.Identity(), ConvModule(cin, branch, 1, pre_act=True)))
        self.scales = nn.ModuleList(scales)
        self.processes = nn.ModuleList(
            [ConvModule(branch, branch, 3, pre_act=True) for _ in range(num_scales - 1)])
        self.compression = ConvModule(branch * num_scales, cout, 1, pre_act=True)
        self.shortcut = ConvModule(cin, cout, 1, pre_act=True)


class _EngineOwner(nn.Module):
    """Lazily (re)builds the CUDA engine from the current parameters."""
    _prefix = ''

    def __init__(self):
        super().__init__()
        self._engine = None
        self.compute_dtype = 'bf16'
        self.register_load_state_dict_post_hook(lambda m, keys: m.reset_engine())

    def reset_engine(self):
        self._engine = None

    def _engine_kwargs(self):
        raise NotImplementedError

    def engine(self):
        if self._engine is None:
            state = {self._prefix + k: v for k, v in self.state_dict().items()}
            self._engine = Engine(state, dtype=self.compute_dtype, allow_partial=True,
                                  **self._engine_kwargs())
        return self._engine

    def set_compute_dtype(self, dtype):
        assert dtype in ('bf16', 'fp32')
        if dtype != self.compute_dtype:
            self.compute_dtype = dtype
            self.reset_engine()
        return self


@MODELS.register_module()
class LEDNet(_EngineOwner):
    _prefix = 'backbone.'

    def __init__(self, in_channels=3, channels=32, ppm_channels=128, align_corners=False,
                 norm_cfg=_BN, act_cfg=_RELU, init_cfg=None, variant='r0'):
        super().__init__()
        if variant != 'r0':
            raise ValueError("only variant='r0' exists: the LED wiring is withheld upstream "
                             '(mmseg/models/backbones/lednet.py:1-8)')
        if align_corners:
            raise ValueError('align_corners=True is not supported by the B200 path')
        C = channels
        self.in_channels, self.channels, self.ppm_channels = in_channels, channels, ppm_channels
        self.align_corners, self.norm_cfg, self.act_cfg, self.init_cfg = align_corners, norm_cfg, act_cfg, init_cfg
        self.stem = nn.Sequential(ConvModule(in_channels, C, 3, 2), ConvModule(C, C, 3, 2),
                                  _layer(BasicBlock, C, C, 2), nn.ReLU(),
                                  _layer(BasicBlock, C, 2 * C, 2, 2), nn.ReLU())
        self.context_branch_layers = nn.ModuleList([
            _layer(BasicBlock, 2 * C, 4 * C, 2, 2), _layer(BasicBlock, 4 * C, 8 * C, 2, 2),
            _layer(Bottleneck, 8 * C, 8 * C, 1, 2)])
        self.compression_1 = ConvModule(4 * C, 2 * C, 1)
        self.down_1 = ConvModule(2 * C, 4 * C, 3, 2)
        self.compression_2 = ConvModule(8 * C, 2 * C, 1)
        self.down_2 = nn.Sequential(ConvModule(2 * C, 4 * C, 3, 2), ConvModule(4 * C, 8 * C, 3, 2))
        self.spatial_branch_layers = nn.ModuleList([
            _layer(BasicBlock, 2 * C, 2 * C, 2), _layer(BasicBlock, 2 * C, 2 * C, 2),
            _layer(Bottleneck, 2 * C, 2 * C, 1)])
        self.spp = DAPPM(16 * C, ppm_channels, 4 * C)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.reset_engine()

    def _engine_kwargs(self):
        return dict(num_classes=2, channels=self.channels, ppm_channels=self.ppm_channels)

    def forward(self, x):
        if self.training:
            raise NotImplementedError(
                'train-mode LEDNet.forward needs the dgrad/wgrad kernels (SURVEY section 8a row T4); '
                'not built yet - call .eval() for inference')
        return self.engine().backbone_forward(x)


@MODELS.register_module()
class LEDHead(_EngineOwner):
    _prefix = 'decode_head.'

    def __init__(self, in_channels, channels, *, num_classes, out_channels=None, threshold=None,
                 dropout_ratio=0.1, conv_cfg=None, norm_cfg=dict(type='BN'), act_cfg=_RELU,
                 in_index=-1, input_transform=None, loss_decode=None, ignore_index=255,
                 sampler=None, align_corners=False, init_cfg=None, tap_channels=32):
        super().__init__()
        # decode_head.py:192-207
        if input_transform is not None:
            assert input_transform in ['resize_concat', 'multiple_select']
            assert isinstance(in_channels, (list, tuple)) and isinstance(in_index, (list, tuple))
            assert len(in_channels) == len(in_index)
            raise NotImplementedError('LEDHead consumes the raw backbone tuple (input_transform=None)')
        assert isinstance(in_channels, int) and isinstance(in_index, int)
        if out_channels is None:                                   # decode_head.py:119-126
            if num_classes == 2:
                warnings.warn('For binary segmentation, we suggest using `out_channels = 1` to define '
                              'the output channels of segmentor, and use `threshold` to convert '
                              '`seg_logits` into a prediction applying a threshold')
            out_channels = num_classes
        if out_channels != num_classes and out_channels != 1:      # decode_head.py:128-133
            raise ValueError('out_channels should be equal to num_classes, except binary segmentation '
                             f'set out_channels == 1 and num_classes == 2, but got out_channels='
                             f'{out_channels} and num_classes={num_classes}')
        if out_channels == 1:
            raise NotImplementedError('sigmoid/threshold heads (out_channels=1) are outside the LED-Net path')
        if dropout_ratio and dropout_ratio > 0:
            raise NotImplementedError('the LED-Net config sets dropout_ratio=0 '
                                      '(configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:35)')
        if align_corners:
            raise ValueError('align_corners=True is not supported by the B200 path')
        if loss_decode is None:
            loss_decode = [dict(type='OhemCrossEntropy', thres=0.9, min_kept=131072, loss_weight=1.0),
                           dict(type='OhemCrossEntropy', thres=0.9, min_kept=131072, loss_weight=0.4)]
        if isinstance(loss_decode, dict):                           # decode_head.py:143-151
            self.loss_decode = MODELS.build(loss_decode)
        elif isinstance(loss_decode, (list, tuple)):
            self.loss_decode = nn.ModuleList([MODELS.build(l) for l in loss_decode])
        else:
            raise TypeError(f'loss_decode must be a dict or sequence of dict, but got {type(loss_decode)}')
        self.in_channels, self.channels, self.num_classes = in_channels, channels, num_classes
        self.out_channels, self.threshold, self.dropout_ratio = out_channels, threshold, dropout_ratio
        self.norm_cfg, self.act_cfg, self.in_index = norm_cfg, act_cfg, in_index
        self.ignore_index, self.align_corners, self.tap_channels = ignore_index, align_corners, tap_channels
        self.dropout = None
        self.conv_seg = nn.Conv2d(channels, out_channels, 1)        # decode_head.py:158
        self.head = self._make_base_head(in_channels, channels)
        self.aux_head = self._make_base_head(in_channels // 2, channels)
        # led_head.py:47-48 hard-codes (32, 2); generalised to (tap_channels, num_classes)
        self.head_x1 = self._make_base_head(tap_channels, num_classes)
        self.head_x2 = self._make_base_head(tap_channels, num_classes)
        self.aux_cls_seg = nn.Conv2d(channels, out_channels, 1)
        self.init_weights()

    @staticmethod
    def _make_base_head(cin, cout):
        return nn.Sequential(ConvModule(cin, cout, 3, pre_act=True), _bn(cout), nn.ReLU(inplace=True))

    def init_weights(self):                                        # led_head.py:53-60
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.reset_engine()

    def _engine_kwargs(self):
        return dict(num_classes=self.num_classes, channels=self.tap_channels,
                    head_channels=self.channels)

    def forward(self, inputs):
        if self.training:
            raise NotImplementedError(
                'train-mode LEDHead.forward needs the dgrad/wgrad kernels (SURVEY section 8a row T4); '
                'loss_by_feat() on given logits is available')
        c5, x1, x2 = inputs
        return self.engine().head_forward(c5, x1, x2)

    def predict(self, inputs, batch_img_metas=None, test_cfg=None):
        return self.predict_by_feat(self.forward(inputs), batch_img_metas)

    def predict_by_feat(self, seg_logits, batch_img_metas=None):
        """decode_head.py:362-379 (img_shape is ignored there too): full-resolution logits."""
        from .ops import head_fuse_argmax
        xc, h1, h2 = seg_logits
        return head_fuse_argmax(xc, h2, h1, want_logits=True)[1]

    def _stack_batch_gt(self, batch_data_samples):                 # decode_head.py:286-290
        return torch.stack([s['gt_sem_seg']['data'] if isinstance(s, dict) else s.gt_sem_seg.data
                            for s in batch_data_samples], dim=0)

    def loss_by_feat(self, seg_logits, batch_data_samples):
        """led_head.py:101-146.  The OHEM loss + accuracy run in the CUDA kernel; the train-time
        resize ladder still uses ATen's differentiable interpolate (its backward kernel is a
        'next' row)."""
        import torch.nn.functional as F
        from .losses import accuracy
        ctx, spa, h1, h2 = seg_logits
        label = self._stack_batch_gt(batch_data_samples)
        hw = label.shape[2:]

        def ladder(t):
            t = h2 + F.interpolate(t, size=tuple(s // 4 for s in hw), mode='bilinear', align_corners=False)
            t = h1 + F.interpolate(t, size=tuple(s // 2 for s in hw), mode='bilinear', align_corners=False)
            return F.interpolate(t, size=tuple(hw), mode='bilinear', align_corners=False)
        ctx, spa = ladder(ctx), ladder(spa)
        label = label.squeeze(1)
        return dict(loss_context=self.loss_decode[0](ctx, label),
                    loss_spatial=self.loss_decode[1](spa, label),
                    acc_seg=accuracy(ctx, label, ignore_index=self.ignore_index))


def build_param_shapes(channels=32, ppm_channels=128, head_channels=64, num_classes=2):
    """name -> shape for every 'backbone.*' / 'decode_head.*' tensor of the R0 model."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        bb = LEDNet(3, channels, ppm_channels)
        hd = LEDHead(4 * channels, head_channels, num_classes=num_classes, dropout_ratio=0.,
                     tap_channels=channels)
    out = {'backbone.' + k: tuple(v.shape) for k, v in bb.state_dict().items()}
    out.update({'decode_head.' + k: tuple(v.shape) for k, v in hd.state_dict().items()})
    return out
