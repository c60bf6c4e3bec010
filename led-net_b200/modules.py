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


def _is_sync(norm_cfg):
    return isinstance(norm_cfg, dict) and norm_cfg.get('type') == 'SyncBN'


def _mark_sync(module, flag=True):
    """Tag every BatchNorm2d under `module`: train_ops.bn_act all-reduces the batch statistics of tagged layers over
    the process group (torch.nn.SyncBatchNorm arithmetic)."""
    for m in module.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.sync = flag


def _as_nhwc(x):
    """[N,C,H,W] tensor (any memory format) -> NHWC-contiguous [N,H,W,C] (a view when the tensor is
    already channels_last, otherwise one layout kernel)."""
    from . import train_ops as T
    if x.dim() != 4:
        raise ValueError(f'expected a 4-D NCHW tensor, got shape {tuple(x.shape)}')
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    return T.to_nhwc(x.contiguous())


def _as_nchw_view(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2)


def _block_train(x, blk, out_relu):
    """BasicBlock / Bottleneck forward (basic_block.py:62-75, 206-221) on the training kernels."""
    from . import train_ops as T
    res = x
    if blk.downsample is not None:
        ds = blk.downsample
        res = T.bn_act(T.conv2d(x, ds[0].weight, None, ds[0].stride[0]), ds[1])
    y = T.conv_module(x, blk.conv1, relu=True)
    if isinstance(blk, Bottleneck):
        y = T.conv_module(y, blk.conv2, relu=True)
        return T.conv_module(y, blk.conv3, relu=out_relu, res=res)
    return T.conv_module(y, blk.conv2, relu=out_relu, res=res)


def _layer_train(x, layer):
    """ddrnet.py:151-180: the first BasicBlock of a layer ends in ReLU, the last block of a layer and
    every Bottleneck do not."""
    n = len(layer)
    for i, blk in enumerate(layer):
        out_relu = isinstance(blk, BasicBlock) and i == 0 and n > 1
        x = _block_train(x, blk, out_relu)
    return x


_DAPPM_POOLS = ((5, 2, 2), (9, 4, 4), (17, 8, 8), (0, 1, 0))     # ppm.py:66-90; k == 0: global average


def _dappm_train(x, spp):
    """DAPPM.forward (ppm.py:119-130); every ConvModule is pre-activation (norm, act, conv)."""
    from . import train_ops as T
    hw = x.shape[1:3]
    feats = [T.pre_conv_module(x, spp.scales[0])]
    for i, (k, s, p) in enumerate(_DAPPM_POOLS, start=1):
        pooled = T.avg_pool(x, k, s, p)
        up = T.resize(T.pre_conv_module(pooled, spp.scales[i][1]), hw)
        feats.append(T.pre_conv_module(T.add(up, feats[i - 1]), spp.processes[i - 1]))
    return T.add(T.pre_conv_module(T.cat_channels(feats), spp.compression), T.pre_conv_module(x, spp.shortcut))


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

    def engine_state(self):
        """This module's tensors under the names the engine expects."""
        return {self._prefix + k: v for k, v in self.state_dict().items()}

    def engine(self):
        if self._engine is None:
            state = self.engine_state()
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

    def __new__(cls, *args, variant='r0', **kwargs):
        """variant='led' builds the LED wiring (led_variant.LEDTrunk: STDC stages + GETB + MFAF + SEAM after the
        authors' speed prototype, tools/speed/ddrnet_speed.py:39-406) instead of the R0 trunk."""
        if variant == 'led' and cls is LEDNet:
            from .led_variant import LEDTrunk
            return LEDTrunk(*args, **kwargs)
        return super().__new__(cls)

    def __init__(self, in_channels=3, channels=32, ppm_channels=128, align_corners=False,
                 norm_cfg=_BN, act_cfg=_RELU, init_cfg=None, variant='r0'):
        super().__init__()
        if variant != 'r0':
            raise ValueError(f"variant must be 'r0' (the DDRNet-23-slim body + stem taps) or 'led' (the LED wiring of the "
                             f"authors' speed prototype), got {variant!r}")
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
        if _is_sync(norm_cfg):
            # Which layers are SyncBN in the reference under a SyncBN config (ddrnet.py): the stem convs, compression /
            # down convs, every downsample BN and every block but the FIRST of a layer get `norm_cfg` (:68-105, 123-138,
            # 151-180); the first block of each `_make_layer` and the DAPPM are built without it and stay plain BN.
            _mark_sync(self.stem[0]), _mark_sync(self.stem[1])
            for mod in (self.compression_1, self.down_1, self.compression_2, self.down_2):
                _mark_sync(mod)
            for layer in [self.stem[2], self.stem[4], *self.context_branch_layers, *self.spatial_branch_layers]:
                if layer[0].downsample is not None:
                    _mark_sync(layer[0].downsample)
                for blk in list(layer)[1:]:
                    _mark_sync(blk)

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
            return self._forward_train(x)
        return self.engine().backbone_forward(x)

    def _forward_train(self, x):
        """Train-mode forward (batch-statistics BatchNorm, autograd tape over the kernels of
        csrc/train.cu).  Wiring = ddrnet.py:182-224 with the two stem taps; returns
        ``(c3, c5, x1, x2)`` as led_head.py:66-75 consumes them: NCHW-shaped tensors (channels_last
        memory, so no copy is made on either side of the boundary)."""
        from . import train_ops as T
        self.reset_engine()                      # parameters are about to change
        size8 = (math.ceil(x.shape[2] / 8), math.ceil(x.shape[3] / 8))       # ddrnet.py:185
        if T.stem_conv_ok(x, self.stem[0].conv):
            # first layer straight from the NCHW image (no NHWC copy of it): csrc/train.cu stem_fwd_kernel
            x1 = T.bn_act(T.stem_conv(x, self.stem[0].conv.weight, self.stem[0].conv.stride[0]), self.stem[0].bn, relu=True)
        else:
            x1 = T.conv_module(_as_nhwc(x), self.stem[0], relu=True)
        x2 = T.conv_module(x1, self.stem[1], relu=True)
        x = T.relu(_layer_train(x2, self.stem[2]))
        x = T.relu(_layer_train(x, self.stem[4]))
        # stage 3 (ddrnet.py:190-201)
        x_c = _layer_train(x, self.context_branch_layers[0])
        x_s = _layer_train(x, self.spatial_branch_layers[0])
        comp = T.conv_module(T.relu(x_c), self.compression_1)
        x_c = T.add(x_c, T.conv_module(T.relu(x_s), self.down_1))
        x_s = T.add(x_s, T.resize(comp, size8))
        c3 = x_s
        # stage 4 (ddrnet.py:203-212)
        x_c = _layer_train(T.relu(x_c), self.context_branch_layers[1])
        x_s = _layer_train(T.relu(x_s), self.spatial_branch_layers[1])
        comp = T.conv_module(T.relu(x_c), self.compression_2)
        d = T.conv_module(T.conv_module(T.relu(x_s), self.down_2[0], relu=True), self.down_2[1])
        x_c = T.add(x_c, d)
        x_s = T.add(x_s, T.resize(comp, size8))
        # stage 5 (ddrnet.py:214-224)
        x_s = _layer_train(T.relu(x_s), self.spatial_branch_layers[2])
        x_c = _dappm_train(_layer_train(T.relu(x_c), self.context_branch_layers[2]), self.spp)
        c5 = T.add(x_s, T.resize(x_c, size8))
        return tuple(_as_nchw_view(t) for t in (c3, c5, x1, x2))


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
        if out_channels == 1 and threshold is None:                # decode_head.py:135-138
            threshold = 0.3
            warnings.warn('threshold is not defined for binary, and defaults to 0.3')
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
        if _is_sync(norm_cfg):
            _mark_sync(self)            # every BN of the head is built from norm_cfg (led_head.py:84-99)
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

    def _cls_params(self, conv):
        """`out_channels=1` (decode_head.py:119-133): the one-channel classifier map is ADDED to the num_classes-channel
        tap heads in predict_by_feat / loss_by_feat (decode_head.py:362-379, led_head.py:101-146), i.e. broadcast over
        the classes - the same as a classifier whose single filter is repeated num_classes times (the final map has
        num_classes channels, so base.py:187-192 takes the argmax branch, never the threshold one)."""
        if self.out_channels == self.num_classes:
            return conv.weight, conv.bias
        return (conv.weight.expand(self.num_classes, -1, -1, -1).contiguous(),
                conv.bias.expand(self.num_classes).contiguous() if conv.bias is not None else None)

    def engine_state(self):
        state = super().engine_state()
        if self.out_channels != self.num_classes:
            for name in ('conv_seg', 'aux_cls_seg'):
                w, b = self._cls_params(getattr(self, name))
                state[self._prefix + name + '.weight'] = w.detach().contiguous()
                state[self._prefix + name + '.bias'] = b.detach().contiguous()
        return state

    def forward(self, inputs):
        if self.training:
            return self._forward_train(inputs)
        c5, x1, x2 = inputs
        return self.engine().head_forward(c5, x1, x2)

    def _base_head_train(self, x, head):
        from . import train_ops as T
        return T.bn_act(T.pre_conv_module(x, head[0]), head[1], relu=True)      # led_head.py:84-99

    def _forward_train(self, inputs):
        """led_head.py:66-75: (context logits, spatial/aux logits, head_x1, head_x2)."""
        from . import train_ops as T
        self.reset_engine()
        c3, c5, x1, x2 = (_as_nhwc(t) for t in inputs)
        ctx = self._base_head_train(c5, self.head)
        ctx = T.conv2d(ctx, *self._cls_params(self.conv_seg))                   # cls_seg, decode_head.py:241-246
        spa = self._base_head_train(c3, self.aux_head)
        spa = T.conv2d(spa, *self._cls_params(self.aux_cls_seg))
        h1 = self._base_head_train(x1, self.head_x1)
        h2 = self._base_head_train(x2, self.head_x2)
        return tuple(_as_nchw_view(t) for t in (ctx, spa, h1, h2))

    def loss(self, inputs, batch_data_samples, train_cfg=None):
        """BaseDecodeHead.loss (decode_head.py:248-265)."""
        return self.loss_by_feat(self.forward(inputs), batch_data_samples)

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
        """led_head.py:101-146: two fused full-resolution logit maps (context, spatial), one OHEM loss
        each plus the accuracy of the context map.  Resize ladder, adds, OHEM CE and accuracy all run
        in the library's kernels (forward and backward)."""
        from . import train_ops as T
        from .losses import accuracy
        ctx, spa, h1, h2 = (_as_nhwc(t) for t in seg_logits)
        label = self._stack_batch_gt(batch_data_samples)
        hw = tuple(label.shape[2:])
        hw4, hw2 = tuple(s // 4 for s in hw), tuple(s // 2 for s in hw)

        def rung(t):
            t = T.add(h2, T.resize(t, hw4))
            return T.add(h1, T.resize(t, hw2))
        label = label.squeeze(1)
        K = ctx.shape[-1]
        fused = (K <= 32 and all(hasattr(l, 'forward_upsampled') for l in self.loss_decode[:2])
                 and self.loss_decode[0].ignore_label == self.ignore_index)
        if fused:
            # the last x2 of the ladder, the softmax / OHEM selection and (for the context map) the accuracy in one kernel
            # family: the full-resolution logits [N,K,H,W] never exist (csrc/ohem.cu, ohem_up_*)
            loss_c = self.loss_decode[0].forward_upsampled(rung(ctx), label, hw)
            acc = self.loss_decode[0].last_stats[2:3].clone()
            loss_s = self.loss_decode[1].forward_upsampled(rung(spa), label, hw)
            return dict(loss_context=loss_c, loss_spatial=loss_s, acc_seg=acc)
        ctx, spa = (T.to_nchw(T.resize(rung(t), hw)) for t in (ctx, spa))
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
