"""`LEDNet(variant='led')`: the LED wiring over the four LED blocks (VERDICT r1 missing item 1).

The registered ``mmseg/models/backbones/lednet.py`` is withheld; the closest public statement of figure 3 is class
``DDRNet1`` of the authors' speed prototype, ``tools/speed/ddrnet_speed.py:39-406``: the DDRNet stem, STDC stages
(``mmseg/models/backbones/stdc.py:15-130``, fusion 'cat') in both branches, a GETB block after each context stage and
after the DAPPM (``UNetFormer_GETB.py:209-226``), ``Muti_AFF`` as the bilateral fusion (``classification/model_utils.py``),
the SEAM edge gate on the spatial branch, ``x_s + resize(x_c)`` as the output.  ``LEDTrunk`` keeps the prototype's module
/ parameter names (its checkpoints load with ``strict=True``) and adds the two stem taps ``LEDHead`` consumes:
eval ``forward(x) -> (c5, x1, x2)``.

Every layer is one C-ABI call into libledb200 (conv layer handles with the BatchNorm folded and the weights resident on
the device - tcgen05 kernel for bf16 tensors of eligible shape, CUDA cores otherwise -, depthwise / pooling / resize /
add kernels of csrc/glue.cu, the GETB / MFAF / SEAM block kernels); torch supplies memory and views only.  Channel
concats (STDC, DAPPM) are slices of one buffer the producers write into.  Eval only: the four blocks have no train mode
yet (README status table).
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import lib as L
from .getb import GETBBlock
from .mfaf import Muti_AFF
from .modules import ConvModule, BasicBlock, DAPPM, _layer, _bn
from .seam import SEAM

_DAPPM_POOLS = ((5, 2, 2), (9, 4, 4), (17, 8, 8), (0, 1, 0))     # ppm.py:66-90; k == 0: global average


class _DWConvModule(nn.Module):
    """ConvModule(c, c, 3, stride=2, padding=1, groups=c, norm_cfg, act_cfg=None) (stdc.py:52-61)."""

    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride, 1, groups=c, bias=False)
        self.bn = _bn(c)


class STDCModule(nn.Module):
    """Parameter container named like mmseg's STDCModule (stdc.py:15-94), fusion_type 'cat'."""

    def __init__(self, in_channels, out_channels, stride, norm_cfg=None, act_cfg=None, num_convs=4, fusion_type='cat'):
        super().__init__()
        assert num_convs > 1
        if fusion_type != 'cat':
            raise NotImplementedError("the LED wiring uses STDC stages with fusion_type='cat' (ddrnet_speed.py:122-123)")
        if out_channels % (8 * 2 ** (num_convs - 1)):
            raise NotImplementedError('STDCModule: every concat slice must hold a multiple of 8 channels')
        self.stride, self.with_downsample, self.out_channels = stride, stride == 2, out_channels
        self.layers = nn.ModuleList([ConvModule(in_channels, out_channels // 2, 1)])
        if self.with_downsample:
            self.downsample = _DWConvModule(out_channels // 2, 2)
            self.skip = nn.AvgPool2d(kernel_size=3, stride=2, padding=1)
        for i in range(1, num_convs):
            out_factor = 2 ** (i + 1) if i != num_convs - 1 else 2 ** i
            self.layers.append(ConvModule(out_channels // 2 ** i, out_channels // out_factor, 3))


def _fold_post(conv, bn):
    """conv -> BN  =>  (w * a, (conv.bias) * a + b)   (float64 fold, fp32 result)."""
    a = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    b = bn.bias.detach().double() - bn.running_mean.detach().double() * a
    w = conv.weight.detach().double() * a.view(-1, 1, 1, 1)
    if conv.bias is not None:
        b = b + conv.bias.detach().double() * a
    return w.float().cpu().contiguous(), b.float().cpu().contiguous()


def _fold_pre(bn):
    a = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    return a.float().cpu().contiguous(), (bn.bias.detach().double() - bn.running_mean.detach().double() * a).float().cpu().contiguous()


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _PreactConv(C.Structure):
    """ledb200_preact_conv (include/ledb200.h): host pointers of one pre-activation ConvModule"""
    _fields_ = [('weight', C.c_void_p), ('bias', C.c_void_p), ('bn_scale', C.c_void_p), ('bn_shift', C.c_void_p)]


def _geom(t):
    """(pixel stride in elements) of an NHWC tensor or channel-slice view; the layout must be pixel-major."""
    n, h, w, c = t.shape
    ld = t.stride(2)
    assert t.stride(3) == 1 and t.stride(1) == w * ld and (n == 1 or t.stride(0) == h * w * ld), 'not an NHWC (slice) view'
    return ld


class LEDTrunk(nn.Module):
    """`LEDNet(variant='led')`.  Constructor kwargs as LEDNet's (configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:24-30)."""
    composed = True          # EncoderDecoder: layer-by-layer module, not the fused R0 engine plan

    def __init__(self, in_channels=3, channels=32, ppm_channels=128, align_corners=False, norm_cfg=None, act_cfg=None,
                 init_cfg=None, num_convs=4):
        super().__init__()
        if align_corners:
            raise ValueError('align_corners=True is not supported by the B200 path')
        if channels != 32:
            raise NotImplementedError('the LED wiring fixes its stage widths to the 32-channel stem (ddrnet_speed.py:82-83)')
        Cc = channels
        self.in_channels, self.channels, self.ppm_channels, self.align_corners = in_channels, channels, ppm_channels, False
        self.variant, self.compute_dtype = 'led', 'bf16'
        ctx, spa = (64, 128, 256, 512), (64, 64, 64, 128)                       # ddrnet_speed.py:82-83
        self.gltb1 = GETBBlock(dim=128, num_heads=8, window_size=8)
        self.gltb2 = GETBBlock(dim=256, num_heads=8, window_size=8)
        self.gltb3 = GETBBlock(dim=128, num_heads=8, window_size=8)
        self.aff1 = Muti_AFF(channels=2 * Cc)
        self.aff2 = Muti_AFF(channels=2 * Cc)
        seam = SEAM(64)                                                         # conv_1 / conv_2 / fusion_kernel live at top level
        object.__setattr__(self, '_seam', seam)
        self.fusion_kernel = seam.fusion_kernel
        self.conv_1, self.conv_2 = seam.conv_1, seam.conv_2
        self.stem = nn.Sequential(ConvModule(in_channels, Cc, 3, 2), ConvModule(Cc, Cc, 3, 2),
                                  _layer(BasicBlock, Cc, Cc, 2), nn.ReLU(),
                                  _layer(BasicBlock, Cc, 2 * Cc, 2, 2), nn.ReLU())
        self.relu = nn.ReLU()
        self.context_branch_layers = nn.ModuleList(
            [nn.Sequential(STDCModule(ctx[i], ctx[i + 1], 2, num_convs=num_convs)) for i in range(3)])
        self.compression_aff = ConvModule(4 * Cc, 2 * Cc, 1)
        self.down_1 = ConvModule(2 * Cc, 4 * Cc, 3, 2)
        self.compression_2 = ConvModule(8 * Cc, 2 * Cc, 1)
        self.down_2 = nn.Sequential(ConvModule(2 * Cc, 4 * Cc, 3, 2), ConvModule(4 * Cc, 8 * Cc, 3, 2))
        self.spatial_branch_layers = nn.ModuleList(
            [nn.Sequential(STDCModule(spa[i], spa[i + 1], 1, num_convs=num_convs)) for i in range(3)])
        self.spp = DAPPM(16 * Cc, ppm_channels, 4 * Cc)
        self._layers = {}
        self._dappm_handle = None
        self.register_load_state_dict_post_hook(lambda m, keys: m.reset_engine())

    # ------------------------------------------------------------------ engine-like surface
    def reset_engine(self):
        lib = L.get()
        for h in self._layers.values():
            lib.ledb200_conv_layer_destroy(h)
        self._layers = {}
        if self._dappm_handle is not None:
            lib.ledb200_dappm_destroy(self._dappm_handle)
            self._dappm_handle = None
        self._seam.reset_engine()

    def __del__(self):
        try:
            self.reset_engine()
        except Exception:
            pass

    def set_compute_dtype(self, dtype):
        assert dtype in ('bf16', 'fp32')
        self.compute_dtype = dtype
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("LEDNet(variant='led') is eval-only: GETB / MFAF / SEAM have no train mode yet")
        self._seam.training = False              # not a registered sub-module (its parameters live at this level)
        return super().train(mode)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.reset_engine()

    # ------------------------------------------------------------------ layer calls
    def _handle(self, key, conv, bn, pre=False):
        """device-resident folded layer, built once per parameter set"""
        if key not in self._layers:
            if pre:                                   # order (norm, act, conv): BN + ReLU in the conv prologue
                w, b = conv.weight.detach().float().cpu().contiguous(), None
                ps, pb = _fold_pre(bn)
            else:
                w, b = _fold_post(conv, bn) if bn is not None else (conv.weight.detach().float().cpu().contiguous(), None)
                ps = pb = None
            h = C.c_void_p()
            L.check(L.get().ledb200_conv_layer_create(_vp(w), _vp(b), _vp(ps), _vp(pb), conv.in_channels, conv.out_channels,
                                                      conv.kernel_size[0], conv.stride[0], conv.groups, C.byref(h)),
                    'ledb200_conv_layer_create')
            self._layers[key] = h
        return self._layers[key]

    def _conv(self, key, conv, bn, x, relu=False, res=None, out=None, pre=False):
        n, h, w, _ = x.shape
        k, s = conv.kernel_size[0], conv.stride[0]
        ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
        if out is None:
            out = torch.empty((n, ho, wo, conv.out_channels), dtype=x.dtype, device=x.device)
        assert tuple(out.shape) == (n, ho, wo, conv.out_channels)
        L.check(L.get().ledb200_conv_layer_forward(self._handle(key, conv, bn, pre), _vp(x), _vp(out), _vp(res),
                                                   L.torch_dtype_code(x), n, h, w, _geom(x), _geom(out),
                                                   _geom(res) if res is not None else 0, int(relu), 0,
                                                   L.stream_ptr(x.device)), f'conv layer {key}')
        return out

    def _cm(self, name, mod, x, relu=False, res=None, out=None):
        """post-norm ConvModule `mod` registered at path `name`"""
        return self._conv(name, mod.conv, mod.bn, x, relu, res, out)

    @staticmethod
    def _avgpool(x, k, s, p, out=None):
        n, h, w, c = x.shape
        ho, wo = ((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1) if k else (1, 1)
        if out is None:
            out = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        L.check(L.get().ledb200_avgpool2d(_vp(x), _vp(out), L.torch_dtype_code(x), n, h, w, c, k, s, p, _geom(x), _geom(out),
                                          L.stream_ptr(x.device)), 'ledb200_avgpool2d')
        return out

    @staticmethod
    def _resize(x, hw):
        n, h, w, c = x.shape
        out = torch.empty((n, hw[0], hw[1], c), dtype=x.dtype, device=x.device)
        L.check(L.get().ledb200_resize_bilinear(_vp(x), _vp(out), L.torch_dtype_code(x), n, h, w, hw[0], hw[1], c, _geom(x),
                                                c, L.stream_ptr(x.device)), 'ledb200_resize_bilinear')
        return out

    @staticmethod
    def _add(a, b=None, relu=False):
        n, h, w, c = a.shape
        out = torch.empty((n, h, w, c), dtype=a.dtype, device=a.device)
        L.check(L.get().ledb200_add_relu(_vp(a), _vp(b), _vp(out), L.torch_dtype_code(a), n * h * w, c, _geom(a),
                                         _geom(b) if b is not None else 0, c, int(relu), L.stream_ptr(a.device)),
                'ledb200_add_relu')
        return out

    @staticmethod
    def _block(mod, *xs):
        """GETB / MFAF / SEAM modules speak NCHW-shaped views over NHWC memory"""
        return mod(*[t.permute(0, 3, 1, 2) for t in xs]).permute(0, 2, 3, 1)

    def _basic_layer(self, name, layer, x, final_relu):
        """ddrnet.py:151-180 + the stage-level nn.ReLU that follows (folded into the last block's epilogue)"""
        nblk = len(layer)
        for i, blk in enumerate(layer):
            p = f'{name}.{i}'
            res = x
            if blk.downsample is not None:
                res = self._conv(p + '.downsample', blk.downsample[0], blk.downsample[1], x)
            y = self._cm(p + '.conv1', blk.conv1, x, relu=True)
            out_relu = (i == 0 and nblk > 1) or (i == nblk - 1 and final_relu)
            x = self._cm(p + '.conv2', blk.conv2, y, relu=out_relu, res=res)
        return x

    def _stdc(self, name, m, x):
        """STDCModule.forward_cat (stdc.py:113-130): the four stages write their slices of ONE output buffer."""
        n, h, w, _ = x.shape
        co = m.out_channels
        ho, wo = ((h - 1) // 2 + 1, (w - 1) // 2 + 1) if m.with_downsample else (h, w)
        buf = torch.empty((n, ho, wo, co), dtype=x.dtype, device=x.device)
        offs = [0, co // 2]
        for i in range(1, len(m.layers)):
            offs.append(offs[-1] + m.layers[i].conv.out_channels)
        if m.with_downsample:
            x0 = self._cm(name + '.layers.0', m.layers[0], x, relu=True)
            self._avgpool(x0, 3, 2, 1, out=buf[..., :co // 2])                      # skip (stdc.py:80, 128-129)
            cur = self._conv(name + '.downsample', m.downsample.conv, m.downsample.bn, x0)
        else:
            cur = self._cm(name + '.layers.0', m.layers[0], x, relu=True, out=buf[..., :co // 2])
        for i in range(1, len(m.layers)):
            cur = self._cm(f'{name}.layers.{i}', m.layers[i], cur, relu=True, out=buf[..., offs[i]:offs[i + 1]])
        return buf

    def _dappm_fused(self, spp, x):
        """the two-launch DAPPM of csrc/dappm.cu behind a handle (bf16, <= 8 tiles of 16 x 8 pixels per image)"""
        n, h, w, c = x.shape
        lib = L.get()
        if self._dappm_handle is None:
            keep = []

            def pc(m):
                a, b = _fold_pre(m.bn)
                wt = m.conv.weight.detach().float().cpu().contiguous()
                bias = m.conv.bias.detach().float().cpu().contiguous() if m.conv.bias is not None else None
                keep.extend([a, b, wt, bias])
                return _PreactConv(_vp(wt).value, _vp(bias).value if bias is not None else None, _vp(a).value, _vp(b).value)
            scales = (_PreactConv * 5)(pc(spp.scales[0]), *[pc(spp.scales[i][1]) for i in range(1, 5)])
            procs = (_PreactConv * 4)(*[pc(m) for m in spp.processes])
            comp, short = pc(spp.compression), pc(spp.shortcut)
            hnd = C.c_void_p()
            L.check(lib.ledb200_dappm_create(c, spp.scales[0].conv.out_channels, spp.shortcut.conv.out_channels,
                                             C.cast(scales, C.c_void_p), C.cast(procs, C.c_void_p), C.byref(comp),
                                             C.byref(short), C.byref(hnd)), 'ledb200_dappm_create')
            self._dappm_handle = hnd
        out = torch.empty((n, h, w, spp.shortcut.conv.out_channels), dtype=x.dtype, device=x.device)
        L.check(lib.ledb200_dappm_forward(self._dappm_handle, _vp(x), _vp(out), L.torch_dtype_code(x), n, h, w,
                                          L.stream_ptr(x.device)), 'ledb200_dappm_forward')
        return out

    def _dappm(self, name, spp, x):
        """DAPPM.forward (ppm.py:119-130); every ConvModule is pre-activation (norm, act, conv)."""
        n, h, w, _ = x.shape
        P = spp.scales[0].conv.out_channels
        nsc = len(spp.scales)
        if (x.is_contiguous() and nsc == 5 and
                L.get().ledb200_dappm_eligible(L.torch_dtype_code(x), n, h, w, x.shape[3], P, spp.shortcut.conv.out_channels)):
            return self._dappm_fused(spp, x)
        cat = torch.empty((n, h, w, nsc * P), dtype=x.dtype, device=x.device)
        prev = self._conv(name + '.scales.0', spp.scales[0].conv, spp.scales[0].bn, x, out=cat[..., :P], pre=True)
        for i, (k, s, p) in enumerate(_DAPPM_POOLS, start=1):
            pooled = self._avgpool(x, k, s, p)
            sc = self._conv(f'{name}.scales.{i}.1', spp.scales[i][1].conv, spp.scales[i][1].bn, pooled, pre=True)
            t = self._add(self._resize(sc, (h, w)), prev)
            prev = self._conv(f'{name}.processes.{i - 1}', spp.processes[i - 1].conv, spp.processes[i - 1].bn, t,
                              out=cat[..., i * P:(i + 1) * P], pre=True)
        y = self._conv(name + '.compression', spp.compression.conv, spp.compression.bn, cat, pre=True)
        return self._conv(name + '.shortcut', spp.shortcut.conv, spp.shortcut.bn, x, res=y, pre=True)

    # ------------------------------------------------------------------ forward (ddrnet_speed.py:272-406, eval branch)
    @torch.no_grad()
    def forward(self, x):
        if self.training:
            raise NotImplementedError("LEDNet(variant='led') is eval-only")
        if not x.is_cuda:
            raise L.LedB200Error("LEDNet(variant='led') needs a CUDA tensor (no CPU fallback)")
        dt = torch.bfloat16 if self.compute_dtype == 'bf16' else torch.float32
        out_size = (math.ceil(x.shape[-2] / 8), math.ceil(x.shape[-1] / 8))
        # stem conv 0 straight from the NCHW fp32 image (bf16: the tensor-core stem kernel of the R0 engine)
        xin = x.float().contiguous()
        n, _, hi, wi = xin.shape
        x1 = torch.empty((n, (hi - 1) // 2 + 1, (wi - 1) // 2 + 1, self.channels), dtype=dt, device=x.device)
        L.check(L.get().ledb200_conv_layer_forward_image(
            self._handle('stem.0', self.stem[0].conv, self.stem[0].bn), _vp(xin), L.IMG_NCHW_F32, _vp(x1),
            L.torch_dtype_code(x1), n, hi, wi, self.channels, 1, L.stream_ptr(x.device)), 'stem.0')
        x2 = self._cm('stem.1', self.stem[1], x1, relu=True)
        t = self._basic_layer('stem.2', self.stem[2], x2, True)
        xs8 = self._basic_layer('stem.4', self.stem[4], t, True)                    # 1/8, 64 channels
        # stage 3
        x_c = self._stdc('context_branch_layers.0.0', self.context_branch_layers[0][0], xs8)
        x_c = self._block(self.gltb1, x_c)
        x_s = self._stdc('spatial_branch_layers.0.0', self.spatial_branch_layers[0][0], xs8)
        comp = self._cm('compression_aff', self.compression_aff, self._add(x_c, relu=True))
        x_c = self._cm('down_1', self.down_1, self._add(x_s, relu=True), res=x_c)
        x_s = self._block(self.aff1, x_s, self._resize(comp, out_size))
        # stage 4
        x_c = self._stdc('context_branch_layers.1.0', self.context_branch_layers[1][0], self._add(x_c, relu=True))
        x_c = self._block(self.gltb2, x_c)
        x_s = self._stdc('spatial_branch_layers.1.0', self.spatial_branch_layers[1][0], self._add(x_s, relu=True))
        comp = self._cm('compression_2', self.compression_2, self._add(x_c, relu=True))
        d = self._cm('down_2.0', self.down_2[0], self._add(x_s, relu=True), relu=True)
        x_c = self._cm('down_2.1', self.down_2[1], d, res=x_c)
        x_s = self._block(self.aff2, x_s, self._resize(comp, out_size))
        x_s = self._block(self._seam, xs8, x_s)                                     # edge gate (:282-338, 388-389)
        # stage 5
        x_s = self._stdc('spatial_branch_layers.2.0', self.spatial_branch_layers[2][0], self._add(x_s, relu=True))
        x_c = self._stdc('context_branch_layers.2.0', self.context_branch_layers[2][0], self._add(x_c, relu=True))
        x_c = self._dappm('spp', self.spp, x_c)
        x_c = self._block(self.gltb3, x_c)
        c5 = self._add(x_s, self._resize(x_c, out_size))
        return tuple(t.permute(0, 3, 1, 2) for t in (c5, x1, x2))
