"""Python face of the engine handle in libledb200.so (ledb200_create / set_param / finalize /
forward_*).  Host code is plumbing only: device memory, streams and tensors come from PyTorch;
all arithmetic on the path runs in the library's CUDA kernels."""
import ctypes as C

import torch

from . import lib as L

MEAN = (123.675, 116.28, 103.53)   # configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:14-15
STD = (58.395, 57.12, 57.375)


def _identity_like(name, shape):
    """Neutral filler for the half of the network a stand-alone module does not own."""
    leaf = name.rsplit('.', 1)[-1]
    if leaf == 'running_var':
        return torch.ones(shape)
    if leaf == 'weight' and len(shape) == 1:
        return torch.ones(shape)
    return torch.zeros(shape)


class Engine:
    """One handle per (process, GPU).  `state` maps reference state-dict names
    ('backbone.*', 'decode_head.*') to tensors; names the engine expects but `state` lacks are an
    error unless `allow_partial` (stand-alone backbone or head modules)."""

    def __init__(self, state, num_classes, channels=32, ppm_channels=128, head_channels=64,
                 dtype='bf16', device=None, mean=MEAN, std=STD, bgr_to_rgb=True,
                 conv_backend=0, allow_partial=False):
        if not torch.cuda.is_available():
            raise L.LedB200Error('no CUDA device: the LED-Net B200 path has no CPU fallback')
        self.lib = L.get()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.dtype = {'fp32': L.F32, 'bf16': L.BF16}[dtype]
        self.num_classes, self.channels = num_classes, channels
        cfg = L.Cfg(in_channels=3, channels=channels, ppm_channels=ppm_channels,
                    head_channels=head_channels, num_classes=num_classes, align_corners=0,
                    dtype=self.dtype, device=self.device.index, variant=0, conv_backend=conv_backend,
                    mean=(C.c_float * 3)(*mean), std=(C.c_float * 3)(*std), bgr_to_rgb=int(bgr_to_rgb))
        h = C.c_void_p()
        L.check(self.lib.ledb200_create(C.byref(cfg), C.byref(h)), 'ledb200_create')
        self.h = h
        n = self.lib.ledb200_num_params(self.h)
        shapes = None
        for i in range(n):
            name = self.lib.ledb200_param_name(self.h, i).decode()
            if name in state:
                t = state[name]
            elif allow_partial:
                if shapes is None:
                    from .modules import build_param_shapes
                    shapes = build_param_shapes(channels, ppm_channels, head_channels, num_classes)
                t = _identity_like(name, shapes[name])
            else:
                raise L.LedB200Error(f'state dict lacks {name}')
            t = t.detach().to('cpu', torch.float32).contiguous()
            shp = (C.c_int64 * max(1, t.dim()))(*t.shape)
            L.check(self.lib.ledb200_set_param(self.h, name.encode(), C.c_void_p(t.data_ptr()), shp,
                                               t.dim(), L.F32), f'set_param({name})')
        L.check(self.lib.ledb200_finalize(self.h), 'ledb200_finalize')

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.ledb200_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ forward calls
    def _img(self, img):
        assert img.is_cuda and img.dim() == 4, 'image batch must be a CUDA tensor [N,3,H,W] or [N,H,W,3]'
        img = img.contiguous()
        if img.dtype == torch.float32:
            assert img.shape[1] == 3
            return img, L.IMG_NCHW_F32, img.shape[0], img.shape[2], img.shape[3]
        assert img.dtype == torch.uint8
        if img.shape[1] == 3:
            return img, L.IMG_NCHW_U8, img.shape[0], img.shape[2], img.shape[3]
        assert img.shape[3] == 3
        return img, L.IMG_NHWC_U8, img.shape[0], img.shape[1], img.shape[2]

    @staticmethod
    def out_hw(H, W):
        """2 * head_x1.shape (decode_head.py:363)."""
        return 2 * ((H - 1) // 2 + 1), 2 * ((W - 1) // 2 + 1)

    def forward_infer(self, img, pred=None, pred_dtype=torch.uint8, logits=None, want_logits=False):
        img, layout, N, H, W = self._img(img)
        Ho, Wo = self.out_hw(H, W)
        if pred is None:
            pred = torch.empty((N, Ho, Wo), dtype=pred_dtype, device=img.device)
        if want_logits and logits is None:
            logits = torch.empty((N, self.num_classes, Ho, Wo), dtype=torch.float32, device=img.device)
        L.check(self.lib.ledb200_forward_infer(
            self.h, C.c_void_p(img.data_ptr()), layout, N, H, W, C.c_void_p(pred.data_ptr()),
            L.torch_dtype_code(pred), C.c_void_p(logits.data_ptr()) if logits is not None else None,
            L.stream_ptr(img.device)), 'ledb200_forward_infer')
        return (pred, logits) if (want_logits or logits is not None) else pred

    def backbone_forward(self, img):
        img, layout, N, H, W = self._img(img)
        Cc = self.channels
        h2, w2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        h4, w4 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
        h8, w8 = (h4 - 1) // 2 + 1, (w4 - 1) // 2 + 1
        kw = dict(dtype=torch.float32, device=img.device)
        c5 = torch.empty((N, 4 * Cc, h8, w8), **kw)
        x1 = torch.empty((N, Cc, h2, w2), **kw)
        x2 = torch.empty((N, Cc, h4, w4), **kw)
        L.check(self.lib.ledb200_backbone_forward(
            self.h, C.c_void_p(img.data_ptr()), layout, N, H, W, C.c_void_p(c5.data_ptr()),
            C.c_void_p(x1.data_ptr()), C.c_void_p(x2.data_ptr()), L.stream_ptr(img.device)),
            'ledb200_backbone_forward')
        return c5, x1, x2

    def head_forward(self, c5, x1, x2):
        c5, x1, x2 = (t.contiguous().float() for t in (c5, x1, x2))
        N, K = c5.shape[0], self.num_classes
        kw = dict(dtype=torch.float32, device=c5.device)
        xc = torch.empty((N, K) + tuple(c5.shape[2:]), **kw)
        h1 = torch.empty((N, K) + tuple(x1.shape[2:]), **kw)
        h2 = torch.empty((N, K) + tuple(x2.shape[2:]), **kw)
        L.check(self.lib.ledb200_head_forward(
            self.h, C.c_void_p(c5.data_ptr()), C.c_void_p(x1.data_ptr()), C.c_void_p(x2.data_ptr()), N,
            c5.shape[2], c5.shape[3], x1.shape[2], x1.shape[3], x2.shape[2], x2.shape[3],
            C.c_void_p(xc.data_ptr()), C.c_void_p(h1.data_ptr()), C.c_void_p(h2.data_ptr()),
            L.stream_ptr(c5.device)), 'ledb200_head_forward')
        return xc, h1, h2

    def head_infer(self, c5, x1, x2, pred=None, pred_dtype=torch.uint8, want_logits=False):
        """Head + fused tail on NHWC features in the engine's dtype (NCHW-shaped channels_last views are taken as they
        are): -> labels [N,2*h2,2*w2] (and fp32 logits [N,K,..] when asked)."""
        want = torch.bfloat16 if self.dtype == L.BF16 else torch.float32
        feats = []
        for t in (c5, x1, x2):
            v = t.permute(0, 2, 3, 1)                      # NCHW-shaped view over NHWC memory -> NHWC
            if v.dtype != want:
                v = v.to(want)
            feats.append(v if v.is_contiguous() else v.contiguous())
        c5n, x1n, x2n = feats
        N = c5n.shape[0]
        Ho, Wo = 2 * x1n.shape[1], 2 * x1n.shape[2]
        if pred is None:
            pred = torch.empty((N, Ho, Wo), dtype=pred_dtype, device=c5n.device)
        logits = torch.empty((N, self.num_classes, Ho, Wo), dtype=torch.float32, device=c5n.device) if want_logits else None
        L.check(self.lib.ledb200_head_infer(
            self.h, C.c_void_p(c5n.data_ptr()), C.c_void_p(x1n.data_ptr()), C.c_void_p(x2n.data_ptr()), N,
            c5n.shape[1], c5n.shape[2], x1n.shape[1], x1n.shape[2], x2n.shape[1], x2n.shape[2],
            C.c_void_p(pred.data_ptr()), L.torch_dtype_code(pred),
            C.c_void_p(logits.data_ptr()) if logits is not None else None, L.stream_ptr(c5n.device)), 'ledb200_head_infer')
        return (pred, logits) if want_logits else pred

    # ------------------------------------------------------------------ introspection
    def debug_fetch(self, name):
        shp = (C.c_int32 * 4)()
        rc = self.lib.ledb200_debug_fetch(self.h, name.encode(), None, 0, shp, L.stream_ptr(self.device))
        if shp[0] == 0:
            L.check(rc, f'debug_fetch({name})')
        out = torch.empty(tuple(shp), dtype=torch.float32)
        L.check(self.lib.ledb200_debug_fetch(self.h, name.encode(), C.c_void_p(out.data_ptr()),
                                             out.numel(), shp, L.stream_ptr(self.device)),
                f'debug_fetch({name})')
        return out

    def plan_launches(self):
        return self.lib.ledb200_plan_launches(self.h)

    def profile_ops(self, iters=5):
        n = self.plan_launches()
        ms = (C.c_float * n)()
        L.check(self.lib.ledb200_profile_ops(self.h, iters, ms, n, L.stream_ptr(self.device)), 'profile_ops')
        return [(self.lib.ledb200_op_name(self.h, i).decode(), ms[i]) for i in range(n)]

    KINDS = ('conv_direct', 'conv_tc', 'upsample_add', 'avgpool', 'affine_relu', 'tail', 'layout')

    def op_info(self):
        """[(name, kind, algorithmic FLOPs, algorithmic bytes)] of the last plan."""
        out = []
        buf = (C.c_double * 3)()
        for i in range(self.plan_launches()):
            L.check(self.lib.ledb200_op_info(self.h, i, buf), 'op_info')
            out.append((self.lib.ledb200_op_name(self.h, i).decode(), self.KINDS[int(buf[2])], buf[0], buf[1]))
        return out
