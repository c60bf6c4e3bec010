"""`OhemCrossEntropy` (MODELS) and `accuracy` - reference surface of
``mmseg/models/losses/ohem_cross_entropy_loss.py:11-94`` and ``losses/accuracy.py:6-61``,
computed by the fused CUDA kernel family in csrc/ohem.cu (radix select instead of a full sort)."""
import torch
import torch.nn as nn

from . import ops
from .registry import MODELS


class _OhemFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, score, target, mod):
        out3, grad = ops.ohem_ce(score, target, mod.ignore_label, mod.thresh, mod.min_kept,
                                 mod.loss_weight, mod.class_weight, want_grad=score.requires_grad)
        ctx.save_for_backward(grad) if grad is not None else None
        ctx.has_grad = grad is not None
        mod.last_stats = out3          # [loss, kept pixels, top-1 accuracy %]
        return out3[0].clone()

    @staticmethod
    def backward(ctx, g):
        if not ctx.has_grad:
            return None, None, None
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


class _OhemUpFn(torch.autograd.Function):
    """OhemCrossEntropy of resize(r1, size, bilinear) for an NHWC rung r1 [N,h,w,K], without the full-resolution logits
    (csrc/ohem.cu: ohem_up_*): the top of LEDHead's training ladder (led_head.py:101-146)."""

    @staticmethod
    def forward(ctx, r1, target, mod, size):
        from . import lib as L
        lib = L.get()
        r1 = r1.contiguous()
        target = target.contiguous().to(torch.int64)
        n, h, w, k = r1.shape
        H, W = int(size[0]), int(size[1])
        ws = torch.empty(lib.ledb200_ohem_workspace_bytes(n * H * W), dtype=torch.uint8, device=r1.device)
        out3 = torch.empty(3, dtype=torch.float32, device=r1.device)
        cw = None
        if mod.class_weight is not None:
            cw = torch.as_tensor(mod.class_weight, dtype=torch.float32, device=r1.device).contiguous()
        L.check(lib.ledb200_ohem_up_fwd(ops._p(r1), ops._p(target), n, k, h, w, H, W, mod.ignore_label, float(mod.thresh),
                                        int(mod.min_kept), float(mod.loss_weight), ops._p(cw), ops._p(out3), ops._p(ws),
                                        L.stream_ptr(r1.device)), 'ledb200_ohem_up_fwd')
        ctx.save_for_backward(r1, target, ws)
        ctx.cfg = (n, k, h, w, H, W, mod.ignore_label, float(mod.loss_weight), cw)
        mod.last_stats = out3
        return out3[0].clone()

    @staticmethod
    def backward(ctx, g):
        from . import lib as L
        r1, target, ws = ctx.saved_tensors
        n, k, h, w, H, W, ign, lw, cw = ctx.cfg
        d = torch.empty_like(r1)
        g = g.reshape(1).float().contiguous()
        L.check(L.get().ledb200_ohem_up_bwd(ops._p(r1), ops._p(target), n, k, h, w, H, W, ign, lw, ops._p(cw), ops._p(g),
                                            ops._p(ws), ops._p(d), L.stream_ptr(r1.device)), 'ledb200_ohem_up_bwd')
        return d, None, None, None


@MODELS.register_module()
class OhemCrossEntropy(nn.Module):

    def __init__(self, ignore_label=255, thres=0.7, min_kept=100000, loss_weight=1.0,
                 class_weight=None, loss_name='loss_ohem'):
        super().__init__()
        self.thresh = thres
        self.min_kept = max(1, min_kept)
        self.ignore_label = ignore_label
        self.loss_weight = loss_weight
        self.loss_name_ = loss_name
        self.class_weight = class_weight
        self.last_stats = None

    def forward(self, score, target):
        return _OhemFn.apply(score, target, self)

    def forward_upsampled(self, r1_nhwc, target, size):
        """loss of resize(r1, size) for an NHWC rung, fused (no full-resolution logits); K <= 32, CUDA fp32."""
        return _OhemUpFn.apply(r1_nhwc, target, self, tuple(size))

    @property
    def loss_name(self):
        return self.loss_name_


def accuracy(pred, target, topk=1, thresh=None, ignore_index=None):
    """accuracy.py:6-61 for the one mode the path uses (topk=1, thresh=None)."""
    if topk != 1 or thresh is not None:
        raise NotImplementedError('only topk=1, thresh=None is on the LED-Net path (led_head.py:143-144)')
    if pred.size(0) == 0:
        return pred.new_tensor(0.)
    assert pred.ndim == target.ndim + 1 and pred.size(0) == target.size(0)
    ign = -(2 ** 31) + 1 if ignore_index is None else ignore_index
    out3, _ = ops.ohem_ce(pred.detach(), target, ign, 0.0, 1, 1.0, None, want_grad=False)
    return out3[2:3].clone()
