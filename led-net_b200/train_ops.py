"""torch.autograd.Function wrappers over the training kernels of libledb200 (csrc/train.cu).

The reference trains through autograd over ATen/cuDNN ops (``EncoderDecoder.loss``,
``mmseg/models/segmentors/encoder_decoder.py:161-185`` -> ``LEDHead.loss``,
``decode_heads/led_head.py:101-146``).  Here every op of that graph is a hand-written kernel pair
(forward, backward); PyTorch only records the tape, owns the memory and runs NCCL.

Layout: activations are NHWC fp32 CUDA tensors ``[N,H,W,C]`` between these functions
(``to_nhwc`` / ``to_nchw`` convert at the boundary to the reference's NCHW tensors);
weights stay in the reference's OIHW state-dict layout.  There is no CPU fallback.
"""
import ctypes as C
import os

import torch

from . import lib as L


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _chk(t, name='tensor'):
    if not t.is_cuda:
        raise L.LedB200Error(f'{name}: LED-Net B200 training ops need CUDA tensors (no CPU fallback)')
    if t.dtype != torch.float32:
        raise L.LedB200Error(f'{name}: training ops are fp32, got {t.dtype}')
    return t.contiguous()


def _st(t):
    return L.stream_ptr(t.device)


def _ws(dev, c):
    """Workspace of the per-channel reductions: accumulators + per-block partials (summed in a fixed order)."""
    return torch.empty(L.get().ledb200_train_bn_workspace_bytes(c) // 8, dtype=torch.float64, device=dev)


def _out_hw(h, w, k, s):
    p = k // 2
    return (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1


class _Layout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, to_nhwc):
        x = _chk(x, 'layout')
        ctx.to_nhwc = to_nhwc
        if to_nhwc:
            n, c, h, w = x.shape
            out = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
        else:
            n, h, w, c = x.shape
            out = torch.empty((n, c, h, w), dtype=x.dtype, device=x.device)
        L.check(L.get().ledb200_train_layout(_p(x), _p(out), n, c, h, w, int(to_nhwc), _st(x)), 'train_layout')
        return out

    @staticmethod
    def backward(ctx, g):
        return _Layout.apply(g, not ctx.to_nhwc), None


def to_nhwc(x):
    return _Layout.apply(x, True)


def to_nchw(x):
    return _Layout.apply(x, False)


# Tensor-core switch of the training convolutions.
#   TENSOR_CORES  True (default): every eligible shape (ledb200_train_conv_tc_ok) runs on the tcgen05 kind::tf32 kernels;
#                 False: the fp32 CUDA-core kernels everywhere (the round-1 path, still used for shapes the tiling rejects).
#   TC_FAST       False (default): error-compensated three-pass products (3 x TF32) - fp32-grade results, which the 1e-2
#                 gradient gate needs (train-mode BatchNorm makes this network's gradient ill-conditioned: tf32 operand
#                 rounding alone moves it by tens of percent, tools/diag_tf32_grads.py);
#                 True: one tf32 pass (what cuDNN does under torch.backends.cudnn.allow_tf32) with tf32-rounded activation
#                 storage, because the tensor core truncates raw fp32 operands.
# Environment: LEDB200_TRAIN_TC = 0 | 1 | fast.
_env = os.environ.get('LEDB200_TRAIN_TC', '1')
TENSOR_CORES = _env != '0'
TC_FAST = _env == 'fast'
# Passes of the weight-gradient kernel: 1 by default (a leaf of the backward pass, a sum over 1e5..1e6 pixels: one tf32 pass
# leaves a uniform ~7e-4 shrink, see ledb200_train_set_wgrad_passes); 3 = fp32-grade.  Environment: LEDB200_WGRAD_PASSES.
WGRAD_PASSES = int(os.environ.get('LEDB200_WGRAD_PASSES', '1'))
_rounding_synced = None
_passes_synced = None


def set_tensor_cores(on, fast=False):
    """Select the convolution path of the training step (see above).  Returns the previous (on, fast) pair."""
    global TENSOR_CORES, TC_FAST
    prev = (TENSOR_CORES, TC_FAST)
    if isinstance(on, tuple):
        on, fast = on
    TENSOR_CORES, TC_FAST = bool(on), bool(fast)
    _sync_mode()
    return prev


def compute_mode():
    """Arithmetic of the training convolutions right now: 'tf32x3' (tensor cores, three error-compensated passes, fp32-grade),
    'tf32' (tensor cores, one pass) or 'f32' (CUDA cores)."""
    return ('tf32' if TC_FAST else 'tf32x3') if TENSOR_CORES else 'f32'


def _sync_mode():
    global _passes_synced
    want = (1 if TC_FAST else 3, 1 if TC_FAST else WGRAD_PASSES)
    if _passes_synced != want:
        L.check(L.get().ledb200_train_set_tf32_passes(want[0]), 'train_set_tf32_passes')
        L.check(L.get().ledb200_train_set_wgrad_passes(want[1]), 'train_set_wgrad_passes')
        _passes_synced = want


def _round_for(shape):
    """tf32 storage for the tensor an element-wise kernel is about to write ([N, H, W, C]): only in the single-pass mode,
    and only for tensors a tensor-core convolution can consume (extents that tile into 16 x 8 pixel blocks)."""
    global _rounding_synced
    on = bool(TENSOR_CORES and TC_FAST and len(shape) == 4 and shape[1] % 16 == 0 and shape[2] % 8 == 0)
    if _rounding_synced is not on:
        L.get().ledb200_train_set_tf32_rounding(int(on))
        _rounding_synced = on


TC_OPS = 7   # diagnostic mask (tools/diag_tf32_grads.py): 1 forward, 2 data gradient, 4 weight gradient on tensor cores


_FWD_MAXCIN = int(os.environ.get('LEDB200_TC_FWD_MAXCIN', '0'))   # diagnostic: forward convolutions wider than this stay on CUDA cores


def _tc_ok(op, n, h, w, cin, cout, k, stride):
    _sync_mode()
    if op == 0 and _FWD_MAXCIN and cin > _FWD_MAXCIN:
        return False
    return (TENSOR_CORES and bool(TC_OPS & (1 << op))
            and bool(L.get().ledb200_train_conv_tc_ok(op, n, h, w, cin, cout, k, stride)))


class _Conv(torch.autograd.Function):
    """nn.Conv2d(k, stride, padding=k//2[, bias]) on NHWC activations."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        x, weight = _chk(x, 'conv input'), _chk(weight, 'conv weight')
        lib = L.get()
        n, h, w, cin = x.shape
        cout, cin_w, k, _ = weight.shape
        assert cin_w == cin, f'conv: input has {cin} channels, weight expects {cin_w}'
        ho, wo = _out_hw(h, w, k, stride)
        y = torch.empty((n, ho, wo, cout), dtype=torch.float32, device=x.device)
        b = _chk(bias, 'conv bias') if bias is not None else None
        if _tc_ok(0, n, h, w, cin, cout, k, stride):
            wp = torch.empty(lib.ledb200_train_packed_weight_tc_floats(cout, cin, k, 0), dtype=torch.float32,
                             device=x.device)
            L.check(lib.ledb200_train_pack_weight_tc(_p(weight), _p(wp), cout, cin, k, 0, _st(x)),
                    'train_pack_weight_tc')
            L.check(lib.ledb200_train_conv_fwd_tc(_p(x), _p(wp), _p(b), _p(y), n, h, w, cin, cout, k, stride,
                                                  _st(x)), 'train_conv_fwd_tc')
        else:
            wp = torch.empty(lib.ledb200_train_packed_weight_floats(cout, cin, k, 0), dtype=torch.float32,
                             device=x.device)
            L.check(lib.ledb200_train_pack_weight(_p(weight), _p(wp), cout, cin, k, 0, _st(x)), 'train_pack_weight')
            L.check(lib.ledb200_train_conv_fwd(_p(x), _p(wp), _p(b), _p(y), n, h, w, cin, cout, k, stride, _st(x)),
                    'train_conv_fwd')
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.has_bias = stride, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _chk(dy, 'conv grad')
        lib = L.get()
        n, h, w, cin = x.shape
        cout, _, k, _ = weight.shape
        dx = dw = db = None
        # An output-channel count that is not a multiple of 32 (the K-class head convs, K = 19) keeps the gradient kernels off
        # the tensor cores: dY's channels are their reduction (dgrad) / N (wgrad) side.  Zero-pad dY and the weight to the
        # next multiple of 32 - one extra pass over dY - and slice the weight gradient afterwards.
        ho, wo = _out_hw(h, w, k, ctx.stride)
        cpad = (cout + 31) // 32 * 32
        want_dx = ctx.needs_input_grad[0]
        want_dw = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        if cpad != cout and TENSOR_CORES:
            dx_tc = want_dx and _tc_ok(1, n, h, w, cin, cpad, k, ctx.stride)
            dw_tc = want_dw and not ctx.has_bias and _tc_ok(2, n, h, w, cin, cpad, k, ctx.stride)
            if dx_tc or dw_tc:
                dyp = torch.zeros((n, ho, wo, cpad), dtype=torch.float32, device=dy.device)
                L.check(lib.ledb200_train_copy_channels(_p(dy), cout, 0, _p(dyp), cpad, 0, n * ho * wo, cout, _st(dy)),
                        'train_copy_channels')
                wpad = torch.zeros((cpad, cin, k, k), dtype=torch.float32, device=weight.device)
                wpad[:cout].copy_(weight)
                dxp, dwp, _ = _Conv._backward_core(ctx, x, wpad, dyp, False, dx_tc, dw_tc)
                dx, dw, db = _Conv._backward_core(ctx, x, weight, dy, ctx.has_bias, want_dx and not dx_tc,
                                                  want_dw and not dw_tc)
                if dx_tc:
                    dx = dxp
                if dw_tc:
                    dw = dwp[:cout].contiguous()
                return dx, dw, db, None
        dx, dw, db = _Conv._backward_core(ctx, x, weight, dy, ctx.has_bias, want_dx, want_dw)
        return dx, dw, db, None

    @staticmethod
    def _backward_core(ctx, x, weight, dy, has_bias, want_dx, want_dw):
        lib = L.get()
        n, h, w, cin = x.shape
        cout, _, k, _ = weight.shape
        dx = dw = db = None
        if want_dx:
            dx = torch.empty_like(x)
            if _tc_ok(1, n, h, w, cin, cout, k, ctx.stride):
                wp = torch.empty(lib.ledb200_train_packed_weight_tc_floats(cout, cin, k, 1), dtype=torch.float32,
                                 device=x.device)
                L.check(lib.ledb200_train_pack_weight_tc(_p(weight), _p(wp), cout, cin, k,
                                                         2 if (ctx.stride == 2 and k == 3) else 1, _st(x)),
                        'train_pack_weight_tc')
                L.check(lib.ledb200_train_conv_dgrad_tc(_p(dy), _p(wp), _p(dx), n, h, w, cin, cout, k, ctx.stride,
                                                        _st(x)), 'train_conv_dgrad_tc')
            else:
                wp = torch.empty(lib.ledb200_train_packed_weight_floats(cout, cin, k, 1), dtype=torch.float32,
                                 device=x.device)
                L.check(lib.ledb200_train_pack_weight(_p(weight), _p(wp), cout, cin, k, 1, _st(x)),
                        'train_pack_weight')
                L.check(lib.ledb200_train_conv_dgrad(_p(dy), _p(wp), _p(dx), n, h, w, cin, cout, k, ctx.stride,
                                                     _st(x)), 'train_conv_dgrad')
        if want_dw and not has_bias and _tc_ok(2, n, h, w, cin, cout, k, ctx.stride):
            dw = torch.empty_like(weight)
            ws = torch.empty(lib.ledb200_train_wgrad_tc_workspace_bytes(n, h, w, cin, cout, k, ctx.stride) // 4,
                             dtype=torch.float32, device=x.device)   # per-CTA partial sums, added in a fixed order
            L.check(lib.ledb200_train_conv_wgrad_tc(_p(x), _p(dy), _p(dw), n, h, w, cin, cout, k, ctx.stride, _p(ws),
                                                    _st(x)), 'train_conv_wgrad_tc')
        elif want_dw:
            dw = torch.empty_like(weight)
            if has_bias:
                db = torch.empty(cout, dtype=torch.float32, device=x.device)
            ws = torch.empty(lib.ledb200_train_wgrad_workspace_bytes(cin, cout, k) // 8, dtype=torch.float64,
                             device=x.device)       # per-CTA partial sums, added in a fixed order (no atomics)
            L.check(lib.ledb200_train_conv_wgrad(_p(x), _p(dy), _p(dw), _p(db), n, h, w, cin, cout, k, ctx.stride,
                                                 _p(ws), _st(x)), 'train_conv_wgrad')
        return dx, dw, db


def conv2d(x, weight, bias=None, stride=1):
    return _Conv.apply(x, weight, bias, stride)


class _StemConv(torch.autograd.Function):
    """The stem's first layer on the NCHW image as the caller holds it: nn.Conv2d(Cin <= 4, Cout, 3, stride, padding=1,
    bias=False) -> NHWC output, without an NHWC copy of the image (ledb200_train_stem_fwd / _stem_wgrad).  The image gets
    no gradient (it is the network input)."""

    @staticmethod
    def forward(ctx, x, weight, stride):
        x, weight = _chk(x, 'stem input'), _chk(weight, 'stem weight')
        n, cin, h, w = x.shape
        cout, _, k, _ = weight.shape
        ho, wo = _out_hw(h, w, k, stride)
        y = torch.empty((n, ho, wo, cout), dtype=torch.float32, device=x.device)
        L.check(L.get().ledb200_train_stem_fwd(_p(x), _p(weight), None, _p(y), n, h, w, cin, cout, stride, _st(x)),
                'train_stem_fwd')
        ctx.save_for_backward(x, weight)
        ctx.stride = stride
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise L.LedB200Error('stem_conv: the image does not get a gradient (use conv2d on an NHWC tensor for that)')
        dw = None
        if ctx.needs_input_grad[1]:
            dy = _chk(dy, 'stem grad')
            lib = L.get()
            n, cin, h, w = x.shape
            cout = weight.shape[0]
            dw = torch.empty_like(weight)
            ws = torch.empty(lib.ledb200_train_wgrad_workspace_bytes(cin, cout, 3) // 8, dtype=torch.float64, device=x.device)
            L.check(lib.ledb200_train_stem_wgrad(_p(x), _p(dy), _p(dw), n, h, w, cin, cout, ctx.stride, _p(ws), _st(x)),
                    'train_stem_wgrad')
        return None, dw, None


def stem_conv_ok(x, conv):
    """whether `conv` on the NCHW image `x` can take the stem kernels (else: to_nhwc + conv2d)"""
    return (x.dim() == 4 and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] <= 4
            and not x.requires_grad
            and conv.bias is None and conv.kernel_size == (3, 3) and conv.stride[0] in (1, 2) and conv.padding == (1, 1)
            and conv.out_channels % 16 == 0 and conv.out_channels <= 128)


def stem_conv(x, weight, stride):
    return _StemConv.apply(x, weight, stride)


def _sync_group(sync):
    """The process group to synchronise BatchNorm statistics over, or None (single process / plain BN)."""
    if not sync:
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return None
    return dist.group.WORLD


class _PeerReduce:
    """SyncBN's statistics all-reduce over NVLink peer memory (ledb200_peer_allreduce_f64) instead of one NCCL call per
    layer and direction.  The symmetric buffer and its peer mappings come from torch.distributed._symmetric_memory (plumbing);
    the exchange itself is the library's kernel.  Used when the process group is NCCL over the GPUs of one box and the
    symmetric-memory rendezvous works; otherwise - or with LEDB200_SYNCBN=nccl - `reduce` falls back to dist.all_reduce."""
    NMAX = 2 * 2048 + 8
    _inst = {}

    @classmethod
    def get(cls, group, device):
        key = (id(group), device.index)
        if key not in cls._inst:
            cls._inst[key] = cls(group, device)
        return cls._inst[key]

    def __init__(self, group, device):
        import torch.distributed as dist
        self.group, self.ok, self.seq = group, False, 0
        if os.environ.get('LEDB200_SYNCBN', 'peer') != 'peer' or dist.get_backend(group) != 'nccl':
            return
        try:
            import torch.distributed._symmetric_memory as symm
            lib = L.get()
            nbytes = lib.ledb200_peer_allreduce_buffer_bytes(self.NMAX)
            self.buf = symm.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
            self.buf.zero_()
            hdl = symm.rendezvous(self.buf, group)
            self.rank, self.world = int(hdl.rank), int(hdl.world_size)
            if self.world > 16:
                return
            self.ptrs = (C.c_uint64 * 16)(*[int(p) for p in hdl.buffer_ptrs], *([0] * (16 - self.world)))
            self._hdl = hdl
            torch.cuda.synchronize(device)
            dist.barrier(group)                      # every rank's flags are zero before anyone signals
            self.ok = True
        except Exception as e:                       # no symmetric memory on this system / group: NCCL does it
            import warnings
            warnings.warn(f'SyncBN peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL')
            self.ok = False

    def reduce(self, vec):
        """in-place sum of the 1-D float64 CUDA tensor `vec` over the group"""
        if not self.ok or vec.numel() > self.NMAX:
            import torch.distributed as dist
            dist.all_reduce(vec, group=self.group)
            return
        self.seq += 1
        L.check(L.get().ledb200_peer_allreduce_f64(_p(vec), vec.numel(), self.rank, self.world, self.ptrs,
                                                   self.seq & 0xFFFFFFFF or 1, self.NMAX, _p(vec), _st(vec)),
                'peer_allreduce_f64')


class _BNAct(torch.autograd.Function):
    """out = [relu](BatchNorm2d_train(y) [+ res]); running stats updated in place.
    With `group` (SyncBN) the per-channel (sum, sum of squares, count) - and in backward (sum dz, sum dz*xhat) - are
    all-reduced over the ranks between the reduction and the apply kernels: torch.nn.SyncBatchNorm's arithmetic
    (statistics over the GLOBAL batch, parameter gradients rank-local, to be averaged by the DDP all-reduce)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, res, running_mean, running_var, momentum, eps, relu, group):
        y = _chk(y, 'bn input')
        _round_for(y.shape)
        c = y.shape[-1]
        npix = y.numel() // c
        out = torch.empty_like(y)
        mean = torch.empty(c, dtype=torch.float32, device=y.device)
        invstd = torch.empty_like(mean)
        r = _chk(res, 'bn residual') if res is not None else None
        lib = L.get()
        total = float(npix)
        if group is None:
            L.check(lib.ledb200_train_bn_fwd(_p(y), _p(gamma), _p(beta), _p(r), _p(out), _p(mean), _p(invstd),
                                             _p(running_mean), _p(running_var), float(momentum), float(eps),
                                             int(relu), npix, c, _p(_ws(y.device, c)), _st(y)), 'train_bn_fwd')
        else:
            import torch.distributed as dist
            ws = _ws(y.device, c)
            L.check(lib.ledb200_train_bn_reduce(_p(y), None, None, None, None, 0, 0, npix, c, _p(ws), _st(y)),
                    'train_bn_reduce')                # (sum, sumsq) at ws[0:2C], this rank's sample count at ws[2C]
            _PeerReduce.get(group, y.device).reduce(ws[:2 * c + 1])   # one packed message per layer: (sum, sumsq, count)
            # the all-reduced sample count stays on the device (total_count < 0: the kernels read it from the workspace) -
            # a .item() here was one host round trip per layer and direction, 5 ms of a 40 ms step at N = 2.  The workspace
            # goes to the backward pass with it (ctx.total), so no copy either.
            total = ws
            L.check(lib.ledb200_train_bn_fwd_apply(_p(y), _p(gamma), _p(beta), _p(r), _p(out), _p(mean), _p(invstd),
                                                   _p(running_mean), _p(running_var), float(momentum), float(eps),
                                                   int(relu), npix, -1.0, c, _p(ws), _st(y)), 'train_bn_fwd_apply')
        ctx.save_for_backward(y, out, gamma, mean, invstd)
        ctx.relu, ctx.has_res, ctx.group, ctx.total = relu, res is not None, group, total
        return out

    @staticmethod
    def backward(ctx, dout):
        y, out, gamma, mean, invstd = ctx.saved_tensors
        dout = _chk(dout, 'bn grad')
        _round_for(y.shape)
        c = y.shape[-1]
        npix = y.numel() // c
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        dres = None
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = torch.empty_like(y) if ctx.relu else dout
        lib = L.get()
        if ctx.group is None:
            L.check(lib.ledb200_train_bn_bwd(_p(dout), _p(y), _p(out), _p(gamma), _p(mean), _p(invstd), _p(dy),
                                             _p(dres) if ctx.relu else None, _p(dgamma), _p(dbeta), int(ctx.relu),
                                             npix, c, _p(_ws(y.device, c)), _st(y)), 'train_bn_bwd')
        else:
            import torch.distributed as dist
            ws = ctx.total                                # the forward workspace: global sample count still at [2C]
            mode = 4 if getattr(ctx, 'count_moved', False) else 3      # a repeated backward must not move the count again
            ctx.count_moved = True
            L.check(lib.ledb200_train_bn_reduce(_p(dout), _p(y), _p(out), _p(mean), _p(invstd), mode, int(ctx.relu), npix,
                                                c, _p(ws), _st(y)), 'train_bn_reduce')   # count -> [4C], sums at [0:2C] and [2C:4C]
            _PeerReduce.get(ctx.group, y.device).reduce(ws[2 * c:4 * c])
            L.check(lib.ledb200_train_bn_bwd_apply(_p(dout), _p(y), _p(out), _p(gamma), _p(mean), _p(invstd), _p(dy),
                                                   _p(dres) if ctx.relu else None, _p(dgamma), _p(dbeta),
                                                   int(ctx.relu), npix, -1.0, c, _p(ws), _st(y)),
                    'train_bn_bwd_apply')
        return dy, dgamma, dbeta, dres, None, None, None, None, None, None


def bn_act(y, bn, res=None, relu=False):
    """`bn`: an nn.BatchNorm2d in training mode (its running stats are updated like PyTorch does).  When the
    module carries `sync = True` (built from norm_cfg type 'SyncBN') and a process group of more than one rank is
    initialised, the batch statistics are those of the GLOBAL batch (torch.nn.SyncBatchNorm)."""
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return _BNAct.apply(y, bn.weight, bn.bias, res, bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu,
                        _sync_group(getattr(bn, 'sync', False)))


class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        x = _chk(x, 'resize input')
        n, h, w, c = x.shape
        H, W = int(size[0]), int(size[1])
        out = torch.empty((n, H, W, c), dtype=torch.float32, device=x.device)
        _round_for(out.shape)
        L.check(L.get().ledb200_train_resize_fwd(_p(x), _p(out), n, h, w, H, W, c, _st(x)), 'train_resize_fwd')
        ctx.shape = (n, h, w, c, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        n, h, w, c, H, W = ctx.shape
        dout = _chk(dout, 'resize grad')
        dx = torch.empty((n, h, w, c), dtype=torch.float32, device=dout.device)
        _round_for(dx.shape)
        L.check(L.get().ledb200_train_resize_bwd(_p(dout), _p(dx), n, h, w, H, W, c, _st(dout)), 'train_resize_bwd')
        return dx, None


def resize(x, size):
    """resize(mode='bilinear', align_corners=False) (mmseg/models/utils/wrappers.py:8-27)."""
    return _Resize.apply(x, tuple(size))


class _AddRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, relu):
        a = _chk(a, 'add input')
        bb = _chk(b, 'add input') if b is not None else None
        if bb is not None:
            assert a.shape == bb.shape, f'add: {tuple(a.shape)} vs {tuple(bb.shape)}'
        out = torch.empty_like(a)
        _round_for(out.shape)
        L.check(L.get().ledb200_train_add_relu(_p(a), _p(bb), _p(out), int(relu), a.numel(), _st(a)),
                'train_add_relu')
        ctx.relu, ctx.has_b = relu, b is not None
        if relu:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _chk(dout, 'add grad')
        if ctx.relu:
            (out,) = ctx.saved_tensors
            dx = torch.empty_like(out)
            _round_for(dx.shape)
            L.check(L.get().ledb200_train_relu_bwd(_p(dout), _p(out), _p(dx), out.numel(), _st(out)),
                    'train_relu_bwd')
        else:
            dx = dout
        return dx, (dx if ctx.has_b else None), None


def add(a, b, relu=False):
    return _AddRelu.apply(a, b, relu)


def relu(x):
    return _AddRelu.apply(x, None, True)


class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, s, p):
        x = _chk(x, 'pool input')
        n, h, w, c = x.shape
        ho, wo = (1, 1) if k == 0 else ((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
        out = torch.empty((n, ho, wo, c), dtype=torch.float32, device=x.device)
        _round_for(out.shape)
        L.check(L.get().ledb200_train_avgpool_fwd(_p(x), _p(out), n, h, w, c, ho, wo, k, s, p, _st(x)),
                'train_avgpool_fwd')
        ctx.cfg = (n, h, w, c, ho, wo, k, s, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        n, h, w, c, ho, wo, k, s, p = ctx.cfg
        dout = _chk(dout, 'pool grad')
        dx = torch.empty((n, h, w, c), dtype=torch.float32, device=dout.device)
        _round_for(dx.shape)
        L.check(L.get().ledb200_train_avgpool_bwd(_p(dout), _p(dx), n, h, w, c, ho, wo, k, s, p, _st(dout)),
                'train_avgpool_bwd')
        return dx, None, None, None


def avg_pool(x, k, s, p):
    """nn.AvgPool2d(k, s, p); k == 0 means nn.AdaptiveAvgPool2d((1, 1))."""
    return _AvgPool.apply(x, k, s, p)


class _Cat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        xs = [_chk(x, 'cat input') for x in xs]
        n, h, w, _ = xs[0].shape
        cs = [x.shape[-1] for x in xs]
        out = torch.empty((n, h, w, sum(cs)), dtype=torch.float32, device=xs[0].device)
        _round_for(out.shape)
        off = 0
        for x, c in zip(xs, cs):
            L.check(L.get().ledb200_train_copy_channels(_p(x), c, 0, _p(out), sum(cs), off, n * h * w, c, _st(x)),
                    'train_copy_channels')
            off += c
        ctx.cs = cs
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _chk(dout, 'cat grad')
        n, h, w, ct = dout.shape
        _round_for(dout.shape)
        outs, off = [], 0
        for c in ctx.cs:
            d = torch.empty((n, h, w, c), dtype=torch.float32, device=dout.device)
            L.check(L.get().ledb200_train_copy_channels(_p(dout), ct, off, _p(d), c, 0, n * h * w, c, _st(dout)),
                    'train_copy_channels')
            outs.append(d)
            off += c
        return tuple(outs)


def cat_channels(xs):
    """torch.cat(xs, dim=1) of the reference (channels are the last axis here)."""
    return _Cat.apply(*xs)


# ---- module-level helpers mirroring mmcv's ConvModule orders ---------------------------------------
def conv_module(x, m, relu=False, res=None):
    """ConvModule order ('conv','norm','act'): conv -> BN (+ residual) -> optional ReLU.
    `m` holds `.conv` (nn.Conv2d) and `.bn` (nn.BatchNorm2d)."""
    y = conv2d(x, m.conv.weight, m.conv.bias, m.conv.stride[0])
    return bn_act(y, m.bn, res=res, relu=relu)


def pre_conv_module(x, m):
    """ConvModule order ('norm','act','conv') (led_head.py:94, ppm.py:42-43): BN -> ReLU -> conv."""
    return conv2d(bn_act(x, m.bn, relu=True), m.conv.weight, m.conv.bias, m.conv.stride[0])
