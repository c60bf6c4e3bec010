"""Optimiser side of the training step: SGD + momentum over ONE flat parameter arena, PolyLR and the
data-parallel gradient all-reduce.

Reference: ``configs/LED_Net/LEDNet_80k_cityscapes-1024x1024.py:64-75`` (SGD lr 0.01, momentum 0.9,
weight_decay 5e-4; PolyLR power 0.9, eta_min 0, by_epoch=False; the `schedule_80k` base it overrides uses 1e-4) driven by mmengine's OptimWrapper
and MMDistributedDataParallel (``tools/train.py``).  Here all parameters (and all gradients) live in
one contiguous fp32 buffer each, so the optimiser step is ONE kernel launch (csrc/train.cu sgd_kernel)
and the DDP exchange is ONE NCCL all-reduce of the flat gradient - the only collective of the
training step.
"""
import ctypes as C

import torch

from . import lib as L


class FlatSGD:

    def __init__(self, params, lr=0.01, momentum=0.9, weight_decay=5e-4, process_group=None,
                 world_size=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('FlatSGD got no trainable parameters')
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.buf = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise L.LedB200Error('FlatSGD: parameters must be fp32')
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)            # parameters become views of the arena
            p.grad = self.flat_grad[off:off + k].view_as(p)       # autograd accumulates in place
            off += k
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.steps = 0
        self.process_group = process_group
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.world_size = world_size

    def zero_grad(self):
        self.flat_grad.zero_()
        off = 0
        for p in self.params:                      # re-attach if something replaced .grad
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + k].view_as(p)
            off += k

    def all_reduce_grads(self):
        """DDP gradient exchange: one sum all-reduce of the flat buffer (averaging is folded into the
        SGD kernel's grad_scale)."""
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group)

    def _gather_stray_grads(self):
        """`model.zero_grad()` (set_to_none) between zero_grad() and backward() makes autograd allocate fresh `.grad`
        tensors: copy them into the flat buffer (and re-attach) instead of silently stepping on zeros."""
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                self.flat_grad[off:off + k].copy_(p.grad.reshape(-1))
                p.grad = self.flat_grad[off:off + k].view_as(p)
            off += k

    def step(self):
        self._gather_stray_grads()
        self.all_reduce_grads()
        if not self.flat.is_cuda:
            raise L.LedB200Error('FlatSGD.step needs CUDA parameters (no CPU fallback)')
        p = lambda t: C.c_void_p(t.data_ptr())     # noqa: E731
        L.check(L.get().ledb200_train_sgd_step(p(self.flat), p(self.flat_grad), p(self.buf), self.flat.numel(),
                                               float(self.lr), float(self.momentum), float(self.weight_decay),
                                               int(self.steps == 0), 1.0 / self.world_size,
                                               L.stream_ptr(self.flat.device)), 'ledb200_train_sgd_step')
        self.steps += 1


class PolyLR:
    """mmengine PolyLR (by_epoch=False): lr_t = (base - eta_min) * (1 - t/T)^power + eta_min."""

    def __init__(self, optimizer, power=0.9, eta_min=0.0, begin=0, end=80000):
        self.opt, self.power, self.eta_min, self.begin, self.end = optimizer, power, eta_min, begin, end
        self.base_lr = optimizer.lr
        self.t = 0

    def lr_at(self, t):
        t = min(max(t - self.begin, 0), self.end - self.begin)
        return (self.base_lr - self.eta_min) * (1 - t / (self.end - self.begin)) ** self.power + self.eta_min

    def step(self):
        self.t += 1
        self.opt.lr = self.lr_at(self.t)
        return self.opt.lr
