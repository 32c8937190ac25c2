"""Flat-arena Adam: the optimizer side of the train step (reference: deeplio/models/optimizer.py:4-16,
``torch.optim.Adam(params, lr, weight_decay)`` -- L2 decay added to the gradient).

All parameters of the model are re-homed into ONE contiguous fp32 arena (each parameter 16-byte aligned), with
a matching gradient arena whose slices are installed as ``param.grad``.  One kernel launch then updates the
whole model (dlio_adam_step), one memset clears all gradients, and one NCCL all-reduce on the gradient arena
is the data-parallel exchange (deeplio_b200.parallel).
"""
import torch

from . import _lib as L
from ._lib import ptr


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam: parameters must live on a CUDA device (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        offsets, total = [], 0
        for p in self.params:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, offsets):
                view = self.flat_param[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
        self.offsets = offsets
        self.step_count = 0
        # zero_grad() below leaves every .grad a zeroed arena slice, so the encoders' backward kernels may write
        # their parameter gradients straight into the arena instead of going through AccumulateGrad (engine.py);
        # this assumes ONE backward per zero_grad(), which is how the reference trains (trainer.py:268-272)
        for p in self.params:
            p._dlio_grad_inplace = True

    def zero_grad(self):
        """One memset for every gradient; the views stay installed so autograd accumulates in place."""
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)

    def step(self, grad_scale=1.0):
        self.step_count += 1
        L.adam_step(ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.numel,
                    self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                    float(grad_scale), torch.cuda.current_stream().cuda_stream)


def create_optimizer(params, cfg, args):
    """Mirror of the reference factory (optimizer.py:4-16) for its default, ``optimizer: adam``."""
    name = cfg.get("optimizer", "adam").lower()
    if name != "adam":
        raise ValueError("deeplio_b200 implements the reference's default optimizer (adam); got %r" % name)
    return FlatAdam(params, lr=args.lr, weight_decay=args.weight_decay)
