"""Flat-arena Adam: the optimizer side of the train step (reference: deeplio/models/optimizer.py:4-16,
``torch.optim.Adam(params, lr, weight_decay)`` -- L2 decay added to the gradient).

All parameters are re-homed into ONE contiguous fp32 arena (each parameter 16-byte aligned), with a matching
gradient arena whose slices are installed as ``param.grad``.  One kernel launch per parameter group then updates
the model (dlio_adam_step), one memset clears all gradients, and one NCCL all-reduce on the gradient arena is the
data-parallel exchange (deeplio_b200.parallel).

The object follows the ``torch.optim.Optimizer`` protocol as far as the reference's Trainer uses it
(trainer.py:56-58,108,161,281; misc.py:131-165 ``PolynomialLRDecay``): it accepts an iterable of parameters or of
param-group dicts, exposes live ``param_groups`` (``step()`` reads ``lr`` / ``betas`` / ``eps`` / ``weight_decay``
from them), ``state_dict()`` / ``load_state_dict()`` in torch.optim.Adam's layout (a checkpoint written by either
loads into the other), ``zero_grad()`` and ``step()``.  Parameters whose ``.grad`` is None at ``step()`` are
skipped like torch.optim.Adam skips them (no decay, no moment update, their step count does not advance).
"""
import torch

from . import _lib as L
from ._lib import ptr

_DEFAULTS = {"lr": 1e-3, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0.0}


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(params)
        if not params:
            raise ValueError("FlatAdam: optimizer got an empty parameter list")
        groups = params if isinstance(params[0], dict) else [{"params": params}]
        defaults = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay}
        self.defaults = dict(defaults)
        self.param_groups = []
        self.params = []
        seen = set()
        for g in groups:
            g = dict(g)
            ps = g["params"]
            ps = [ps] if torch.is_tensor(ps) else list(ps)
            ps = [p for p in ps if p.requires_grad]
            for p in ps:
                if id(p) in seen:
                    raise ValueError("FlatAdam: some parameters appear in more than one parameter group")
                seen.add(id(p))
            for k, v in defaults.items():
                g.setdefault(k, v)
            g["params"] = ps
            g["_first"] = len(self.params)       # index of the group's first parameter in self.params
            self.params.extend(ps)
            self.param_groups.append(g)
        if not self.params:
            raise ValueError("FlatAdam: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam: parameters must live on a CUDA device (no CPU fallback)")
        offsets, total = [], 0
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("FlatAdam: all parameters must be float32 on one CUDA device")
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
        self.steps = [0] * len(self.params)      # per-parameter step counts (they differ only after skipped steps)
        # zero_grad() leaves every .grad a zeroed arena slice, so the encoders' backward kernels may write their
        # parameter gradients straight into the arena instead of going through AccumulateGrad (engine.Run.param_grad).
        # ``_dlio_grad_dirty`` marks a slice that already holds a gradient: a second backward before the next
        # zero_grad() (gradient accumulation) then adds instead of overwriting.
        for p in self.params:
            p._dlio_grad_inplace = True
            p._dlio_grad_dirty = False

    # ------------------------------------------------------------------ torch.optim protocol
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        for g in self.param_groups:
            g["lr"] = v

    @property
    def step_count(self):
        return max(self.steps)

    def _grad_view(self, i):
        p, off = self.params[i], self.offsets[i]
        return self.flat_grad[off:off + p.numel()].view_as(p)

    def zero_grad(self, set_to_none=False):
        """One memset for every gradient; the views stay installed so autograd accumulates in place
        (``set_to_none`` is accepted for signature compatibility and ignored: the arena is the gradient storage)."""
        self.flat_grad.zero_()
        base = self.flat_grad.data_ptr()
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                p.grad = self._grad_view(i)
            p._dlio_grad_dirty = False

    def _rehome(self):
        """Make the arena hold every parameter's gradient.  A ``.grad`` that no longer aliases its arena slice
        (``model.zero_grad(set_to_none=True)`` followed by a backward, or user code assigning ``.grad``) is copied in
        and the view re-installed; returns the indices of parameters without a gradient (skipped by this step)."""
        base = self.flat_grad.data_ptr()
        skipped = []
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            g = p.grad
            if g is None:
                skipped.append(i)
            elif g.data_ptr() != base + 4 * off:
                view = self._grad_view(i)
                view.copy_(g)
                p.grad = view
        return skipped

    def step(self, grad_scale=1.0, closure=None):
        loss = closure() if closure is not None else None
        skipped = set(self._rehome())
        st = torch.cuda.current_stream().cuda_stream
        ends = self.offsets[1:] + [self.numel]
        for gi, g in enumerate(self.param_groups):
            first = g["_first"]
            last = first + len(g["params"])
            # contiguous runs of parameters that take this step and share one step count: one launch each
            i = first
            while i < last:
                if i in skipped:
                    i += 1
                    continue
                j = i
                while j + 1 < last and (j + 1) not in skipped and self.steps[j + 1] == self.steps[i]:
                    j += 1
                a, b = self.offsets[i], ends[j]
                step = self.steps[i] + 1
                L.adam_step(ptr(self.flat_param) + 4 * a, ptr(self.flat_grad) + 4 * a, ptr(self.exp_avg) + 4 * a,
                            ptr(self.exp_avg_sq) + 4 * a, b - a, float(g["lr"]), float(g["betas"][0]),
                            float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), step,
                            float(grad_scale), st)
                for k in range(i, j + 1):
                    self.steps[k] = step
                i = j + 1
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout: ``state[i] = {step, exp_avg, exp_avg_sq}`` per parameter index and
        ``param_groups`` with index lists (what trainer.py:161 saves)."""
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            if self.steps[i] == 0:
                continue
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.steps[i])),
                        "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        groups = []
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k not in ("params", "_first")}
            d.setdefault("amsgrad", False)
            d["params"] = list(range(g["_first"], g["_first"] + len(g["params"])))
            groups.append(d)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(
                len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("FlatAdam: loaded state dict has different parameter groups")
        if any(g.get("amsgrad", False) for g in groups):
            raise ValueError("FlatAdam: amsgrad state cannot be loaded (the fused kernel implements plain Adam)")
        for src, dst in zip(groups, self.param_groups):
            for k, v in src.items():
                if k != "params":
                    dst[k] = tuple(v) if k == "betas" else v
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.steps = [0] * len(self.params)
        order = [i for g in groups for i in g["params"]]       # saved index of our k-th parameter
        with torch.no_grad():
            for k, saved in enumerate(order):
                st = sd["state"].get(saved)
                if st is None:
                    continue
                p, off = self.params[k], self.offsets[k]
                n = p.numel()
                self.steps[k] = int(st["step"])
                self.exp_avg[off:off + n].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[off:off + n].view_as(p).copy_(st["exp_avg_sq"])


def create_optimizer(params, cfg, args):
    """The reference factory (optimizer.py:4-16) for its default, ``optimizer: adam``; ``params`` may be the list of
    param-group dicts trainer.py:56 builds."""
    name = cfg.get("optimizer", "adam").lower()
    if name != "adam":
        raise ValueError("deeplio_b200 implements the reference's default optimizer (adam); got %r" % name)
    return FlatAdam(params, lr=args.lr, weight_decay=args.weight_decay)
