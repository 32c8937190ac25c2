"""Pose losses: drop-ins for ``deeplio.losses`` (deeplio/losses/__init__.py:4-30, losses/losses.py:11-96).

``HWSLoss`` keeps the reference's learnable homoscedastic weights ``sx`` / ``sq`` as ``nn.Parameter``s (they join the
optimizer's second parameter group, trainer.py:56-58, and -- with ``optim.FlatAdam`` -- the flat arena and the
gradient all-reduce); ``LWSLoss`` is the beta-weighted sum.  Same constructor signatures, same ``forward`` signature
(eight tensors, the global ones already sliced by the caller), same ``loss_Types`` attribute the trainer reads
(trainer.py:249-258).  The four MSE terms, the weighting and every gradient are one launch each way (dlio_pose_loss).
"""
import torch
from torch import nn

from . import _lib as L
from . import functional as Fn
from ._lib import ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _view3(t, name):
    """The tensor as a [B, G, C] view with a contiguous last dimension (copied only if it has none)."""
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("pose loss: %s must be a float32 CUDA tensor (no CPU path)" % name)
    if t.dim() != 3:
        t = t.reshape(t.shape[0], -1, t.shape[-1])
    if t.stride(2) != 1 and t.shape[2] > 1:
        t = t.contiguous()
    return t


def _term(pred, gt, dpred):
    if pred is None:
        return L.LossTerm()
    B, G, Cc = pred.shape
    if tuple(gt.shape) != (B, G, Cc):
        raise RuntimeError("pose loss: prediction %s and ground truth %s differ in shape" % (tuple(pred.shape), tuple(gt.shape)))
    return L.LossTerm(ptr(pred), ptr(gt), ptr(dpred), B, G, Cc, pred.stride(0), pred.stride(1), gt.stride(0),
                      gt.stride(1), dpred.stride(0) if dpred is not None else 0, dpred.stride(1) if dpred is not None else 0)


class _PoseLoss(torch.autograd.Function):
    """loss(pred_t, pred_w, pred_p, pred_q | gt...) with sx, sq as differentiable device scalars (HWS) or beta (LWS)."""

    @staticmethod
    def forward(ctx, pt, pw, pp, pq, gt_t, gt_w, gt_p, gt_q, sx, sq, lws, beta, local, glob):
        preds = [_view3(t, "prediction") if (t is not None and use) else None
                 for t, use in ((pt, local), (pw, local), (pp, glob), (pq, glob))]
        gts = [_view3(t, "ground truth") if p is not None else None for t, p in zip((gt_t, gt_w, gt_p, gt_q), preds)]
        dev = next(p for p in preds if p is not None).device
        loss = torch.empty((1,), device=dev, dtype=torch.float32)
        terms = [_term(p, g, None) for p, g in zip(preds, gts)]
        L.pose_loss(*terms, ptr(sx), ptr(sq), 1 if lws else 0, float(beta), None, ptr(loss), None, None, _stream())
        ctx.save_for_backward(*[t for t in preds + gts if t is not None], *([sx, sq] if not lws else []))
        ctx.mask = [p is not None for p in preds]
        ctx.lws, ctx.beta = lws, float(beta)
        ctx.in_shapes = [t.shape if t is not None else None for t in (pt, pw, pp, pq)]
        return loss.view(())

    @staticmethod
    def backward(ctx, up):
        saved = list(ctx.saved_tensors)
        n = sum(ctx.mask)
        it_p, it_g = iter(saved[:n]), iter(saved[n:2 * n])
        preds = [next(it_p) if m else None for m in ctx.mask]
        gts = [next(it_g) if m else None for m in ctx.mask]
        sx, sq = (saved[2 * n], saved[2 * n + 1]) if not ctx.lws else (None, None)
        need = ctx.needs_input_grad
        ds = [torch.empty(p.shape, device=p.device, dtype=torch.float32) if (p is not None and need[i]) else None
              for i, p in enumerate(preds)]
        pending = []       # sx / sq living in the FlatAdam arena take their gradient in place (functional._grad_buffer)
        dsx, dsx_ret = Fn._grad_buffer(sx, pending) if (sx is not None and need[8]) else (None, None)
        dsq, dsq_ret = Fn._grad_buffer(sq, pending) if (sq is not None and need[9]) else (None, None)
        terms = [_term(p, g, d) for p, g, d in zip(preds, gts, ds)]
        up = up.contiguous()
        L.pose_loss(*terms, ptr(sx), ptr(sq), 1 if ctx.lws else 0, ctx.beta, ptr(up), None, ptr(dsx), ptr(dsq), _stream())
        Fn._flush_pending(pending)
        outs = [d.view(s) if d is not None else None for d, s in zip(ds, ctx.in_shapes)]
        return (*outs, None, None, None, None, dsx_ret, dsq_ret, None, None, None, None)


class HWSLoss(nn.Module):
    """Homoscedastic weighted loss (losses.py:51-96): (L_p + L_t) e^-sx + sx + (L_q + L_w) e^-sq + sq."""

    def __init__(self, sx=0.0, sq=-2.5, learn_hyper_params=True, device="cuda", loss_Types=(True, True)):
        super().__init__()
        self.learn_hyper_params = learn_hyper_params
        self.loss_Types = list(loss_Types)
        self.sx = nn.Parameter(torch.tensor(float(sx), device=device), requires_grad=learn_hyper_params)
        self.sq = nn.Parameter(torch.tensor(float(sq), device=device), requires_grad=learn_hyper_params)

    def forward(self, pred_f2f_x, pred_f2f_r, pred_f2g_x, pred_f2g_r, gt_f2f_x, gt_f2f_r, gt_f2g_x, gt_f2g_q):
        return _PoseLoss.apply(pred_f2f_x, pred_f2f_r, pred_f2g_x, pred_f2g_r, gt_f2f_x, gt_f2f_r, gt_f2g_x, gt_f2g_q,
                               self.sx, self.sq, False, 0.0, bool(self.loss_Types[0]), bool(self.loss_Types[1]))

    def __repr__(self):
        return _describe(self.loss_Types)


class LWSLoss(nn.Module):
    """Linear weighted sum loss (losses.py:11-49): (L_p + L_t) + beta (L_q + L_w)."""

    def __init__(self, beta=1125.0, gamma=1.0, loss_Types=(True, True)):
        super().__init__()
        self.beta = beta
        self.gamma = gamma
        self.loss_Types = list(loss_Types)

    def forward(self, pred_f2f_x, pred_f2f_r, pred_f2g_x, pred_f2g_r, gt_f2f_x, gt_f2f_r, gt_f2g_x, gt_f2g_q):
        return _PoseLoss.apply(pred_f2f_x, pred_f2f_r, pred_f2g_x, pred_f2g_r, gt_f2f_x, gt_f2f_r, gt_f2g_x, gt_f2g_q,
                               None, None, True, float(self.beta), bool(self.loss_Types[0]), bool(self.loss_Types[1]))

    def __repr__(self):
        return _describe(self.loss_Types)


def _describe(loss_types):
    if loss_types[0] and loss_types[1]:
        return "HWSLoss with f2f and f2g loss."
    if loss_types[0]:
        return "HWSLoss with only f2f loss."
    if loss_types[1]:
        return "HWSLoss with only f2g loss."
    return "Wrong loss combination!"


def get_loss_function(cfg, device):
    """The reference factory (deeplio/losses/__init__.py:4-30): same config keys, defaults and errors."""
    loss_cfg = cfg["losses"]
    loss_name = loss_cfg["active"].lower()
    params = loss_cfg.get(loss_name, {}).get("params", {})
    loss_type = loss_cfg["loss-type"].lower()
    if "+" in loss_type:
        loss_types = [True, True]
    elif loss_type == "global":
        loss_types = [False, True]
    elif loss_type == "local":
        loss_types = [True, False]
    else:
        raise ValueError("Wrong loss type selected!")
    if loss_name == "hwsloss":
        return HWSLoss(sx=params.get("sx", 0.0), sq=params.get("sq", -2.5), learn_hyper_params=params.get("learn", False),
                       device=device, loss_Types=loss_types)
    if loss_name == "lwsloss":
        return LWSLoss(beta=params.get("beta", 1125.0), loss_Types=loss_types)
    raise ValueError("Loss {} is not supported!".format(loss_name))
