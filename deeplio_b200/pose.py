"""Pose chaining and the train loop's sanity guards on the device.

``se3_to_SE3`` replaces ``Trainer.se3_to_SE3`` (deeplio/models/trainer.py:324-351): a Python loop over batch x pairs
with ``SO3.exp``, two 3x3 products, ``SO3.from_matrix(normalize=True).to_quaternion()`` and two ``torch.det`` host
synchronisations per pair becomes one forward and one backward launch (dlio_se3_chain_fwd / _bwd).  ``check_finite``
replaces the six ``isnan().any() or isinf().any()`` pairs of trainer.py:221-229,240-243 (each a reduction and a host
synchronisation) with one pass over all tensors and one flag word.

The reference raises ``ValueError`` from inside its loops; kernels cannot, so they OR flags into a per-device
status word.  ``raise_for_status()`` reads it (one host synchronisation) and raises the reference's errors;
``se3_to_SE3(..., check=True)`` does that immediately (drop-in behaviour), ``check=False`` leaves it to the caller
(e.g. once every N steps, or through ``pipeline.LaggedScalar``).
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import ptr

ST_NONFINITE, ST_DET_STEP, ST_DET_CHAIN, ST_INVALID_GT = 1, 2, 4, 8
_status = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def status_word(device):
    """The per-device int32 status word the pose / ground-truth kernels OR their flags into."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("deeplio_b200.pose runs on CUDA devices only (no CPU fallback)")
    key = device.index if device.index is not None else torch.cuda.current_device()
    t = _status.get(key)
    if t is None:
        t = _status[key] = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", key))
    return t


def raise_for_status(device, clear=True):
    """Synchronises, reads the status word and raises what the reference raises (trainer.py:341-348, misc.py:106)."""
    t = status_word(device)
    st = int(t.item())
    if clear and st:
        t.zero_()
    if st & ST_NONFINITE:
        raise ValueError("pose: NaN / Inf in the frame-to-frame predictions or the ground truth")
    if st & ST_DET_STEP:
        raise ValueError("Det error: det(exp(w)) is not close to 1")
    if st & ST_DET_CHAIN:
        raise ValueError("Det error: det(R) of the accumulated rotation is not close to 1")
    if st & ST_INVALID_GT:
        raise ValueError("Invalid rotation matrix. Use normalize=True to handle rounding errors.")
    return st


class _Se3Chain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        if not (x.is_cuda and w.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32):
            raise RuntimeError("se3_to_SE3: float32 CUDA tensors expected (no CPU path)")
        x, w = x.contiguous(), w.contiguous()
        B, S, _ = x.shape
        ox = torch.empty((B, S, 3), device=x.device, dtype=torch.float32)
        oq = torch.empty((B, S, 4), device=x.device, dtype=torch.float32)
        L.se3_chain_fwd(ptr(x), ptr(w), B, S, ptr(ox), ptr(oq), ptr(status_word(x.device)), _stream())
        ctx.save_for_backward(x, w)
        return ox, oq

    @staticmethod
    def backward(ctx, gx, gq):
        x, w = ctx.saved_tensors
        B, S, _ = x.shape
        gx = gx.contiguous() if gx is not None else None
        gq = gq.contiguous() if gq is not None else None
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        L.se3_chain_bwd(ptr(x), ptr(w), B, S, ptr(gx), ptr(gq), ptr(dx), ptr(dw), _stream())
        return dx, dw


def se3_to_SE3(f2f_x, f2f_r, check=True):
    """[B,S,3] translations and so(3) rotations of consecutive pairs -> (f2g_x [B,S,3], f2g_q [B,S,4] wxyz), the
    poses relative to the first frame; differentiable."""
    ox, oq = _Se3Chain.apply(f2f_x, f2f_r)
    if check:
        raise_for_status(f2f_x.device)
    return ox, oq


def check_finite(named_tensors, check=True):
    """One fused NaN / Inf scan over up to 8 tensors: ``named_tensors`` is a list of (name, tensor).  Returns the
    flag tensor (bit i <=> tensor i); with ``check`` raises ValueError naming the first offending tensor, as the
    reference's guards do (trainer.py:221-229,240-243)."""
    items = [(n, t.contiguous()) for n, t in named_tensors if torch.is_tensor(t) and t.numel() > 0]
    if not items:
        return None
    if len(items) > 8:
        raise ValueError("check_finite: at most 8 tensors per call")
    dev = items[0][1].device
    for n, t in items:
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError("check_finite: %s must be a float32 CUDA tensor" % n)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    ptrs = (C.c_void_p * len(items))(*[t.data_ptr() for _, t in items])
    sizes = (C.c_longlong * len(items))(*[t.numel() for _, t in items])
    L.finite_check(ptrs, sizes, len(items), ptr(flags), _stream())
    if check:
        f = int(flags.item())
        for i, (n, _) in enumerate(items):
            if f & (1 << i):
                raise ValueError("%s: NaN / Inf" % n)
    return flags
