"""torch.autograd.Function shims over the small-tensor C-ABI ops (dense layers, RNNs, dropout, gating).

torch supplies memory, streams and the autograd graph between these calls; every arithmetic op is a
kernel from libdeeplio_b200.so.  Inputs must be fp32 CUDA tensors.
"""
import torch

from . import _lib as L
from ._lib import ptr

_ACTS = {None: L.ACT_NONE, "none": L.ACT_NONE, "relu": L.ACT_RELU, "leaky_relu": L.ACT_LEAKY,
         "sigmoid": L.ACT_SIGMOID, "tanh": L.ACT_TANH}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("deeplio_b200: %s must be a float32 CUDA tensor (got %s on %s); there is no CPU path"
                           % (name, t.dtype, t.device))


def _grad_buffer(p, pending):
    """Where a backward kernel writes the gradient of input ``p``: returns (buffer, value handed to autograd).

    Parameters owned by ``optim.FlatAdam`` carry ``_dlio_grad_inplace``: their ``.grad`` is a zeroed slice of the flat
    gradient arena, so the kernel overwrites it directly and autograd gets None (no AccumulateGrad add kernel, no
    temporary).  A slice that already holds a gradient (``_dlio_grad_dirty``: the parameter is used more than once in
    a backward pass -- the IMU RNN runs once per window -- or gradients are being accumulated over several backward
    passes) receives a temporary instead, added by ``_flush_pending``.  Anything else: a fresh tensor for autograd."""
    g = p.grad
    if (getattr(p, "_dlio_grad_inplace", False) and g is not None and g.is_contiguous() and g.shape == p.shape
            and g.dtype == torch.float32):
        if not getattr(p, "_dlio_grad_dirty", False):
            p._dlio_grad_dirty = True
            return g, None
        tmp = torch.empty_like(g)
        pending.append((g, tmp))
        return tmp, None
    buf = torch.empty(p.shape, device=p.device, dtype=torch.float32)
    return buf, buf


def _flush_pending(pending):
    for dst, tmp in pending:
        L.axpby(ptr(dst), 1.0, ptr(tmp), 1.0, ptr(dst), dst.numel(), _stream())


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        _check(x, "input")
        _check(w, "weight")
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        m, k = x2.shape
        n = w.shape[0]
        if w.shape[1] != k:
            raise RuntimeError("linear: mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)" % (m, k, w.shape[1], n))
        y = torch.empty((m, n), device=x.device, dtype=torch.float32)
        L.linear_fwd(ptr(x2), x2.stride(0), ptr(w), ptr(b), m, n, k, act, ptr(y), n, _stream())
        ctx.save_for_backward(x2, w, y)
        ctx.act, ctx.has_b, ctx.in_shape = act, b is not None, x.shape
        ctx.bias = b if (b is not None and b.requires_grad) else None     # the bias is only needed for its .grad slot
        return y.view(*x.shape[:-1], n)

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        m, k = x2.shape
        n = w.shape[0]
        dy2 = dy.reshape(m, n)
        if dy2.stride(-1) != 1 or dy2.stride(0) < n:
            dy2 = dy2.contiguous()
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_b and ctx.needs_input_grad[2]
        dx = torch.empty((m, k), device=dy.device, dtype=torch.float32) if need_x else None
        pending = []
        dw, dw_ret = _grad_buffer(w, pending) if need_w else (None, None)
        db, db_ret = _grad_buffer(ctx.bias, pending) if need_b else (None, None)
        scr = torch.empty((m, n), device=dy.device, dtype=torch.float32) if ctx.act != L.ACT_NONE else None
        L.linear_bwd(ptr(x2), x2.stride(0), ptr(w), ptr(y), n, ptr(dy2), dy2.stride(0), m, n, k, ctx.act, ptr(dx), k,
                     ptr(dw), ptr(db), ptr(scr), _stream())
        _flush_pending(pending)
        return (dx.view(ctx.in_shape) if need_x else None), dw_ret, db_ret, None


def linear(x, w, b=None, act=None):
    """act(x @ w.T + b) -- nn.Linear + activation (aten::linear at lidar_feat_nets.py:97,144,185,233 etc.)."""
    return _Linear.apply(x, w, b, _ACTS[act])


class _Mul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        L.mul(ptr(a), ptr(b), ptr(out), a.numel(), _stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, d):
        a, b = ctx.saved_tensors
        d = d.contiguous()
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(a)
            L.mul(ptr(d), ptr(b), ptr(da), a.numel(), _stream())
        if ctx.needs_input_grad[1]:
            db = torch.empty_like(b)
            L.mul(ptr(d), ptr(a), ptr(db), a.numel(), _stream())
        return da, db


def mul(a, b):
    """Element-wise product of two same-shape tensors (soft-fusion gating, fusion_nets.py:72-73)."""
    assert a.shape == b.shape
    return _Mul.apply(a, b)


class _CatLast(torch.autograd.Function):
    """torch.cat(tensors, dim=-1) for tensors that agree on the leading dimensions: one strided copy per input."""

    @staticmethod
    def forward(ctx, *xs):
        xs = [x.contiguous() for x in xs]
        widths = [x.shape[-1] for x in xs]
        rows = xs[0].numel() // widths[0]
        out = torch.empty(xs[0].shape[:-1] + (sum(widths),), device=xs[0].device, dtype=torch.float32)
        off = 0
        for x, w in zip(xs, widths):
            L.copy2d(ptr(x), w, out.data_ptr() + 4 * off, sum(widths), rows, w, _stream())
            off += w
        ctx.widths, ctx.rows = widths, rows
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        total, off, grads = sum(ctx.widths), 0, []
        for i, w in enumerate(ctx.widths):
            g = None
            if ctx.needs_input_grad[i]:
                g = torch.empty(d.shape[:-1] + (w,), device=d.device, dtype=torch.float32)
                L.copy2d(d.data_ptr() + 4 * off, total, ptr(g), w, ctx.rows, w, _stream())
            grads.append(g)
            off += w
        return tuple(grads)


def cat_last(*xs):
    """Concatenation along the last axis (fusion_nets.py:27,70,75) through the library's strided copy."""
    for x in xs:
        _check(x, "cat_last input")
        assert x.shape[:-1] == xs[0].shape[:-1]
    return _CatLast.apply(*xs)


class _StackMid(torch.autograd.Function):
    """torch.stack([B, C] tensors, dim=1) -> [B, S, C]; the inputs may be row-strided views (``x[:, -1, :C]``)."""

    @staticmethod
    def forward(ctx, *xs):
        b, c = xs[0].shape
        s = len(xs)
        out = torch.empty((b, s, c), device=xs[0].device, dtype=torch.float32)
        for i, x in enumerate(xs):
            if x.stride(1) != 1:
                x = x.contiguous()
            L.copy2d(ptr(x), x.stride(0), out.data_ptr() + 4 * i * c, s * c, b, c, _stream())
        ctx.shape = (b, s, c)
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        b, s, c = ctx.shape
        grads = []
        for i in range(s):
            g = None
            if ctx.needs_input_grad[i]:
                g = torch.empty((b, c), device=d.device, dtype=torch.float32)
                L.copy2d(d.data_ptr() + 4 * i * c, s * c, ptr(g), c, b, c, _stream())
            grads.append(g)
        return tuple(grads)


def stack_mid(xs):
    """torch.stack(xs, dim=1) of [B, C] tensors (the per-window IMU features, imu_feat_nets.py:81-83)."""
    for x in xs:
        _check(x, "stack_mid input")
        assert x.dim() == 2 and x.shape == xs[0].shape
    return _StackMid.apply(*xs)


_drop_counter = [0]
# device int64 counter mixed into every dropout seed at run time (None: not used).  deeplio_b200.graph sets it while
# it captures a train step and advances it inside the graph, so that replays draw fresh masks.
_seed_epoch = [None]


def dropout_mask(shape, p, device):
    """Pre-scaled Bernoulli keep mask from the library's counter-based generator.  The stream is seeded by
    torch.initial_seed() and a per-process call counter, so torch.manual_seed() makes runs repeatable."""
    mask = torch.empty(shape, device=device, dtype=torch.float32)
    _drop_counter[0] += 1
    seed = (torch.initial_seed() * 0x9E3779B1 + _drop_counter[0] * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF
    L.dropout_mask(ptr(mask), mask.numel(), float(p), seed, ptr(_seed_epoch[0]), _stream())
    return mask


def dropout(x, p, training):
    """nn.Dropout: identity in eval mode or for p == 0."""
    if not training or p <= 0.0:
        return x
    return mul(x, dropout_mask(x.shape, p, x.device))


class _SumMid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        a, t, c = x.shape
        out = torch.empty((a, c), device=x.device, dtype=torch.float32)
        L.sum_mid(ptr(x), ptr(out), a, t, c, _stream())
        ctx.t = t
        return out

    @staticmethod
    def backward(ctx, d):
        return d.unsqueeze(1).expand(-1, ctx.t, -1)


def sum_mid(x):
    """[A, T, C] -> [A, C], sum over T (ImuFeatFC time sum, imu_feat_nets.py:50)."""
    return _SumMid.apply(x)


class _Axpby(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, alpha, beta):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        L.axpby(ptr(a), alpha, ptr(b), beta, ptr(out), a.numel(), _stream())
        ctx.alpha, ctx.beta = alpha, beta
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()

        def scaled(s):
            if s == 1.0:
                return d
            o = torch.empty_like(d)
            L.axpby(ptr(d), s, ptr(d), 0.0, ptr(o), d.numel(), _stream())
            return o
        return scaled(ctx.alpha), scaled(ctx.beta), None, None


def axpby(a, b, alpha=1.0, beta=1.0):
    return _Axpby.apply(a, b, float(alpha), float(beta))


class _Rnn(torch.autograd.Function):
    """Multi-layer (bi)LSTM / GRU, batch_first (aten::lstm / aten::gru at imu_feat_nets.py:79-82,
    odom_feat_nets.py:80).  Returns (out [B,T,D*H], h_n, c_n)."""

    @staticmethod
    def forward(ctx, x, h0, c0, kind, L_, D, H, drop_mask, *weights):
        _check(x, "rnn input")
        x = x.contiguous()
        B, T, I = x.shape
        dev = x.device
        for w in weights:
            _check(w, "rnn weight")
        ws = [w if w.is_contiguous() else w.contiguous() for w in weights]
        out = torch.empty((B, T, D * H), device=dev, dtype=torch.float32)
        hn = torch.empty((L_ * D, B, H), device=dev, dtype=torch.float32)
        cn = torch.empty((L_ * D, B, H), device=dev, dtype=torch.float32)
        reserve = torch.empty((L.rnn_reserve_floats(kind, L_, D, B, T, I, H),), device=dev, dtype=torch.float32)
        h0c = h0.contiguous() if h0 is not None else None
        c0c = c0.contiguous() if c0 is not None else None
        L.rnn_fwd(kind, L_, D, B, T, I, H, L.ptr_array(ws), ptr(x), ptr(h0c), ptr(c0c), ptr(drop_mask), ptr(out),
                  ptr(hn), ptr(cn), ptr(reserve), _stream())
        ctx.save_for_backward(x, reserve, drop_mask, *ws)
        ctx.weight_objs = weights      # the parameters themselves (their .grad slots), not contiguous copies
        ctx.dims = (kind, L_, D, B, T, I, H)
        ctx.has_state = (h0 is not None, c0 is not None)
        ctx.mark_non_differentiable()
        return out, hn, cn

    @staticmethod
    def backward(ctx, dout, dhn, dcn):
        x, reserve, drop_mask, *ws = ctx.saved_tensors
        kind, L_, D, B, T, I, H = ctx.dims
        dev = x.device
        pending = []
        bufs = [_grad_buffer(w, pending) for w in ctx.weight_objs]
        grads = [b for b, _ in bufs]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dh0 = torch.empty((L_ * D, B, H), device=dev, dtype=torch.float32)
        dc0 = torch.empty((L_ * D, B, H), device=dev, dtype=torch.float32)
        nscr = L.rnn_bwd_scratch_floats(kind, L_, D, B, T, I, H)
        scratch = torch.empty((nscr,), device=dev, dtype=torch.float32)
        dout = dout.contiguous() if dout is not None else None
        dhn = dhn.contiguous() if dhn is not None else None
        dcn = dcn.contiguous() if (dcn is not None and kind == 0) else None
        L.rnn_bwd(kind, L_, D, B, T, I, H, L.ptr_array(ws), ptr(x), ptr(drop_mask), ptr(dout), ptr(dhn), ptr(dcn),
                  ptr(reserve), L.ptr_array(grads), ptr(dx), ptr(dh0), ptr(dc0), ptr(scratch), nscr, _stream())
        _flush_pending(pending)
        rets = [r if need else None for (_, r), need in zip(bufs, ctx.needs_input_grad[8:])]
        return (dx, dh0 if ctx.has_state[0] else None, dc0 if ctx.has_state[1] else None,
                None, None, None, None, None, *rets)


def rnn(x, state, kind, num_layers, bidirectional, hidden_size, weights, dropout_p=0.0, training=False):
    """kind 'lstm' | 'gru'; state None | (h0, c0) | h0; weights in torch's flat order
    (w_ih, w_hh, b_ih, b_hh per layer and direction).  Returns (out, new_state) like nn.LSTM / nn.GRU."""
    k = 0 if kind == "lstm" else 1
    D = 2 if bidirectional else 1
    h0 = c0 = None
    if state is not None:
        h0, c0 = state if k == 0 else (state, None)
    mask = None
    if training and dropout_p > 0.0 and num_layers > 1:
        B, T, _ = x.shape
        mask = dropout_mask((num_layers - 1, B, T, D * hidden_size), dropout_p, x.device)
    out, hn, cn = _Rnn.apply(x, h0, c0, k, num_layers, D, hidden_size, mask, *weights)
    return out, ((hn, cn) if k == 0 else hn)


class _HwsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, ori, gt_pos, gt_ori, sx, sq):
        pos, ori, gt_pos, gt_ori = (t.contiguous() for t in (pos, ori, gt_pos, gt_ori))
        loss = torch.empty((1,), device=pos.device, dtype=torch.float32)
        dpos, dori = torch.empty_like(pos), torch.empty_like(ori)
        L.hws_loss(ptr(pos), ptr(ori), ptr(gt_pos), ptr(gt_ori), pos.numel(), sx, sq, ptr(loss), ptr(dpos), ptr(dori),
                   _stream())
        ctx.save_for_backward(dpos, dori)
        return loss.view(())

    @staticmethod
    def backward(ctx, d):
        dpos, dori = ctx.saved_tensors
        return mul(dpos, d.expand_as(dpos).contiguous()), mul(dori, d.expand_as(dori).contiguous()), None, None, None, None


def hws_loss(pos, ori, gt_pos, gt_ori, sx=0.0, sq=-3.0):
    """Frame-to-frame HWSLoss with fixed weights (losses.py:68-86): one launch for the loss and its gradient."""
    return _HwsLoss.apply(pos, ori, gt_pos, gt_ori, float(sx), float(sq))
