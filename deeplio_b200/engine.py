"""Host-side executor for the convolutional encoders: a small tape over the C ABI.

The LiDAR encoders do not go through torch autograd op by op.  A forward pass runs a straight-line
"program" of fused C-ABI calls (conv + bias + ReLU + BN statistics; BN-apply + residual + ReLU + max-pool;
SE; global average) on padded-NHWC buffers and records one backward closure per call; the backward pass
replays the closures in reverse.  torch is used for device memory (the caching allocator) and the
current stream only.

Reference semantics implemented here: base_net.py:55-71 (conv helper), lidar_feat_nets.py:306-342
(Simple-1 blocks), pointseg_modules.py:110-141,203-221 (Fire / SE), torchvision resnet.py:88-103
(BasicBlock), BatchNorm2d train/eval statistics.
"""
import os

import torch

from . import _lib as L
from ._lib import ptr

# defaults of nn.BatchNorm2d; the values actually used come from the BatchNorm2d module that owns the layer's
# parameters (Run.bn_cfg: PointSeg passes momentum=bn_d, pointseg_modules.py:98-106)
BN_MOMENTUM = 0.1
BN_EPS = 1e-5
# the two encoders of the siamese pair are independent until their feature vectors meet: with ENC_STREAMS they are
# enqueued on two streams, so the HBM-bound BN / pool passes of one overlap the tensor-bound convolutions of the other
ENC_STREAMS = os.environ.get("DLIO_ENC_STREAMS", "0") == "1"
# layers without pooling / residual: the gradient at the BN output is the incoming gradient (masked by the ReLU inside
# the apply pass), so backward pass 1 only takes the sums and dz never goes through HBM
SKIP_DZ = os.environ.get("DLIO_SKIP_DZ", "1") == "1"
# tcgen05 (3xTF32) convolution path: stride-1 convolutions whose input has a multiple of 32 channels
USE_TC = True
# tcgen05 fp16 split path ("3xF16", twice the TF32 rate): stride-1 convolutions, input channels a multiple of 64;
# operands are packed fp16 hi|lo planes scaled from device-resident bounds (include/deeplio_b200.h)
USE_F16 = True
# optim.FlatAdam marks its parameters with ``_dlio_grad_inplace``: their .grad is a zeroed slice of one gradient
# arena before each backward, so the backward kernels may write parameter gradients straight into it (overwrite
# == accumulate into zeros) and hand autograd None -- one AccumulateGrad add kernel less per parameter tensor
# (116 launches per step).  One backward per zero_grad(), as in the reference's train loop.
# conv -> (ReLU) -> BN -> max-pool layers: BN-backward sums from the pooled side + un-pooling inside the apply pass
# (dlio_pool_bwd_sums / dlio_bn_pool_bwd_apply) instead of materialising dz
FUSED_POOL_BWD = True
# debugging aid (scripts/repeat_parity.py): when a dict, every conv_bn backward stores clones of its tensors here
DEBUG_TRACE = None
# bench aid: when a dict, conv_bn adds the ALGORITHMIC FLOPs (2 * Cout * Ho * Wo * Cin * kh * kw per image; real channel
# counts, no padding, no split-precision factor) of every convolution it runs to the kernel class that runs it
# (the dlio_prof_kind names), so bench.py can quote a roofline for any workload
FLOPS = None
# test aid: when a dict, every conv_bn forward stores the ReLU mask of its layer (bool, NCHW) under
# Run.trace_prefix + conv name, so that tests can count mask disagreements with the oracle (oracle.TRACE)
MASK_TRACE = None
# profiling aid (DLIO_NVTX=1, the same switch as the library's dlio/<kernel class> ranges): an NVTX range per layer,
# "<encoder>.<conv name>" around its forward launches and "<...>/bwd" around its backward ones
NVTX = os.environ.get("DLIO_NVTX", "0") == "1"


def _trace_mask(run, cname, y, bnv, pre_relu, relu, res, res_mode, c_off):
    if pre_relu:
        m = y.t > 0                                       # y holds relu(conv + bias): positive <=> the ReLU passed it
    elif relu:
        # the kernels evaluate fmaf(scale, y, shift): a single rounding of the exact value, which fp64 holds exactly
        v = (y.t.double() * bnv[2].double() + bnv[3].double()).float()
        if res is not None and res_mode == 1:
            if res.t is None:
                return
            v = v + res.t[:, res.ph:res.ph + res.h, res.pw:res.pw + res.w, c_off:c_off + y.c]
        m = v > 0
    else:
        return
    MASK_TRACE[getattr(run, "trace_prefix", "") + cname] = m.permute(0, 3, 1, 2).cpu()


def tc_ok(cin, cout, stride):
    return USE_TC and tuple(stride) == (1, 1) and cin % 32 == 0 and cout % 16 == 0


def f16_ok(cin, cout, stride):
    return USE_TC and USE_F16 and tuple(stride) == (1, 1) and cin % 64 == 0 and cout % 16 == 0


# backward of narrow 3xTF32 layers (Fire squeezes) on the tensor cores through zero-padded dy channels / weight rows
NARROW_TC_BWD = os.environ.get("DLIO_NARROW_TC_BWD", "1") == "1"
# first layer on the fp16 kernels through the "folded split" operands (csrc/conv_s2d.cu); DLIO_FIRST_F16=0 keeps 3xTF32
FIRST_F16 = os.environ.get("DLIO_FIRST_F16", "1") == "1"


def first_layer_f16_ok(wshape, stride, width):
    """The first convolution (<= 8 input channels) can run on packed fp16 input planes: W stride 1 or 2 (four or two
    outputs per 4-pixel group -> 256 or 128 tensor-core columns), 64 output channels (one dy box per output pixel
    in wgrad)."""
    cout, cin, kh, kw = wshape
    return (USE_TC and USE_F16 and FIRST_F16 and tuple(stride) in ((1, 1), (1, 2)) and cout == 64 and cin <= 8
            and kw <= 7 and kw % 2 == 1 and width % 4 == 0)


def pair_ok(x, cin, cout, kh, kw, stride):
    """W-stride-2 convolution on the tensor cores through the pixel-pair view (csrc/conv_s2d.cu): the input's fp16
    planes must be in the pixel-pair layout (producer called with ``out_group=2``); H stride 2 is computed over the
    whole grid and decimated by the epilogue."""
    sh, sw = stride
    return (USE_TC and USE_F16 and sw == 2 and sh in (1, 2) and kw in (1, 3, 5) and x.h2 is not None and x.group == 2
            and x.c == cin and cin % 32 == 0 and cout % 64 == 0 and x.w % 2 == 0 and x.pw % 2 == 0
            and x.pw // 2 >= (1 if kw > 1 else 0) and x.ph >= (kh - 1) // 2 and (sh == 1 or x.h > 1))


def consumer_reads_f16_only(cin, cout, k, stride, w_out, h_out=2):
    """True when a convolution (cin -> cout, kernel k, stride) that is the ONLY consumer of a tensor of extent
    ``h_out`` x ``w_out`` runs its forward, dgrad and wgrad on the fp16 tensor-core kernels, so that the producer may
    skip the fp32 copy (``conv_bn(out_f32=False)``) -- the stride-1 path or the pixel-pair view of a W-stride-2 layer.
    Same predicate as ``f16_ok`` / ``pair_ok`` for a producer that pads for this consumer (even row pads >= 2 when
    the consumer is W-strided)."""
    kh, kw = (k, k) if isinstance(k, int) else k
    if not (USE_TC and USE_F16) or cout % 64 != 0:
        return False
    if tuple(stride) == (1, 1):
        return cin % 64 == 0
    return (stride[1] == 2 and stride[0] in (1, 2) and kw in (1, 3, 5) and cin % 32 == 0 and w_out % 2 == 0
            and (stride[0] == 1 or h_out > 1))


def stream():
    return torch.cuda.current_stream().cuda_stream


def pool_out(n, s, ceil_mode):
    """Output extent of MaxPool2d(3, s, padding=1) (torch ceil_mode rule: the last window must start
    inside the input or its left padding)."""
    if not ceil_mode:
        return (n + 2 - 3) // s + 1
    o = -(-(n + 2 - 3) // s) + 1
    if (o - 1) * s >= n + 1:
        o -= 1
    return o


class Act:
    """Padded NHWC activation: memory [n][h+2ph][w+2pw][c], zero pads.  Up to three representations:
    ``t`` fp32 (what every non-tensor-core consumer reads), ``lo`` the low-order TF32 plane x - trunc_tf32(x) that
    goes with ``t`` on the 3xTF32 path, ``h2`` the packed fp16 hi|lo planes [n][hp][wp][2][c] of the 3xF16 path.
    ``bound`` (device float[1]) >= max |x|: defines the scale of ``h2`` and feeds the bound of a residual sum."""
    __slots__ = ("n", "h", "w", "c", "ph", "pw", "t", "lo", "h2", "bound", "needs_grad", "group")

    def __init__(self, n, h, w, c, ph=0, pw=0, device=None, t=None, needs_grad=True, split=False, lo=None,
                 f32=True, f16=False):
        self.n, self.h, self.w, self.c, self.ph, self.pw = n, h, w, c, ph, pw
        shape = (n, h + 2 * ph, w + 2 * pw, c)
        self.t = t if t is not None else (torch.empty(shape, device=device, dtype=torch.float32) if f32 else None)
        self.lo = lo if lo is not None else (torch.empty_like(self.t) if split else None)
        self.h2 = torch.empty(shape[:3] + (2, c), device=device, dtype=torch.float16) if f16 else None
        self.bound = None
        self.needs_grad = needs_grad
        self.group = 1            # layout of h2: 1 plain, 2 pixel pairs (input of a W-stride-2 tensor-core convolution)

    @property
    def t4(self):
        return L.Tensor4(self.n, self.h, self.w, self.c, self.ph, self.pw)

    @property
    def t4_unpadded(self):
        return L.Tensor4(self.n, self.h, self.w, self.c, 0, 0)


class Run:
    """One forward pass (and, if ``record``, its backward tape)."""

    def __init__(self, params, buffers, device, training, record):
        self.params = params      # name -> tensor (OIHW conv weights, BN affine, SE linears)
        self.buffers = buffers    # name -> tensor (BN running statistics)
        self.device = device
        self.training = training
        self.record = record
        self.tape = []
        self.agrad = {}           # id(Act) -> unpadded NHWC gradient tensor
        self.fgrad = {}           # id(feature buffer) -> gradient tensor [N, ld]
        self.pgrad = {}           # parameter name -> gradient tensor (absent: written in place into param.grad)
        self.keep = []            # keeps Acts alive so that id() stays unique
        self.param_objs = {}      # name -> nn.Parameter (in-place parameter gradients)
        self.bn_cfg = {}          # BatchNorm2d name -> (momentum or None, eps)
        self.accum = []           # (param.grad, temporary): gradients to ADD after the tape (second backward)
        self._zblock, self._zoff = None, 0   # current block of the small-zeros arena

    # -- helpers
    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, device=self.device, dtype=dtype)

    ZBLOCK = 1 << 18

    def zeros(self, *shape, dtype=torch.float32):
        """Zeroed tensor.  Small ones (the per-layer fp64 statistics / sums: ~6 per convolution) are carved out of
        256 KB blocks zeroed with one fill each instead of one fill kernel per tensor."""
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if nbytes > 16384:
            return torch.zeros(shape, device=self.device, dtype=dtype)
        if self._zblock is None or self._zoff + nbytes > self.ZBLOCK:
            self._zblock, self._zoff = torch.zeros(self.ZBLOCK, device=self.device, dtype=torch.uint8), 0
        t = self._zblock[self._zoff:self._zoff + nbytes].view(dtype).view(shape)
        self._zoff = (self._zoff + nbytes + 15) & ~15
        return t

    def param_grad(self, name, like):
        """Tensor the backward kernels write the gradient of parameter ``name`` into: the parameter's own .grad
        (FlatAdam parameters; nothing is returned to autograd for it) or a fresh tensor registered in ``pgrad``."""
        p = self.param_objs.get(name)
        if (p is not None and getattr(p, "_dlio_grad_inplace", False) and p.grad is not None and p.grad.is_contiguous()
                and p.grad.shape == like.shape and p.grad.dtype == torch.float32):
            if not getattr(p, "_dlio_grad_dirty", False):
                p._dlio_grad_dirty = True      # cleared by FlatAdam.zero_grad()
                return p.grad
            # the slice already holds a gradient (a second backward before zero_grad(): gradient accumulation, as
            # AccumulateGrad would do): the kernels write a temporary that is added after the tape
            g = torch.empty_like(like)
            self.accum.append((p.grad, g))
            return g
        g = torch.empty_like(like)
        self.pgrad[name] = g
        return g

    def grad_slot(self, act, zeroed=False):
        """Returns (buffer, existed).  A fresh buffer is uninitialised unless ``zeroed``."""
        g = self.agrad.get(id(act))
        if g is not None:
            return g, True
        make = self.zeros if zeroed else self.empty
        g = make(act.n, act.h, act.w, act.c)
        self.agrad[id(act)] = g
        return g, False

    def add_grad(self, act, write, accumulates=False):
        """``write(buf)`` overwrites ``buf`` with a full gradient contribution for ``act``.  ``accumulates``: the
        writer takes a second argument, ``write(buf, 1)`` ADDS its contribution to ``buf`` (the dgrad kernels'
        epilogue), so a second contribution needs neither a temporary nor an add pass."""
        g, existed = self.grad_slot(act)
        if not existed:
            write(g)
        elif accumulates:
            write(g, 1)
        else:
            tmp = torch.empty_like(g)
            write(tmp)
            L.axpby(ptr(g), 1.0, ptr(tmp), 1.0, ptr(g), g.numel(), stream())

    def backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []
        for dst, tmp in self.accum:
            L.axpby(ptr(dst), 1.0, ptr(tmp), 1.0, ptr(dst), dst.numel(), stream())
        self.accum = []


def pack_input(run, view, c_pad, ph, pw):
    """[N, T, C, H, W] strided view (H, W contiguous) -> padded NHWC Act with c_pad channels
    (lidar_feat_nets.py:216-218 reshape; channel order t-major)."""
    n, t, c, h, w = view.shape
    if view.stride(4) != 1 or view.stride(3) != w:
        view = view.contiguous()
    x = Act(n, h, w, c_pad, ph, pw, device=run.device, needs_grad=False, split=USE_TC)
    L.pack_input(ptr(view), view.stride(0), view.stride(1), view.stride(2), t, c, x.t4, ptr(x.t), ptr(x.lo), stream())
    run.keep.append((x, view))
    return x


def conv_bn(run, x, cname, *args, **kwargs):
    if not NVTX:
        return _conv_bn(run, x, cname, *args, **kwargs)
    label = getattr(run, "trace_prefix", "") + cname
    n_tape = len(run.tape)
    torch.cuda.nvtx.range_push(label)
    try:
        return _conv_bn(run, x, cname, *args, **kwargs)
    finally:
        torch.cuda.nvtx.range_pop()
        for i in range(n_tape, len(run.tape)):
            run.tape[i] = _nvtx_wrapped(run.tape[i], label + "/bwd")


def _nvtx_wrapped(fn, label):
    def call():
        torch.cuda.nvtx.range_push(label)
        try:
            fn()
        finally:
            torch.cuda.nvtx.range_pop()
    return call


def _conv_bn(run, x, cname, bname, stride=(1, 1), pre_relu=False, relu=True, pool=None, ceil=False, res=None,
            res_mode=0, out=None, c_off=0, out_c=None, out_pad=(0, 0), feat=None, feat_ld=0, feat_off=0,
            out_f32=True, out_group=1, out_c_pad=None):
    """conv(+bias)(+ReLU if pre_relu) -> BatchNorm (batch statistics when run.training) -> (+res)(ReLU)(+res)
    -> optional 3x3 max-pool, or -> global average into ``feat`` [N, feat_ld] at column feat_off.

    ``out_f32=False``: the caller guarantees that the output is consumed only by fp16 tensor-core convolutions
    (forward, dgrad and wgrad), so no fp32 copy of it is written.  ``out_group=2``: the output feeds W-stride-2
    convolutions only; its fp16 planes are written in the pixel-pair layout (needs an even padded width, else
    ignored).  ``out_c_pad``: allocate the output with this many channels (> cout); the extra channels are zeros, so
    that a narrow tensor (a Fire squeeze output: 16 .. 80 channels) meets the 64-channel granularity of the fp16
    tensor-core kernels of its consumers.  Returns the output Act (None when ``feat``)."""
    p, st = run.params, stream()
    w = p[cname + ".weight"]
    b = p.get(cname + ".bias")
    gamma, beta = p[bname + ".weight"], p[bname + ".bias"]
    rm, rv = run.buffers[bname + ".running_mean"], run.buffers[bname + ".running_var"]
    cout, cin, kh, kw = w.shape
    assert x.c >= cin, (cname, x.c, cin)      # extra input channels are zeros (and get zero weights)
    cin_pad = x.c
    sh, sw = stride
    cph, cpw = (kh - 1) // 2, (kw - 1) // 2
    ho, wo = (x.h + 2 * cph - kh) // sh + 1, (x.w + 2 * cpw - kw) // sw + 1
    cv = L.Conv(kh, kw, sh, sw, cph, cpw)
    n = x.n
    pads_ok = x.ph >= cph and x.pw >= cpw
    pair = pair_ok(x, cin, cout, kh, kw, stride)
    fwd_f16 = x.h2 is not None and x.group == 1 and f16_ok(cin_pad, cout, stride) and pads_ok
    fwd_tc = (not fwd_f16) and x.lo is not None and tc_ok(cin_pad, cout, stride) and pads_ok
    # first layer (8-channel input): stride-1 kh x 3 convolution over the space-to-depth views (csrc/conv_s2d.cu)
    s2d_f16 = (x.h2 is not None and x.c == 8 and x.group == 1 and x.pw == 4 and x.ph >= cph and not x.needs_grad
               and first_layer_f16_ok(w.shape, stride, x.w))
    s2d = s2d_f16 or (USE_TC and x.lo is not None and x.c == 8 and sh == 1 and sw in (1, 2) and kw <= 7
                      and x.w % 4 == 0 and x.pw == 4 and x.ph >= cph and (4 // sw * cout) % 128 == 0
                      and not x.needs_grad)
    if s2d:
        fwd_f16 = fwd_tc = False
    assert fwd_f16 or pair or s2d_f16 or x.t is not None, (cname, "input has no fp32 plane and the fp16 path does not apply")
    flops = 2.0 * cout * ho * wo * cin * kh * kw * x.n
    if FLOPS is not None:
        k = "conv_fwd_tc" if (s2d or pair or fwd_f16 or fwd_tc) else "conv_fwd_simt"
        FLOPS[k] = FLOPS.get(k, 0.0) + flops
    y = Act(n, ho, wo, cout, device=run.device)
    # the batch statistics also bound the BN output (-> scale of the fp16 planes), so they are taken in eval mode too
    stats = run.zeros(2 * cout, dtype=torch.float64)
    act = L.ACT_RELU if pre_relu else L.ACT_NONE
    w_ohwi = w_lo = w_h2 = w_bound = None
    if s2d:
        R = 4 // sw
        x4_t4 = L.Tensor4(n, x.h, x.w // 4, 32, x.ph, 1)
        y4_t4 = L.Tensor4(n, ho, x.w // 4, R * cout, 0, 0)
        cv4 = L.Conv(kh, 3, 1, 1, cph, 1)
        bias4 = b.repeat(R) if b is not None else None
        stats4 = run.zeros(2 * R * cout, dtype=torch.float64)
        if s2d_f16:
            # folded split: a 4-pixel group of the packed input planes is one row of 64 halves (hi and lo of 32 values)
            x4_t4 = L.Tensor4(n, x.h, x.w // 4, 64, x.ph, 1)
            w_bound = run.empty(1)
            w_h2 = run.empty(R * cout, 2, kh * 3 * 64, dtype=torch.float16)
            L.weight_to_s2d_f16(ptr(w), cout, cin, kh, kw, sw, ptr(w_bound), ptr(w_h2), st)
            L.conv2d_fwd_f16_folded(x4_t4, ptr(x.h2), ptr(x.bound), ptr(w_h2), ptr(w_bound), ptr(bias4), cv4, act,
                                    y4_t4, ptr(y.t), ptr(stats4), st)
        else:
            w4, w4_lo = run.empty(R * cout, kh, 3, 32), run.empty(R * cout, kh, 3, 32)
            L.weight_to_s2d(ptr(w), cout, cin, kh, kw, sw, ptr(w4), ptr(w4_lo), st)
            L.conv2d_fwd(x4_t4, ptr(x.t), ptr(x.lo), ptr(w4), ptr(w4_lo), ptr(bias4), cv4, act, y4_t4, ptr(y.t),
                         ptr(stats4), st)
        L.fold_stats(ptr(stats4), R, cout, ptr(stats), st)
    elif pair:
        # [n, h, w/2, 2 cin] view of the same memory, stride-1 kh x kw2 convolution, even rows kept when sh == 2
        kw2 = 3 if kw > 1 else 1
        x2_t4 = L.Tensor4(n, x.h, x.w // 2, 2 * cin, x.ph, x.pw // 2)
        cv2 = L.Conv(kh, kw2, 1, 1, cph, (kw2 - 1) // 2)
        w_bound = run.empty(1)
        w_h2 = run.empty(cout, 2, kh * kw2 * 2 * cin, dtype=torch.float16)
        L.weight_pack_pair_f16(ptr(w), cout, cin, kh, kw, 0, 1, ptr(w_bound), ptr(w_h2), st)
        L.conv2d_fwd_f16(x2_t4, ptr(x.h2), ptr(x.bound), ptr(w_h2), ptr(w_bound), ptr(b),
                         L.Conv(kh, kw2, sh, 1, cph, (kw2 - 1) // 2), act, y.t4, ptr(y.t), ptr(stats), st)
    elif fwd_f16:
        w_bound = run.empty(1)
        w_h2 = run.empty(cout, 2, kh * kw * cin_pad, dtype=torch.float16)
        L.weight_pack_f16(ptr(w), cout, cin, kh, kw, cin_pad, 0, 1, ptr(w_bound), ptr(w_h2), st)
        L.conv2d_fwd_f16(x.t4, ptr(x.h2), ptr(x.bound), ptr(w_h2), ptr(w_bound), ptr(b), cv, act, y.t4, ptr(y.t),
                         ptr(stats), st)
    else:
        w_ohwi = run.empty(cout, kh, kw, cin_pad)
        w_lo = torch.empty_like(w_ohwi) if fwd_tc else None    # low-order TF32 plane of the weights
        L.weight_to_ohwi(ptr(w), cout, cin, kh, kw, cin_pad, ptr(w_ohwi), ptr(w_lo), st)
        L.conv2d_fwd(x.t4, ptr(x.t), ptr(x.lo) if fwd_tc else None, ptr(w_ohwi), ptr(w_lo), ptr(b), cv, act, y.t4,
                     ptr(y.t), ptr(stats), st)
    bnv = run.empty(4, cout)  # mean, invstd, scale, shift
    count = n * ho * wo
    # the output of a Fire concat has two producers with separate statistics: no common bound, no fp16 planes
    single = out is None and c_off == 0 and out_c is None
    out_bound = run.empty(1) if (single and feat is None) else None
    if res is not None and out_bound is not None and res.bound is None:
        out_bound = None
    momentum, eps = run.bn_cfg.get(bname, (BN_MOMENTUM, BN_EPS))
    # num_batches_tracked is advanced by the finalize kernel itself; momentum None = cumulative average (-1)
    nbt = run.buffers.get(bname + ".num_batches_tracked") if run.training else None
    L.bn_finalize(ptr(stats), count, cout, ptr(gamma), ptr(beta), ptr(rm), ptr(rv),
                  -1.0 if momentum is None else float(momentum), float(eps),
                  0 if run.training else 1, ptr(bnv[0]), ptr(bnv[1]), ptr(bnv[2]), ptr(bnv[3]),
                  ptr(res.bound) if (res is not None and out_bound is not None) else None, ptr(out_bound),
                  ptr(nbt), st)
    if MASK_TRACE is not None:
        _trace_mask(run, cname, y, bnv, pre_relu, relu, res, res_mode, c_off)
    bp = L.BnPool(1 if relu else 0, res_mode if res is not None else 0, 3 if pool else 1,
                  pool[0] if pool else 1, pool[1] if pool else 1, c_off, 1)
    idx = ymax = None
    dummy = y.t4
    if feat is not None:
        assert res is None and not pool
        L.spatial_mean_fwd(y.t4, ptr(y.t), ptr(bnv[2]), ptr(bnv[3]), 1 if relu else 0, ptr(feat), feat_ld, feat_off, st)
        result = None
    else:
        oh, ow = (pool_out(ho, pool[0], ceil), pool_out(wo, pool[1], ceil)) if pool else (ho, wo)
        if out is None:
            oc = out_c_pad or out_c or cout
            o16 = USE_TC and USE_F16 and out_bound is not None and oc % 64 == 0
            assert out_f32 or o16, (cname, "out_f32=False needs an fp16-capable output")
            out = Act(n, oh, ow, oc, out_pad[0], out_pad[1], device=run.device, f32=out_f32, f16=o16,
                      split=USE_TC and out_f32 and not o16 and oc % 32 == 0)
            out.bound = out_bound
            if out_c_pad and oc > cout:      # the BN pass also zeroes channels [cout, oc)
                assert not pool
                bp.zero_tail = 1
            if out_group == 2 and o16 and (ow + 2 * out_pad[1]) % 2 == 0:
                out.group = bp.out_group = 2
        assert out.h == oh and out.w == ow and out.n == n
        if pool and run.record:
            idx = run.empty(n, oh, ow, cout, dtype=torch.uint8)
            # BN directly followed by the pool (Simple-1): the backward pass takes its sums from the pooled side and
            # un-pools inside the apply pass, so dz never goes through HBM; needs y at the arg-max from this pass
            if (FUSED_POOL_BWD and not relu and res is None and single and out.c == cout and pool[0] in (1, 2)
                    and pool[1] in (1, 2) and n * (oh + 2 * out.ph) * (ow + 2 * out.pw) < 2 ** 31):
                ymax = run.empty(n, oh, ow, cout)
        L.bn_act_pool_fwd(y.t4, ptr(y.t), ptr(bnv[2]), ptr(bnv[3]), res.t4 if res is not None else dummy,
                          ptr(res.t) if res is not None else None, bp, out.t4, ptr(out.t), ptr(out.lo), ptr(out.h2),
                          ptr(out.bound) if out.h2 is not None else None, ptr(idx), ptr(ymax), st)
        if MASK_TRACE is not None and idx is not None:
            MASK_TRACE[getattr(run, "trace_prefix", "") + cname + "#pool"] = idx.permute(0, 3, 1, 2).cpu()
        result = out
    if not run.record:
        return result
    def bwd():
        st = stream()
        if feat is not None:
            dout = run.fgrad[id(feat)]
            src, dout_t4, ld = L.GRAD_AVG, dummy, feat_ld
            bpb = L.BnPool(bp.relu, 0, 1, 1, 1, feat_off)
        else:
            # the writer of channel 0 ran first in the forward pass, so it is the last reader of the gradient
            dout = run.agrad.pop(id(out)) if c_off == 0 else run.agrad[id(out)]
            src, dout_t4, ld = (L.GRAD_POOL if pool else L.GRAD_DIRECT), out.t4_unpadded, 0
            bpb = bp
        # which kernels consume dy: 3xF16 / 3xTF32 tensor-core kernels or the fp32 CUDA-core ones
        if s2d:
            wg, dg = ("f16" if s2d_f16 else "tf32"), None
        elif pair:
            wg, dg = "f16", ("f16" if x.needs_grad else None)
        else:
            wg = "f16" if (fwd_f16 and cout % 64 == 0) else ("tf32" if (fwd_tc and cout % 32 == 0) else "simt")
            dg = None
            if x.needs_grad:
                dg = ("f16" if (f16_ok(cout, cin_pad, stride)) else ("tf32" if tc_ok(cout, cin_pad, stride) else "simt"))
        # narrow layer on the 3xTF32 forward path (Fire squeeze: 16 / 48 / 80 output channels): its backward runs on the
        # tensor cores too when dy carries zero channels up to a multiple of 32 (written by dlio_bn_bwd_apply) and the
        # weights zero rows -- instead of the CUDA-core dgrad / wgrad kernels (1.9 ms of the PointSeg step)
        cb = cout
        if (NARROW_TC_BWD and not s2d and not pair and wg == "simt" and fwd_tc and ymax is None and cout % 4 == 0
                and tuple(stride) == (1, 1)):
            cb = (cout + 31) // 32 * 32
            wg = "tf32"
            if x.needs_grad:
                dg = "tf32" if tc_ok(cb, cin_pad, stride) else "simt"
        assert wg != "simt" or x.t is not None, (cname, "wgrad on the CUDA cores needs the fp32 input plane")
        modes = (wg, dg)
        if FLOPS is not None:
            k = "conv_wgrad_simt" if wg == "simt" else "conv_wgrad_tc"
            FLOPS[k] = FLOPS.get(k, 0.0) + flops
            if dg is not None:
                k = "conv_dgrad_simt" if dg == "simt" else "conv_dgrad_tc"
                FLOPS[k] = FLOPS.get(k, 0.0) + flops
        sums = run.zeros(2 * cout + 1, dtype=torch.float64)
        dz = None
        if ymax is not None:
            windows = (3 if pool[0] == 1 else 2) * (3 if pool[1] == 1 else 2)
            L.pool_bwd_sums(dout_t4, ptr(dout), 0, cout, ptr(ymax), ptr(bnv[0]), ptr(bnv[1]), windows, ptr(sums), st)
        else:
            # no pooling, no residual, one producer: dz is dout itself (masked by the ReLU, if any, inside the apply
            # pass), so this pass only takes the sums
            direct = (SKIP_DZ and src == L.GRAD_DIRECT and res is None and single and out.c == cout
                      and dout.is_contiguous())
            dz = None if direct else run.empty(n, ho, wo, cout)
            dres, dres_c, dres_acc = None, 0, 0
            if res is not None and res.needs_grad:
                partial = res.c != cout
                dres, existed = run.grad_slot(res, zeroed=partial)
                dres_c, dres_acc = res.c, 1 if (existed or partial) else 0
            L.bn_act_pool_bwd_reduce(y.t4, ptr(y.t), ptr(bnv[2]), ptr(bnv[3]), ptr(bnv[0]), ptr(bnv[1]),
                                     res.t4 if res is not None else dummy, ptr(res.t) if res is not None else None,
                                     bpb, src, dout_t4, ptr(dout), ld, ptr(idx), ptr(dz), ptr(dres), dres_c, dres_acc,
                                     ptr(sums), 1 if "f16" in modes else 0, st)
        # backward on the tensor cores: dy is produced as padded split planes.  wgrad wants dy on x's padded grid,
        # dgrad wants pads >= (k - 1 - pad); x's pads satisfy both for the "same" stride-1 convolutions.
        on_x_grid = wg in ("f16", "tf32")
        dpad = (x.ph, x.pw) if on_x_grid else ((kh - 1 - cph, kw - 1 - cpw) if dg in ("f16", "tf32") else (0, 0))
        if s2d:
            dpad = (x.ph, 4 // sw)    # x's padded grid expressed in output pixels
        if pair:
            # dy on the padded grid of the pixel-pair view; with sh == 2 its rows are spread over the even input rows
            dya = Act(n, x.h, wo, cout, x.ph, x.pw // 2, device=run.device, f32=False, f16=True)
        else:
            dya = Act(n, ho, wo, cb, dpad[0], dpad[1], device=run.device, f32="simt" in modes or "tf32" in modes,
                      split="tf32" in modes, f16="f16" in modes)
        dya.bound = run.empty(1) if dya.h2 is not None else None
        dgamma, dbeta = run.param_grad(bname + ".weight", gamma), run.param_grad(bname + ".bias", beta)
        dbs = run.zeros(cout, dtype=torch.float64) if b is not None else None
        if ymax is not None:
            L.bn_pool_bwd_apply(y.t4, ptr(y.t), bp, dout_t4, ptr(dout), ptr(idx), ptr(sums), count, ptr(bnv[2]),
                                ptr(bnv[0]), ptr(bnv[1]), 1 if pre_relu else 0, 1 if run.training else 0, dya.t4,
                                ptr(dya.t), ptr(dya.lo), ptr(dya.h2), ptr(dya.bound), ptr(dgamma), ptr(dbeta), ptr(dbs), st)
        else:
            L.bn_bwd_apply(y.t4, ptr(y.t), ptr(dz if dz is not None else dout), ptr(sums), count, ptr(bnv[2]),
                           ptr(bnv[0]), ptr(bnv[1]), 1 if pre_relu else 0, 1 if run.training else 0, dya.t4, ptr(dya.t),
                           ptr(dya.lo), ptr(dya.h2), ptr(dya.bound), ptr(dgamma), ptr(dbeta), ptr(dbs),
                           ptr(bnv[3]), 1 if (dz is None and bpb.relu) else 0, st)
        if DEBUG_TRACE is not None and dz is not None:
            DEBUG_TRACE[(id(run), cname)] = dict(dout=dout.clone(), dz=dz.clone(), sums=sums.clone(),
                                                 dy=(dya.t if dya.t is not None else dya.h2).clone(), y=y.t.clone(),
                                                 bnv=bnv.clone())
        del dz
        if b is not None:
            L.f64_to_f32(ptr(dbs), ptr(run.param_grad(cname + ".bias", b)), cout, st)
        dw = run.param_grad(cname + ".weight", w)
        if pair:
            cv2s1 = L.Conv(kh, kw2, 1, 1, cph, (kw2 - 1) // 2)
            dw2 = run.empty(cout, kh, kw2, 2 * cin)
            L.conv2d_bwd_weight_f16(x2_t4, ptr(x.h2), ptr(x.bound), dya.t4, ptr(dya.h2), ptr(dya.bound), cv2s1,
                                    ptr(dw2), st)
            L.weight_grad_from_pair(ptr(dw2), cout, cin, kh, kw, ptr(dw), st)
            if dg == "f16":
                wt_h2 = run.empty(2 * cin, 2, kh * kw2 * cout, dtype=torch.float16)
                L.weight_pack_pair_f16(ptr(w), cout, cin, kh, kw, 1, 0, ptr(w_bound), ptr(wt_h2), st)
                run.add_grad(x, lambda buf, acc=0: L.conv2d_bwd_data_f16(
                    dya.t4, ptr(dya.h2), ptr(dya.bound), ptr(wt_h2), ptr(w_bound), cv2s1,
                    L.Tensor4(n, x.h, x.w // 2, 2 * cin, 0, 0), ptr(buf), acc, st), accumulates=True)
            return
        if s2d_f16:
            # dy planes per output pixel [64 hi | 64 lo]; R output pixels per row of the space-to-depth view
            dw64 = run.empty(R * cout, kh, 3, 64)
            L.conv2d_bwd_weight_f16_folded(x4_t4, ptr(x.h2), ptr(x.bound), L.Tensor4(n, ho, x.w // 4, R * cout, x.ph, 1),
                                           ptr(dya.h2), ptr(dya.bound), cv4, ptr(dw64), st)
            L.weight_grad_from_s2d_f16(ptr(dw64), cout, cin, kh, kw, sw, ptr(dw), st)
        elif s2d:
            dw4 = run.empty(R * cout, kh, 3, 32)
            L.conv2d_bwd_weight(x4_t4, ptr(x.t), ptr(x.lo), L.Tensor4(n, ho, x.w // 4, R * cout, x.ph, 1), ptr(dya.t),
                                ptr(dya.lo), cv4, ptr(dw4), st)
            L.weight_grad_from_s2d(ptr(dw4), cout, cin, kh, kw, sw, ptr(dw), st)
        else:
            dw_ohwi = run.empty(cb, kh, kw, cin_pad)      # rows [cout, cb): gradients of the zero rows, not read
            if wg == "f16":
                L.conv2d_bwd_weight_f16(x.t4, ptr(x.h2), ptr(x.bound), dya.t4, ptr(dya.h2), ptr(dya.bound), cv,
                                        ptr(dw_ohwi), st)
            else:
                L.conv2d_bwd_weight(x.t4, ptr(x.t), ptr(x.lo) if wg == "tf32" else None, dya.t4, ptr(dya.t),
                                    ptr(dya.lo) if wg == "tf32" else None, cv, ptr(dw_ohwi), st)
            L.weight_grad_to_oihw(ptr(dw_ohwi), cout, cin, kh, kw, cin_pad, ptr(dw), st)
        if dg == "f16":
            wb = w_bound if w_bound is not None else run.empty(1)
            wt_h2 = run.empty(cin_pad, 2, kh * kw * cout, dtype=torch.float16)
            L.weight_pack_f16(ptr(w), cout, cin, kh, kw, cin_pad, 1, 0 if w_bound is not None else 1, ptr(wb),
                              ptr(wt_h2), st)
            run.add_grad(x, lambda buf, acc=0: L.conv2d_bwd_data_f16(dya.t4, ptr(dya.h2), ptr(dya.bound), ptr(wt_h2),
                                                                     ptr(wb), cv, x.t4_unpadded, ptr(buf), acc, st),
                         accumulates=True)
        elif dg is not None:
            wo_, wl_ = w_ohwi, w_lo
            if wo_ is None:
                wo_ = run.empty(cout, kh, kw, cin_pad)
                L.weight_to_ohwi(ptr(w), cout, cin, kh, kw, cin_pad, ptr(wo_), None, st)
            if cb != cout:
                wo_p, wl_p = run.zeros(cb, kh, kw, cin_pad), run.zeros(cb, kh, kw, cin_pad)
                wo_p[:cout].copy_(wo_)
                if wl_ is not None:
                    wl_p[:cout].copy_(wl_)
                wo_, wl_ = wo_p, wl_p
            wt_hi = wt_lo = None
            if dg == "tf32":
                wt_hi, wt_lo = torch.empty_like(wo_), torch.empty_like(wo_)
                L.weight_flip_transpose(ptr(wo_), cb, cin_pad, kh, kw, ptr(wt_hi), ptr(wt_lo), st)
            run.add_grad(x, lambda buf, acc=0: L.conv2d_bwd_data(dya.t4, ptr(dya.t), ptr(dya.lo), ptr(wo_), ptr(wl_),
                                                                 ptr(wt_hi), ptr(wt_lo), cv, x.t4_unpadded, ptr(buf),
                                                                 acc, st), accumulates=True)

    run.tape.append(bwd)
    run.keep.append((x, y, out, res))
    return result


def se_layer(run, x, prefix, out_pad=(0, 0)):
    """SELayer (pointseg_modules.py:203-221): x * sigmoid(W2 relu(W1 mean_hw(x))), no biases."""
    p, st = run.params, stream()
    w1, w2 = p[prefix + "fc.0.weight"], p[prefix + "fc.2.weight"]
    n, c, cr = x.n, x.c, w1.shape[0]
    m = run.empty(n, c)
    L.spatial_mean_fwd(x.t4, ptr(x.t), None, None, 0, ptr(m), c, 0, st)
    hid = run.empty(n, cr)
    L.linear_fwd(ptr(m), c, ptr(w1), None, n, cr, c, L.ACT_RELU, ptr(hid), cr, st)
    gate = run.empty(n, c)
    L.linear_fwd(ptr(hid), cr, ptr(w2), None, n, c, cr, L.ACT_SIGMOID, ptr(gate), c, st)
    out = Act(n, x.h, x.w, c, out_pad[0], out_pad[1], device=run.device, split=USE_TC and c % 32 == 0)
    L.channel_scale_fwd(x.t4, ptr(x.t), ptr(gate), out.t4, ptr(out.t), ptr(out.lo), st)
    out.bound = x.bound       # gate <= 1
    if not run.record:
        return out

    def bwd():
        st = stream()
        dout = run.agrad.pop(id(out))
        dgate = run.empty(n, c)
        L.spatial_dot(out.t4_unpadded, ptr(dout), x.t4, ptr(x.t), ptr(dgate), st)
        dhid, dw2, scr = run.empty(n, cr), run.param_grad(prefix + "fc.2.weight", w2), run.empty(n, c)
        L.linear_bwd(ptr(hid), cr, ptr(w2), ptr(gate), c, ptr(dgate), c, n, c, cr, L.ACT_SIGMOID, ptr(dhid), cr,
                     ptr(dw2), None, ptr(scr), st)
        dm, dw1, scr2 = run.empty(n, c), run.param_grad(prefix + "fc.0.weight", w1), run.empty(n, cr)
        L.linear_bwd(ptr(m), c, ptr(w1), ptr(hid), cr, ptr(dhid), cr, n, cr, c, L.ACT_RELU, ptr(dm), c, ptr(dw1),
                     None, ptr(scr2), st)
        run.add_grad(x, lambda buf: L.channel_scale_bwd(ptr(dout), ptr(gate), ptr(dm), n, x.h * x.w, c, ptr(buf), st))

    run.tape.append(bwd)
    run.keep.append((x, out))
    return out


def max_pool(run, x, stride, ceil=False, out_pad=(0, 0), name=None):
    """MaxPool2d(3, stride, padding=1) on its own (PointSeg pools, pointseg_net.py:28,35,43,50)."""
    st = stream()
    oh, ow = pool_out(x.h, stride[0], ceil), pool_out(x.w, stride[1], ceil)
    out = Act(x.n, oh, ow, x.c, out_pad[0], out_pad[1], device=run.device, split=USE_TC and x.c % 32 == 0)
    bp = L.BnPool(0, 0, 3, stride[0], stride[1], 0)
    idx = run.empty(x.n, oh, ow, x.c, dtype=torch.uint8) if run.record else None
    L.bn_act_pool_fwd(x.t4, ptr(x.t), None, None, x.t4, None, bp, out.t4, ptr(out.t), ptr(out.lo), None, None,
                      ptr(idx), None, st)
    if MASK_TRACE is not None and idx is not None and name:
        MASK_TRACE[getattr(run, "trace_prefix", "") + name] = idx.permute(0, 3, 1, 2).cpu()
    out.bound = x.bound
    if not run.record:
        return out

    def bwd():
        dout = run.agrad.pop(id(out))
        if FUSED_POOL_BWD and stride[0] in (1, 2) and stride[1] in (1, 2) and x.c % 16 == 0 and x.t is not None:
            # the un-pooling of dlio_bn_pool_bwd_apply with an identity BatchNorm (no statistics, scale 1): rows of dout
            # and of the arg-max bytes staged by bulk copies instead of the generic gather (233 -> ~90 us per PointSeg pool)
            run.add_grad(x, lambda buf: L.bn_pool_bwd_apply(
                x.t4, ptr(x.t), bp, out.t4_unpadded, ptr(dout), ptr(idx), None, 1, None, None, None, 0, 0,
                x.t4_unpadded, ptr(buf), None, None, None, None, None, None, stream()))
            return
        run.add_grad(x, lambda buf: L.bn_act_pool_bwd_reduce(
            x.t4, ptr(x.t), None, None, None, None, x.t4, None, bp, L.GRAD_POOL, out.t4_unpadded, ptr(dout), 0,
            ptr(idx), ptr(buf), None, 0, 0, None, 0, stream()))

    run.tape.append(bwd)
    run.keep.append((x, out))
    return out


def global_avg(run, x, feat, feat_ld, feat_off):
    """adaptive_avg_pool2d((1,1)) of an activation into feat[:, feat_off : feat_off + c]."""
    L.spatial_mean_fwd(x.t4, ptr(x.t), None, None, 0, ptr(feat), feat_ld, feat_off, stream())
    if not run.record:
        return
    bp = L.BnPool(0, 0, 1, 1, 1, feat_off)

    def bwd():
        dout = run.fgrad[id(feat)]
        run.add_grad(x, lambda buf: L.bn_act_pool_bwd_reduce(
            x.t4, ptr(x.t), None, None, None, None, x.t4, None, bp, L.GRAD_AVG, x.t4, ptr(dout), feat_ld, None,
            ptr(buf), None, 0, 0, None, 0, stream()))

    run.tape.append(bwd)
    run.keep.append((x,))
