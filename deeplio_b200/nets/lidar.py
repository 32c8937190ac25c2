"""LiDAR feature nets: two independent encoders (xyz pair, normal pair) + add/sub/cat + fc1 -> 128.

Drop-in classes for the reference's ``LidarSimpleFeat1``, ``LidarPointSegFeat``, ``LidarResNetFeat`` and
``LidarFlowNetFeat`` (lidar_feat_nets.py:46-237) with their encoders ``FeatureNetSimple1`` (:270-342),
``FlowNetEncoder`` (:240-267), ``PSEncoder`` (pointseg_net.py:9-82) and ``ResNetEncoder`` (resnet.py:14-112).
Same constructor signatures, ``state_dict`` keys and outputs; the convolutional work runs as a tape of
fused C-ABI calls (deeplio_b200.engine) instead of nn.Module forwards.

Differences from the reference, on purpose:
  * the output shape is computed analytically (the reference runs a dummy CPU forward, :33-40);
  * ``fusion: cat`` builds ``fc1`` with 2*C inputs -- the reference hard-codes C (:66,116,161,204) and
    crashes on the first forward;
  * Simple-1 ``bypass: true`` raises at construction: the reference adds a 128- to a 256-channel tensor
    (:322-326) and cannot run either.
"""
import torch
from torch import nn

from .. import engine as E
from .. import functional as Fn
from .. import _lib as L
from .._lib import ptr
from ..config import get_config_container
from ..data import PairedFrames, frames_bound, pair_gather
from .base import BaseNet, ParamTree, conv_params, require_cuda


# ----------------------------------------------------------------------------- encoder definitions
class _Encoder(ParamTree):
    """Parameter container + forward program of one encoder.  ``out_channels`` = pooled feature width."""
    out_channels = 0
    first_pad = (0, 0)
    first_conv, first_stride = None, (1, 1)      # parameter prefix and stride of the first convolution

    def _conv(self, name, cin, cout, k, bias):
        return self.put(name, conv_params(cin, cout, k, bias))

    def _bn(self, name, c, momentum=0.1):
        return self.put(name, nn.BatchNorm2d(c, momentum=momentum))

    def bn_cfg(self):
        """BatchNorm2d name -> (momentum, eps) as the modules declare them (engine.conv_bn)."""
        cfg = self.__dict__.get("_bn_cfg")
        if cfg is None:
            cfg = {k: (m.momentum, m.eps) for k, m in self.named_modules() if isinstance(m, nn.BatchNorm2d)}
            self.__dict__["_bn_cfg"] = cfg
        return cfg

    def program(self, run, x, feat, ld, off):
        raise NotImplementedError


class Simple1Encoder(_Encoder):
    """conv(+bias) -> ReLU -> BN, seven times; ceil-mode max-pools after blocks 1, 2, 4, 6; global mean."""
    out_channels = 512
    first_pad = (2, 3)
    first_conv, first_stride = "conv1", (1, 2)
    SPEC = [(1, 64, (5, 7)), (2, 128, (3, 5)), (3, 128, 3), (4, 256, 3), (5, 256, 3), (6, 512, 3), (7, 512, 3)]

    def __init__(self, cin, bypass=False):
        super().__init__()
        if bypass:
            raise ValueError("lidar-feat-simple-1: bypass=true adds tensors of different channel counts "
                             "(reference lidar_feat_nets.py:322-326) and is not runnable")
        for i, cout, k in self.SPEC:
            self._conv("conv%d" % i, cin, cout, k, True)
            self._bn("bn%d" % i, cout)
            cin = cout

    def program(self, run, x, feat, ld, off):
        # every intermediate feeds only a stride-1 convolution with >= 128 output channels: fp16 planes suffice
        def blk(i, t, stride=(1, 1), pool=None, pad=(1, 1), **kw):
            return E.conv_bn(run, t, "conv%d" % i, "bn%d" % i, stride, pre_relu=True, relu=False, pool=pool,
                             ceil=True, out_pad=pad, out_f32=not (E.USE_TC and E.USE_F16) or "feat" in kw, **kw)
        t = blk(1, x, (1, 2), pool=(1, 2), pad=(1, 2))
        t = blk(2, t, pool=(1, 2))
        t = blk(3, t)
        t = blk(4, t, pool=(2, 2))
        t = blk(5, t)
        t = blk(6, t, pool=(2, 2))
        blk(7, t, feat=feat, feat_ld=ld, feat_off=off)


class FlowNetEncoder(_Encoder):
    """Nine conv(no bias) -> BN -> ReLU blocks (base_net.py:55-71), global mean."""
    out_channels = 1024
    first_pad = (2, 3)
    first_conv, first_stride = "conv1.0", (1, 2)
    SPEC = [("conv1", 64, (5, 7), (1, 2)), ("conv2", 128, (3, 5), (1, 2)), ("conv3", 256, (3, 5), (1, 2)),
            ("conv3_1", 256, 3, (1, 1)), ("conv4", 512, 3, (2, 2)), ("conv4_1", 512, 3, (1, 1)),
            ("conv5", 512, 3, (2, 2)), ("conv5_1", 512, 3, (1, 1)), ("conv6", 1024, 3, (2, 2))]

    def __init__(self, cin):
        super().__init__()
        for name, cout, k, _ in self.SPEC:
            self._conv(name + ".0", cin, cout, k, False)
            self._bn(name + ".1", cout)
            cin = cout

    def program(self, run, x, feat, ld, off):
        t = x
        for i, (name, _, k, stride) in enumerate(self.SPEC):
            last = i == len(self.SPEC) - 1
            if last:
                E.conv_bn(run, t, name + ".0", name + ".1", stride, feat=feat, feat_ld=ld, feat_off=off)
            else:
                cout, (ncout, nk, nstride) = self.SPEC[i][1], self.SPEC[i + 1][1:]
                nkh, nkw = (nk, nk) if isinstance(nk, int) else nk
                kh, kw = (k, k) if isinstance(k, int) else k
                w_out = (t.w + 2 * ((kw - 1) // 2) - kw) // stride[1] + 1
                h_out = (t.h + 2 * ((kh - 1) // 2) - kh) // stride[0] + 1
                # the next layer is the only consumer: no fp32 copy when it runs on the fp16 tensor-core kernels
                f32 = not E.consumer_reads_f16_only(cout, ncout, nk, nstride, w_out, h_out)
                if nstride[1] == 2:
                    # the next layer is W-strided: pixel-pair layout, even row pads (engine.pair_ok)
                    t = E.conv_bn(run, t, name + ".0", name + ".1", stride, out_pad=((nkh - 1) // 2, 2), out_group=2,
                                  out_f32=f32)
                else:
                    t = E.conv_bn(run, t, name + ".0", name + ".1", stride, out_pad=((nkh - 1) // 2, (nkw - 1) // 2),
                                  out_f32=f32)


class ResNetEncoder(_Encoder):
    """conv1 5x7 (bias) -> BN -> ReLU -> max-pool s(1,2); layers [3,3,3,2] of BasicBlocks
    (torchvision resnet.py:59-105) with first-block strides (1,2),(1,2),(2,2),(2,2); global mean."""
    out_channels = 512
    first_pad = (2, 3)
    first_conv, first_stride = "conv1", (1, 1)
    LAYERS = [("layer1", 3, 64, (1, 2)), ("layer2", 3, 128, (1, 2)), ("layer3", 3, 256, (2, 2)),
              ("layer4", 2, 512, (2, 2))]

    def __init__(self, cin):
        super().__init__()
        self._conv("conv1", cin, 64, (5, 7), True)
        self._bn("bn1", 64)
        inpl = 64
        for lname, nblk, planes, _ in self.LAYERS:
            for b in range(nblk):
                q = "%s.%d." % (lname, b)
                self._conv(q + "conv1", inpl, planes, 3, False)
                self._bn(q + "bn1", planes)
                self._conv(q + "conv2", planes, planes, 3, False)
                self._bn(q + "bn2", planes)
                if b == 0:
                    self._conv(q + "downsample.0", inpl, planes, 1, False)
                    self._bn(q + "downsample.1", planes)
                inpl = planes
        for m in self.modules():  # resnet.py:51-56
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def program(self, run, x, feat, ld, off):
        # the first block of every layer is W-strided (conv1 and the 1x1 downsample): its input is produced in the
        # pixel-pair layout with even row pads (engine.pair_ok)
        t = E.conv_bn(run, x, "conv1", "bn1", (1, 1), pool=(1, 2), ceil=False, out_pad=(1, 2), out_group=2)
        for li, (lname, nblk, _, stride) in enumerate(self.LAYERS):
            for b in range(nblk):
                q = "%s.%d." % (lname, b)
                s = stride if b == 0 else (1, 1)
                planes = run.params[q + "conv2.weight"].shape[0]
                w_out = (t.w - 1) // s[1] + 1
                # conv1's output feeds conv2 only (stride 1, planes -> planes)
                o = E.conv_bn(run, t, q + "conv1", q + "bn1", s, out_pad=(1, 1),
                              out_f32=not E.consumer_reads_f16_only(planes, planes, 3, (1, 1), w_out))
                idn = t
                if b == 0:
                    idn = E.conv_bn(run, t, q + "downsample.0", q + "downsample.1", s, relu=False)
                to_strided = b == nblk - 1 and li + 1 < len(self.LAYERS)
                t = E.conv_bn(run, o, q + "conv2", q + "bn2", (1, 1), relu=True, res=idn, res_mode=1,
                              out_pad=(1, 2) if to_strided else (1, 1), out_group=2 if to_strided else 1)
        E.global_avg(run, t, feat, ld, off)


class PointSegEncoder(_Encoder):
    """PSEncoder: conv1a 3x5 s(1,2) -> BN -> ReLU -> pool; five Fire blocks with SE layers and pools."""
    out_channels = 768
    first_pad = (1, 2)
    first_conv, first_stride = "conv1a.0", (1, 2)
    # block -> entries: ('F', cin, squeeze, expand) fire | ('S', c) SE | ('P', stride) pool
    BLOCKS = [("fire_blk1", [("F", 64, 16, 64), ("F", 128, 16, 64), ("S", 128), ("P", (1, 2))]),
              ("fire_blk2", [("F", 128, 32, 128), ("F", 256, 32, 128), ("S", 256), ("P", (1, 2))]),
              ("fire_blk3", [("F", 256, 48, 192), ("F", 384, 48, 192), ("F", 384, 64, 256), ("F", 512, 64, 256),
                             ("S", 512), ("P", (2, 2))]),
              ("fire_blk4", [("F", 512, 64, 256), ("F", 512, 64, 256), ("S", 512), ("P", (2, 2))]),
              ("fire_blk5", [("F", 512, 80, 384), ("FN", 768, 80, 384)])]

    def __init__(self, cin, bypass="simple", bn_d=0.1):
        super().__init__()
        if bypass not in (None, False, "simple"):
            # the reference's "complex" bypass adds a 1x1 `upsample` convolution per Fire (pointseg_modules.py
            # :110-113,134-137): different parameters and state_dict keys -- not built here, so refuse loudly
            raise ValueError("lidar-feat-pointseg: bypass=%r is not implemented (supported: null, 'simple')" % (bypass,))
        self.bypass = bypass
        self._conv("conv1a.0", cin, 64, (3, 5), True)
        self._bn("conv1a.1", 64, bn_d)
        for bname, entries in self.BLOCKS:
            for i, e in enumerate(entries):
                q = "%s.%d." % (bname, i)
                if e[0] in ("F", "FN"):
                    _, ci, sq, ex = e
                    self._conv(q + "squeeze", ci, sq, 1, True)
                    self._bn(q + "squeeze_bn", sq, bn_d)
                    self._conv(q + "expand1x1", sq, ex, 1, True)
                    self._bn(q + "expand1x1_bn", ex, bn_d)
                    self._conv(q + "expand3x3", sq, ex, 3, True)
                    self._bn(q + "expand3x3_bn", ex, bn_d)
                elif e[0] == "S":
                    c = e[1]
                    self.put(q + "fc.0", nn.Linear(c, c // 2, bias=False))
                    self.put(q + "fc.2", nn.Linear(c // 2, c, bias=False))

    def _fire(self, run, q, x, ex, bypass):
        """Fire (pointseg_modules.py:110-141): squeeze1x1 -> {expand1x1 || expand3x3} -> cat (+ x)."""
        # the squeeze output (16 .. 80 channels) is allocated with a multiple of 64 channels, zeros beyond the real
        # ones, so that both expand convolutions run on the fp16 tensor-core kernels
        sq = run.params[q + "squeeze.weight"].shape[0]
        # ... and its only consumers are those two convolutions, so no fp32 copy of it is written
        c_pad = (sq + 63) // 64 * 64
        s = E.conv_bn(run, x, q + "squeeze", q + "squeeze_bn", out_pad=(1, 1), out_c_pad=c_pad,
                      out_f32=not E.consumer_reads_f16_only(c_pad, ex, 3, (1, 1), x.w, x.h))
        res = x if (bypass and x.c == 2 * ex) else None
        out = E.conv_bn(run, s, q + "expand1x1", q + "expand1x1_bn", res=res, res_mode=2, c_off=0, out_c=2 * ex,
                        out_pad=(1, 1))
        E.conv_bn(run, s, q + "expand3x3", q + "expand3x3_bn", res=res, res_mode=2, out=out, c_off=ex)
        return out

    def program(self, run, x, feat, ld, off):
        t = E.conv_bn(run, x, "conv1a.0", "conv1a.1", (1, 2), pool=(1, 2), ceil=False, out_pad=(1, 1))
        for bname, entries in self.BLOCKS:
            for i, e in enumerate(entries):
                q = "%s.%d." % (bname, i)
                if e[0] == "F":
                    t = self._fire(run, q, t, e[3], self.bypass == "simple")
                elif e[0] == "FN":
                    t = self._fire(run, q, t, e[3], False)
                elif e[0] == "S":
                    t = E.se_layer(run, t, q, out_pad=(1, 1))
                else:
                    t = E.max_pool(run, t, e[1], ceil=False, out_pad=(1, 1), name=q + "pool")
        E.global_avg(run, t, feat, ld, off)  # adaptive_avg_pool2d outside the encoder (lidar_feat_nets.py:84-85)


# ----------------------------------------------------------------------------- siamese pair as one autograd node
class _EncoderPair(torch.autograd.Function):
    """encoder1(xyz) (+|-|cat) encoder2(normals) -> [N, C] (or [N, 2C]); forward records a tape per encoder,
    backward replays them."""

    @staticmethod
    def forward(ctx, net, record, xyz, normals, *params):
        enc = (net.encoder1, net.encoder2)
        names = net._param_names
        n1 = len(names[0])
        groups = (params[:n1], params[n1:])
        dev = xyz.device
        c = enc[0].out_channels
        cat = net.fusion == "cat"
        n = xyz.shape[0] * xyz.shape[1] if isinstance(xyz, PairedFrames) else xyz.shape[0]
        feats = [torch.empty((n, 2 * c if cat else c), device=dev, dtype=torch.float32)]
        feats.append(feats[0] if cat else torch.empty((n, c), device=dev, dtype=torch.float32))
        runs = []
        bounds = {}
        streams = _fork(dev)
        for e, view in enumerate((xyz, normals)):
            pd = dict(zip(names[e], groups[e]))
            bufs = dict(enc[e].named_buffers())
            run = E.Run(pd, bufs, dev, net.training, record)
            run.param_objs = dict(enc[e].named_parameters())
            run.bn_cfg = enc[e].bn_cfg()
            run.trace_prefix = "encoder%d." % (e + 1)
            with torch.cuda.stream(streams[e]):
                # row pads of 4 pixels: space-to-depth first layer
                if isinstance(view, PairedFrames):
                    # frames paired on the fly (deeplio_b200.data): no [B,S,2,C,H,W] tensor, no reshape copy
                    w1 = pd.get(str(enc[e].first_conv) + ".weight")
                    if (w1 is not None and view.frames.is_contiguous()
                            and E.first_layer_f16_ok(w1.shape, enc[e].first_stride, view.shape[5])):
                        # packed fp16 planes (per pixel [8 hi | 8 lo]) for the folded-split first layer; one bound
                        # per frames tensor, shared by the two encoders when both views are cut from it
                        key = view.frames.data_ptr()
                        if bounds.get("key") != key:
                            bounds["key"], bounds["bound"] = key, frames_bound(view)
                        h2, bnd = pair_gather(dev, view, 8, enc[e].first_pad[0], 4, False, f16=True, bound=bounds["bound"])
                        x0 = E.Act(n, view.shape[4], view.shape[5], 8, enc[e].first_pad[0], 4, needs_grad=False, f32=False)
                        x0.h2, x0.bound = h2, bnd
                    else:
                        hi, lo = pair_gather(dev, view, 8, enc[e].first_pad[0], 4, E.USE_TC)
                        x0 = E.Act(n, view.shape[4], view.shape[5], 8, enc[e].first_pad[0], 4, t=hi, lo=lo, needs_grad=False)
                    run.keep.append((x0, view.frames))
                else:
                    x0 = E.pack_input(run, view, 8, enc[e].first_pad[0], 4)
                enc[e].program(run, x0, feats[e], feats[e].shape[1], c if (cat and e == 1) else 0)
            runs.append(run)
        _join(streams)
        if cat:
            y = feats[0]
        else:
            y = torch.empty_like(feats[0])
            L.axpby(ptr(feats[0]), 1.0, ptr(feats[1]), 1.0 if net.fusion == "add" else -1.0, ptr(y), y.numel(),
                    E.stream())
        if record:
            ctx.runs, ctx.feats, ctx.fusion, ctx.names = runs, feats, net.fusion, names
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        runs, feats = ctx.runs, ctx.feats
        grads = []
        ds = [dy, dy]
        if ctx.fusion not in ("cat", "add"):
            ds[1] = torch.empty_like(dy)
            L.axpby(ptr(dy), -1.0, ptr(dy), 0.0, ptr(ds[1]), dy.numel(), E.stream())
        streams = _fork(dy.device)
        for e, run in enumerate(runs):
            run.fgrad[id(feats[e])] = ds[e]
            with torch.cuda.stream(streams[e]):
                run.backward()
            grads.extend(run.pgrad.get(k) for k in ctx.names[e])
        _join(streams)
        ctx.runs = None
        return (None, None, None, None, *grads)


_side_streams = {}


def _fork(dev):
    """Streams the two encoders are enqueued on: (current, current) or, with engine.ENC_STREAMS, (current, side) with
    the side stream ordered after everything enqueued on the current one so far.  Everything a run allocates stays
    on its stream for the forward AND the backward pass, and every fork starts with this wait, so the caching
    allocator's per-stream reuse is safe."""
    cur = torch.cuda.current_stream(dev)
    if not E.ENC_STREAMS:
        return (cur, cur)
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(dev)
    side.wait_stream(cur)
    return (cur, side)


def _join(streams):
    if streams[1] is not streams[0]:
        streams[0].wait_stream(streams[1])


class BaseLidarFeatNet(BaseNet):
    """Common part of the four LiDAR feature nets (lidar_feat_nets.py:12-43)."""
    tail = "relu_then_drop"   # pointseg / flownet: drop(relu(fc1(x)))

    def __init__(self, input_shape, cfg):
        super().__init__()
        self.p = cfg["dropout"]
        self.fusion = cfg["fusion"]
        self.cfg_container = get_config_container()
        self.seq_size = self.cfg_container.seq_size
        self.timestamps = self.cfg_container.timestamps
        self.combinations = self.cfg_container.combinations
        self.input_shape = input_shape
        self.split_backward = False
        self._cut = None

    def _finish(self, width):
        self.fc1 = nn.Linear(width * (2 if self.fusion == "cat" else 1), 128)
        self._param_names = tuple(tuple(k for k, _ in e.named_parameters()) for e in (self.encoder1, self.encoder2))
        self.output_shape = torch.Size([1, self.seq_size, 128])

    def forward(self, x):
        imgs_xyz, imgs_normals = x[0], x[1]
        require_cuda(imgs_xyz, "lidar input")
        b, s, t, c, h, w = imgs_xyz.shape
        if imgs_xyz.dtype != torch.float32 or imgs_normals.dtype != torch.float32:
            raise RuntimeError("deeplio_b200: lidar inputs must be float32")

        def as5(v):  # [B,S,T,C,H,W] -> [B*S,T,C,H,W] without copying when (b, s) collapse
            if v.stride(0) != s * v.stride(1):
                v = v.contiguous()
            return v.as_strided((b * s, t, c, h, w), (v.stride(1),) + tuple(v.stride()[2:]), v.storage_offset())
        params = [p for e in (self.encoder1, self.encoder2) for _, p in e.named_parameters()]
        record = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        views = [v if isinstance(v, PairedFrames) else as5(v) for v in (imgs_xyz, imgs_normals)]
        y = _EncoderPair.apply(self, record, views[0], views[1], *params)
        if self.split_backward and y.requires_grad:
            # cut the autograd graph at the encoders' feature vector: ``loss.backward()`` then stops here with the
            # gradients of everything downstream complete, and ``backward_encoders()`` runs the encoders' backward as
            # a second step -- the seam where a data-parallel step starts reducing the downstream gradients
            # (deeplio_b200.graph.GraphedTrainStep, deeplio_b200.parallel.OverlappedGradReducer)
            cut = y.detach().requires_grad_(True)
            self._cut = (y, cut)
            y = cut
        y = self._tail(y)
        return y.view(b, s, -1)

    def backward_encoders(self):
        """Second half of a split backward pass (``split_backward = True``): the encoders' backward from the gradient
        that ``loss.backward()`` left at the cut."""
        y, cut = self._cut
        self._cut = None
        if cut.grad is not None:
            y.backward(cut.grad)

    def _tail(self, y):
        if self.tail == "drop_then_leaky":      # Simple-1 (lidar_feat_nets.py:230-233)
            return Fn.linear(Fn.dropout(y, self.p, self.training), self.fc1.weight, self.fc1.bias, "leaky_relu")
        if self.tail == "drop_then_relu":       # ResNet (lidar_feat_nets.py:182-185)
            return Fn.linear(Fn.dropout(y, self.p, self.training), self.fc1.weight, self.fc1.bias, "relu")
        return Fn.dropout(Fn.linear(y, self.fc1.weight, self.fc1.bias, "relu"), self.p, self.training)


class LidarSimpleFeat1(BaseLidarFeatNet):
    tail = "drop_then_leaky"

    def __init__(self, input_shape, cfg):
        super().__init__(input_shape, cfg)
        c = input_shape[0]
        self.encoder1 = Simple1Encoder(2 * c, cfg["bypass"])
        self.encoder2 = Simple1Encoder(2 * c, cfg["bypass"])
        self._finish(512)


class LidarFlowNetFeat(BaseLidarFeatNet):
    def __init__(self, input_shape, cfg):
        super().__init__(input_shape, cfg)
        c = input_shape[0]
        self.encoder1 = FlowNetEncoder(2 * c)
        self.encoder2 = FlowNetEncoder(2 * c)
        self._finish(1024)


class LidarResNetFeat(BaseLidarFeatNet):
    tail = "drop_then_relu"

    def __init__(self, input_shape, cfg):
        super().__init__(input_shape, cfg)
        c = input_shape[0]
        self.encoder1 = ResNetEncoder(2 * c)
        self.encoder2 = ResNetEncoder(2 * c)
        self._finish(512)


class LidarPointSegFeat(BaseLidarFeatNet):
    def __init__(self, input_shape, cfg, bn_d=0.1):
        super().__init__(input_shape, cfg)
        self.part = cfg["part"].lower()
        self.bn_d = bn_d
        c = input_shape[0]
        self.encoder1 = PointSegEncoder(2 * c, cfg.get("bypass"), bn_d)
        self.encoder2 = PointSegEncoder(2 * c, cfg.get("bypass"), bn_d)
        self._finish(768)
