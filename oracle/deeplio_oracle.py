"""CPU fp32 oracle for the DeepLIO hot path (functional restatement).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``deeplio_b200``) never does.

This file restates, as pure functions over a flat ``state_dict`` that uses the
reference's key names, what the four nn.Module subsystems under
``/root/reference/deeplio/models/nets`` compute.  It is written from the
algorithm, not from the module classes: there are no nn.Modules here, only
``torch.nn.functional`` primitives (conv2d / batch_norm / max_pool2d / linear)
and hand-written LSTM / GRU recurrences.  Gradients come from torch autograd
over these functions.

Pinning: ``tests/test_oracle_vs_reference.py`` runs this file against the
unmodified reference modules imported from ``/root/reference`` (container only)
and ``tests/test_oracle_golden.py`` checks it against the committed golden
vectors in ``tests/golden/`` that ``oracle/make_golden.py`` generated from the
reference itself.  The reference has no tests / golden vectors of its own
(SURVEY.md section 4), so those reference-generated fixtures are the pin.

Documented deviations from the reference:
  P3  lidar ``fusion: cat`` -- the reference hard-codes fc1's in-features to one
      branch (lidar_feat_nets.py:66,116,161,204) and crashes; here fc1 is taken
      from the state_dict, so a ``[128, 2C]`` weight works.
  P4  soft fusion -- the reference multiplies in place (fusion_nets.py:72-73),
      which breaks autograd for ResNet/FlowNet; same values out of place here.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1  # nn.BatchNorm2d default; PointSeg passes bn_d=0.1 explicitly


# --------------------------------------------------------------------------- helpers
def _bn(sd, p, x, training):
    """nn.BatchNorm2d (train: batch mean / biased var; running stats momentum 0.1)."""
    rm, rv = sd[p + ".running_mean"], sd[p + ".running_var"]
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training and (p + ".num_batches_tracked") in sd:
        sd[p + ".num_batches_tracked"] += 1
    return y


# When a dict, every ReLU of the LiDAR encoders stores its mask (input > 0, bool NCHW) here under the name of the
# convolution that feeds it (tests count ReLU-mask disagreements between the B200 path and this oracle:
# tests/test_gpu_model.py, tests/test_gpu_fullsize.py)
TRACE = None
# When a dict (same keys, bool masks), the listed ReLUs apply THAT mask instead of their own input > 0: the gradient
# tests impose the B200 path's ReLU decisions on the fp64 oracle, so that the comparison is about arithmetic and not
# about which side of zero a 1e-7 input landed on (a parameter gradient is a sum of up to 10^8 terms of random sign:
# one flipped decision moves it by ~1 / sqrt(terms), far above the 2e-4 parity bar)
FORCE_MASKS = None


def _relu(x, key):
    if TRACE is not None:
        TRACE[key] = x.detach() > 0
    if FORCE_MASKS is not None and key in FORCE_MASKS:
        return x * FORCE_MASKS[key].to(x.dtype)
    return F.relu(x)


def _pool(x, stride, ceil_mode, key):
    """MaxPool2d(3, stride, padding=1).  With TRACE: stores the window-relative arg-max (uint8, dy * 3 + dx, NCHW) under
    ``key``; with FORCE_MASKS[key]: takes the window element THAT index names instead of the maximum (the B200
    path's arg-max decisions imposed on the oracle, as for the ReLU masks)."""
    if TRACE is None and (FORCE_MASKS is None or key not in FORCE_MASKS):
        return F.max_pool2d(x, 3, stride, 1, ceil_mode=ceil_mode)
    sh, sw = stride
    n, c, h, w = x.shape
    y, flat = F.max_pool2d(x, 3, stride, 1, ceil_mode=ceil_mode, return_indices=True)
    oh, ow = y.shape[2:]
    top = (torch.arange(oh) * sh - 1).view(1, 1, oh, 1)
    left = (torch.arange(ow) * sw - 1).view(1, 1, 1, ow)
    if TRACE is not None:
        TRACE[key] = (((flat // w) - top) * 3 + ((flat % w) - left)).to(torch.uint8)
    if FORCE_MASKS is not None and key in FORCE_MASKS:
        r = FORCE_MASKS[key].long()
        iy, ix = top + r // 3, left + r % 3
        return x.flatten(2).gather(2, (iy * w + ix).flatten(2)).view(n, c, oh, ow)
    return y


def _conv(sd, p, x, stride=(1, 1)):
    w = sd[p + ".weight"]
    pad = ((w.shape[2] - 1) // 2, (w.shape[3] - 1) // 2)
    return F.conv2d(x, w, sd.get(p + ".bias"), stride=stride, padding=pad)


def _drop(x, p, training):
    return F.dropout(x, p, training) if (p > 0.0 and training) else x


# --------------------------------------------------------------------------- LiDAR encoders
def simple1_encoder(sd, p, x, training, bypass=False):
    """FeatureNetSimple1 (lidar_feat_nets.py:279-342): conv -> ReLU -> BN, ceil-mode pools."""
    def blk(i, t, stride=(1, 1)):
        return _bn(sd, "%sbn%d" % (p, i), _relu(_conv(sd, "%sconv%d" % (p, i), t, stride), "%sconv%d" % (p, i)), training)

    def pool(t, stride, i):
        return _pool(t, stride, True, "%sconv%d#pool" % (p, i))

    t = pool(blk(1, x, (1, 2)), (1, 2), 1)
    t = pool(blk(2, t), (1, 2), 2)
    t = blk(3, t)
    ident = t
    t = blk(4, t)
    if bypass:
        t = t + ident
    t = pool(t, (2, 2), 4)
    t = blk(5, t)
    ident = t
    t = blk(6, t)
    if bypass:
        t = t + ident
    t = pool(t, (2, 2), 6)
    t = blk(7, t)
    return t.mean(dim=(2, 3))


_FLOWNET = [("conv1", (1, 2)), ("conv2", (1, 2)), ("conv3", (1, 2)), ("conv3_1", (1, 1)),
            ("conv4", (2, 2)), ("conv4_1", (1, 1)), ("conv5", (2, 2)), ("conv5_1", (1, 1)),
            ("conv6", (2, 2))]


def flownet_encoder(sd, p, x, training):
    """FlowNetEncoder (lidar_feat_nets.py:248-267) with base_net.conv (base_net.py:55-71):
    conv(no bias) -> BN -> ReLU, nine times, then global mean."""
    t = x
    for name, stride in _FLOWNET:
        t = _relu(_bn(sd, "%s%s.1" % (p, name), _conv(sd, "%s%s.0" % (p, name), t, stride), training), "%s%s.0" % (p, name))
    return t.mean(dim=(2, 3))


_RESNET_LAYERS = [("layer1", 3, (1, 2)), ("layer2", 3, (1, 2)), ("layer3", 3, (2, 2)), ("layer4", 2, (2, 2))]


def resnet_encoder(sd, p, x, training):
    """ResNetEncoder (resnet.py:36-108) over torchvision BasicBlock
    (torchvision/models/resnet.py:59-105, v0.26.0)."""
    t = _relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x), training), p + "conv1")
    t = _pool(t, (1, 2), False, p + "conv1#pool")
    for lname, nblk, stride in _RESNET_LAYERS:
        for b in range(nblk):
            q = "%s%s.%d." % (p, lname, b)
            s = stride if b == 0 else (1, 1)
            o = _relu(_bn(sd, q + "bn1", _conv(sd, q + "conv1", t, s), training), q + "conv1")
            o = _bn(sd, q + "bn2", _conv(sd, q + "conv2", o), training)
            if (q + "downsample.0.weight") in sd:
                t = _bn(sd, q + "downsample.1", _conv(sd, q + "downsample.0", t, s), training)
            t = _relu(o + t, q + "conv2")
    return t.mean(dim=(2, 3))


def _fire(sd, q, x, training, bypass):
    """Fire (pointseg_modules.py:86-142): squeeze1x1 -> {expand1x1 || expand3x3} -> cat (+x)."""
    s = _relu(_bn(sd, q + "squeeze_bn", _conv(sd, q + "squeeze", x), training), q + "squeeze")
    e1 = _relu(_bn(sd, q + "expand1x1_bn", _conv(sd, q + "expand1x1", s), training), q + "expand1x1")
    e3 = _relu(_bn(sd, q + "expand3x3_bn", _conv(sd, q + "expand3x3", s), training), q + "expand3x3")
    out = torch.cat([e1, e3], 1)
    if bypass == "simple" and out.shape[1] == x.shape[1]:
        out = out + x
    return out


def _se(sd, q, x):
    """SELayer (pointseg_modules.py:203-221), reduction 2, no biases."""
    y = x.mean(dim=(2, 3))
    y = torch.sigmoid(F.linear(F.relu(F.linear(y, sd[q + "fc.0.weight"])), sd[q + "fc.2.weight"]))
    return x * y[:, :, None, None]


# (block name, [entries]) -- entries: 'F' fire, 'FN' fire without bypass, 'S' SE, 'P12'/'P22' pools
_POINTSEG = [("fire_blk1", ["F", "F", "S", "P12"]), ("fire_blk2", ["F", "F", "S", "P12"]),
             ("fire_blk3", ["F", "F", "F", "F", "S", "P22"]), ("fire_blk4", ["F", "F", "S", "P22"]),
             ("fire_blk5", ["F", "FN"])]


def pointseg_encoder(sd, p, x, training, bypass="simple"):
    """PSEncoder (pointseg_net.py:18-71)."""
    t = _relu(_bn(sd, p + "conv1a.1", _conv(sd, p + "conv1a.0", x, (1, 2)), training), p + "conv1a.0")
    t = _pool(t, (1, 2), False, p + "conv1a.0#pool")
    for bname, entries in _POINTSEG:
        for i, e in enumerate(entries):
            q = "%s%s.%d." % (p, bname, i)
            if e == "F":
                t = _fire(sd, q, t, training, bypass)
            elif e == "FN":
                t = _fire(sd, q, t, training, None)
            elif e == "S":
                t = _se(sd, q, t)
            elif e == "P12":
                t = _pool(t, (1, 2), False, q + "pool")
            else:
                t = _pool(t, (2, 2), False, q + "pool")
    return t.mean(dim=(2, 3))  # adaptive_avg_pool2d outside the encoder (lidar_feat_nets.py:84-85)


def lidar_feat(sd, cfg, name, xyz, normals, training, p="lidar_feat_net."):
    """Lidar*Feat.forward (lidar_feat_nets.py:73-101,119-148,165-189,208-237)."""
    c = cfg[name]
    b, s, t, ch, h, w = xyz.shape
    x0 = xyz.reshape(b * s, t * ch, h, w)
    x1 = normals.reshape(b * s, t * ch, h, w)
    if name == "lidar-feat-simple-1":
        f0 = simple1_encoder(sd, p + "encoder1.", x0, training, c.get("bypass", False))
        f1 = simple1_encoder(sd, p + "encoder2.", x1, training, c.get("bypass", False))
    elif name == "lidar-feat-flownet":
        f0 = flownet_encoder(sd, p + "encoder1.", x0, training)
        f1 = flownet_encoder(sd, p + "encoder2.", x1, training)
    elif name == "lidar-feat-resnet":
        f0 = resnet_encoder(sd, p + "encoder1.", x0, training)
        f1 = resnet_encoder(sd, p + "encoder2.", x1, training)
    elif name == "lidar-feat-pointseg":
        f0 = pointseg_encoder(sd, p + "encoder1.", x0, training, c.get("bypass"))
        f1 = pointseg_encoder(sd, p + "encoder2.", x1, training, c.get("bypass"))
    else:
        raise ValueError("Wrong feature network {}".format(name))
    fusion = c["fusion"]
    if fusion == "cat":
        y = torch.cat((f0, f1), dim=1)
    elif fusion == "add":
        y = f0 + f1
    else:
        y = f0 - f1
    pd = float(c["dropout"])
    w1, b1 = sd[p + "fc1.weight"], sd[p + "fc1.bias"]
    if name == "lidar-feat-simple-1":
        y = F.leaky_relu(F.linear(_drop(y, pd, training), w1, b1), 0.01)
    elif name == "lidar-feat-resnet":
        y = F.relu(F.linear(_drop(y, pd, training), w1, b1))
    else:  # pointseg / flownet: drop(relu(fc1(x)))
        y = _drop(F.relu(F.linear(y, w1, b1)), pd, training)
    return y.view(b, s, -1)


# --------------------------------------------------------------------------- recurrent cells
def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, h0, c0, reverse):
    """One direction of one LSTM layer.  x [B,T,I]; gate order i,f,g,o (torch.nn.LSTM)."""
    T = x.shape[1]
    gx = F.linear(x, w_ih, b_ih)
    h, c = h0, c0
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = gx[:, t] + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1), h, c


def _gru_dir(x, w_ih, w_hh, b_ih, b_hh, h0, reverse):
    """One direction of one GRU layer; gate order r,z,n (torch.nn.GRU)."""
    T = x.shape[1]
    gx = F.linear(x, w_ih, b_ih)
    h = h0
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = F.linear(h, w_hh, b_hh)
        xr, xz, xn = gx[:, t].chunk(3, dim=1)
        hr, hz, hn = gh.chunk(3, dim=1)
        r = torch.sigmoid(xr + hr)
        z = torch.sigmoid(xz + hz)
        n = torch.tanh(xn + r * hn)
        h = (1.0 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1), h


def rnn_stack(sd, p, x, kind, num_layers, bidirectional, state, dropout, training):
    """Multi-layer (bi)LSTM/GRU, batch_first.  ``state`` is None or (h[L*D,B,H], c[L*D,B,H]) /
    h[L*D,B,H]; returns (out [B,T,D*H], new_state) exactly like torch.nn.LSTM / GRU."""
    D = 2 if bidirectional else 1
    B = x.shape[0]
    H = sd[p + "weight_hh_l0"].shape[1]
    inp = x
    hs, cs = [], []
    for l in range(num_layers):
        outs = []
        for d in range(D):
            sfx = "_l%d%s" % (l, "_reverse" if d == 1 else "")
            w_ih, w_hh = sd[p + "weight_ih" + sfx], sd[p + "weight_hh" + sfx]
            b_ih, b_hh = sd[p + "bias_ih" + sfx], sd[p + "bias_hh" + sfx]
            k = l * D + d
            if kind == "lstm":
                if state is None:
                    h0 = x.new_zeros(B, H)
                    c0 = x.new_zeros(B, H)
                else:
                    h0, c0 = state[0][k], state[1][k]
                o, h, c = _lstm_dir(inp, w_ih, w_hh, b_ih, b_hh, h0, c0, d == 1)
                cs.append(c)
            else:
                h0 = x.new_zeros(B, H) if state is None else state[k]
                o, h = _gru_dir(inp, w_ih, w_hh, b_ih, b_hh, h0, d == 1)
            hs.append(h)
            outs.append(o)
        inp = torch.cat(outs, dim=2) if D == 2 else outs[0]
        if l < num_layers - 1:
            inp = _drop(inp, dropout, training)
    new_state = (torch.stack(hs), torch.stack(cs)) if kind == "lstm" else torch.stack(hs)
    return inp, new_state


# --------------------------------------------------------------------------- IMU / fusion / odometry
def imu_feat(sd, cfg, name, imus, training, p="imu_feat_net."):
    c = cfg[name]
    b, s, t, n = imus.shape
    if name == "imu-feat-fc":
        # ImuFeatFC.forward (imu_feat_nets.py:38-53): per (b,s) MLP over the T samples, summed over T.
        y = imus.reshape(b * s * t, n)
        for i in range(len(c["hidden-size"])):
            y = F.leaky_relu(F.linear(y, sd["%snet.%d.weight" % (p, i)], sd["%snet.%d.bias" % (p, i)]), 0.01)
        y = _drop(y, float(c["dropout"]), training)
        return y.view(b, s, t, -1).sum(dim=2)
    if name == "imu-feat-rnn":
        # ImufeatRNN0.forward (imu_feat_nets.py:75-83): state carried from window s to s+1;
        # feature = last-layer forward-direction hidden at the last time step.
        kind = c["type"].lower()
        kind = "gru" if kind == "gru" else "lstm"
        H = c.get("hidden-size", 6)
        bi = c.get("bidirectional", False)
        L = c.get("num-layers", 2)
        state = None
        feats = []
        for q in range(s):
            out, state = rnn_stack(sd, p + "rnn.", imus[:, q], kind, L, bi, state, float(c["dropout"]), training)
            feats.append(out[:, -1, :H])
        return torch.stack(feats, dim=1)
    raise ValueError("Wrong feature network {}".format(name))


def fusion_feat(sd, cfg, name, lidar, imu, p="fusion_net."):
    if name == "fusion-layer-cat":
        return torch.cat((lidar, imu), dim=2)  # fusion_nets.py:24-31
    if name == "fusion-layer-soft":
        # DeepLIOFusionSoft.forward (fusion_nets.py:64-75); out-of-place (patch P4)
        cat = torch.cat((lidar, imu), dim=2)
        s1 = torch.sigmoid(F.linear(cat, sd[p + "layers.0.weight"], sd[p + "layers.0.bias"]))
        s2 = torch.sigmoid(F.linear(cat, sd[p + "layers.1.weight"], sd[p + "layers.1.bias"]))
        return torch.cat((lidar * s1, imu * s2), dim=2)
    raise ValueError("Wrong feature network {}".format(name))


def odom_feat(sd, cfg, name, x, training, p="odom_feat_net."):
    c = cfg[name]
    if name == "odom-feat-fc":
        # OdomFeatFC (odom_feat_nets.py:8-45) reads cfg['hidden-size'] (the yaml key is 'size'),
        # so the effective widths come from the state_dict (default [256,128]).
        b, s, n = x.shape
        y = x.reshape(b * s, n)
        i = 0
        while ("%slayers.%d.weight" % (p, i)) in sd:
            y = F.leaky_relu(F.linear(y, sd["%slayers.%d.weight" % (p, i)], sd["%slayers.%d.bias" % (p, i)]), 0.01)
            i += 1
        y = _drop(y, float(c.get("dropout", 0.0)), training)
        return y.view(b, s, -1)
    if name == "odom-feat-rnn":
        # OdomFeatRNN.forward (odom_feat_nets.py:72-83): RNN over the S axis, forward direction kept.
        kind = "gru" if c["type"].lower() == "gru" else "lstm"
        H = c.get("hidden-size", 6)
        out, _ = rnn_stack(sd, p + "rnn.", x, kind, c.get("num-layers", 2), c.get("bidirectional", False),
                           None, float(c.get("dropout", 0.0)), training)
        return out[:, :, :H]
    raise ValueError("Wrong odometry feature network {}".format(name))


# --------------------------------------------------------------------------- whole model
def deeplio_forward(sd, cfg, xyz, normals, imus, training=True, return_feats=False):
    """DeepLIO.forward (deeplio_nets.py:60-90) behind get_model (nets/__init__.py:16-79).

    sd       flat state_dict with the reference key names (BN buffers are updated in place
             when ``training``).
    cfg      the parsed config.yaml dict.
    xyz, normals  [B,S,2,3,H,W];  imus [B,S,T,6].
    Dropout follows ``training`` and the configured p (parity runs use p = 0).
    """
    arch = cfg["deeplio"]
    names = {k: (arch[k].get("name") if arch.get(k) else None)
             for k in ("lidar-feat-net", "imu-feat-net", "fusion-net", "odom-feat-net")}
    feats = {}
    last = None
    fl = fi = None
    if names["lidar-feat-net"]:
        fl = lidar_feat(sd, cfg, names["lidar-feat-net"].lower(), xyz, normals, training)
        feats["lidar"] = last = fl
    if names["imu-feat-net"]:
        fi = imu_feat(sd, cfg, names["imu-feat-net"].lower(), imus, training)
        feats["imu"] = last = fi
    if fl is not None and fi is not None and names["fusion-net"]:
        last = fusion_feat(sd, cfg, names["fusion-net"].lower(), fl, fi)
        feats["fusion"] = last
    if names["odom-feat-net"]:
        last = odom_feat(sd, cfg, names["odom-feat-net"].lower(), last, training)
        feats["odom"] = last
    last = _drop(last, float(arch.get("dropout", 0.0)), training)
    pos = F.linear(last, sd["fc_pos.weight"], sd["fc_pos.bias"])
    ori = F.linear(last, sd["fc_ori.weight"], sd["fc_ori.bias"])
    if return_feats:
        return pos, ori, feats
    return pos, ori


# --------------------------------------------------------------------------- shapes / synthetic state
def conv_out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def pool_out(n, s, ceil_mode):
    """MaxPool2d(3, s, 1) output size (ceil_mode semantics of torch)."""
    if not ceil_mode:
        return (n + 2 - 3) // s + 1
    o = -(-(n + 2 - 3) // s) + 1
    if (o - 1) * s >= n + 1:
        o -= 1
    return o


def _rnn_shapes(p, kind, I, H, L, bi, out):
    G = 4 if kind == "lstm" else 3
    D = 2 if bi else 1
    for l in range(L):
        for d in range(D):
            sfx = "_l%d%s" % (l, "_reverse" if d else "")
            out[p + "weight_ih" + sfx] = (G * H, I if l == 0 else D * H)
            out[p + "weight_hh" + sfx] = (G * H, H)
            out[p + "bias_ih" + sfx] = (G * H,)
            out[p + "bias_hh" + sfx] = (G * H,)


def _bn_shapes(p, c, out):
    out[p + ".weight"] = (c,)
    out[p + ".bias"] = (c,)
    out[p + ".running_mean"] = (c,)
    out[p + ".running_var"] = (c,)
    out[p + ".num_batches_tracked"] = ()


def state_shapes(cfg):
    """Key -> shape for the model ``cfg`` selects (reference key order is not needed: values are
    drawn per key from a hash-seeded generator, see ``synthetic_state``)."""
    arch = cfg["deeplio"]
    out = {}
    lname = arch["lidar-feat-net"].get("name")
    feat = None
    if lname:
        lname = lname.lower()
        for e in ("lidar_feat_net.encoder1.", "lidar_feat_net.encoder2."):
            if lname == "lidar-feat-simple-1":
                spec = [(1, 6, 64, 5, 7), (2, 64, 128, 3, 5), (3, 128, 128, 3, 3), (4, 128, 256, 3, 3),
                        (5, 256, 256, 3, 3), (6, 256, 512, 3, 3), (7, 512, 512, 3, 3)]
                for i, ci, co, kh, kw in spec:
                    out["%sconv%d.weight" % (e, i)] = (co, ci, kh, kw)
                    out["%sconv%d.bias" % (e, i)] = (co,)
                    _bn_shapes("%sbn%d" % (e, i), co, out)
                width = 512
            elif lname == "lidar-feat-flownet":
                spec = [("conv1", 6, 64, 5, 7), ("conv2", 64, 128, 3, 5), ("conv3", 128, 256, 3, 5),
                        ("conv3_1", 256, 256, 3, 3), ("conv4", 256, 512, 3, 3), ("conv4_1", 512, 512, 3, 3),
                        ("conv5", 512, 512, 3, 3), ("conv5_1", 512, 512, 3, 3), ("conv6", 512, 1024, 3, 3)]
                for n, ci, co, kh, kw in spec:
                    out["%s%s.0.weight" % (e, n)] = (co, ci, kh, kw)
                    _bn_shapes("%s%s.1" % (e, n), co, out)
                width = 1024
            elif lname == "lidar-feat-resnet":
                out[e + "conv1.weight"] = (64, 6, 5, 7)
                out[e + "conv1.bias"] = (64,)
                _bn_shapes(e + "bn1", 64, out)
                inpl = 64
                for (ln, nb, _), planes in zip(_RESNET_LAYERS, (64, 128, 256, 512)):
                    for b in range(nb):
                        q = "%s%s.%d." % (e, ln, b)
                        out[q + "conv1.weight"] = (planes, inpl, 3, 3)
                        _bn_shapes(q + "bn1", planes, out)
                        out[q + "conv2.weight"] = (planes, planes, 3, 3)
                        _bn_shapes(q + "bn2", planes, out)
                        if b == 0:
                            out[q + "downsample.0.weight"] = (planes, inpl, 1, 1)
                            _bn_shapes(q + "downsample.1", planes, out)
                        inpl = planes
                width = 512
            elif lname == "lidar-feat-pointseg":
                out[e + "conv1a.0.weight"] = (64, 6, 3, 5)
                out[e + "conv1a.0.bias"] = (64,)
                _bn_shapes(e + "conv1a.1", 64, out)
                fires = {"fire_blk1": [(64, 16, 64), (128, 16, 64)], "fire_blk2": [(128, 32, 128), (256, 32, 128)],
                         "fire_blk3": [(256, 48, 192), (384, 48, 192), (384, 64, 256), (512, 64, 256)],
                         "fire_blk4": [(512, 64, 256), (512, 64, 256)], "fire_blk5": [(512, 80, 384), (768, 80, 384)]}
                se = {"fire_blk1": (2, 128), "fire_blk2": (2, 256), "fire_blk3": (4, 512), "fire_blk4": (2, 512)}
                for bn_, fl in fires.items():
                    for i, (ci, sq, ex) in enumerate(fl):
                        q = "%s%s.%d." % (e, bn_, i)
                        out[q + "squeeze.weight"] = (sq, ci, 1, 1)
                        out[q + "squeeze.bias"] = (sq,)
                        _bn_shapes(q + "squeeze_bn", sq, out)
                        out[q + "expand1x1.weight"] = (ex, sq, 1, 1)
                        out[q + "expand1x1.bias"] = (ex,)
                        _bn_shapes(q + "expand1x1_bn", ex, out)
                        out[q + "expand3x3.weight"] = (ex, sq, 3, 3)
                        out[q + "expand3x3.bias"] = (ex,)
                        _bn_shapes(q + "expand3x3_bn", ex, out)
                    if bn_ in se:
                        i, c = se[bn_]
                        out["%s%s.%d.fc.0.weight" % (e, bn_, i)] = (c // 2, c)
                        out["%s%s.%d.fc.2.weight" % (e, bn_, i)] = (c, c // 2)
                width = 768
            else:
                raise ValueError("Wrong feature network {}".format(lname))
        fin = width * (2 if cfg[lname]["fusion"] == "cat" else 1)
        out["lidar_feat_net.fc1.weight"] = (128, fin)
        out["lidar_feat_net.fc1.bias"] = (128,)
        feat = 128
    iname = arch["imu-feat-net"].get("name")
    ifeat = None
    if iname:
        iname = iname.lower()
        c = cfg[iname]
        if iname == "imu-feat-fc":
            prev = c["input-size"]
            for i, h in enumerate(c["hidden-size"]):
                out["imu_feat_net.net.%d.weight" % i] = (h, prev)
                out["imu_feat_net.net.%d.bias" % i] = (h,)
                prev = h
            ifeat = prev
        else:
            kind = "gru" if c["type"].lower() == "gru" else "lstm"
            _rnn_shapes("imu_feat_net.rnn.", kind, c["input-size"], c["hidden-size"], c.get("num-layers", 2),
                        c.get("bidirectional", False), out)
            ifeat = c["hidden-size"]
    fname = arch["fusion-net"].get("name") if arch.get("fusion-net") else None
    if feat is not None and ifeat is not None and fname:
        if fname.lower() == "fusion-layer-soft":
            out["fusion_net.layers.0.weight"] = (feat, feat + ifeat)
            out["fusion_net.layers.0.bias"] = (feat,)
            out["fusion_net.layers.1.weight"] = (ifeat, feat + ifeat)
            out["fusion_net.layers.1.bias"] = (ifeat,)
        feat = feat + ifeat
    elif feat is None:
        feat = ifeat
    oname = arch["odom-feat-net"].get("name")
    if oname:
        oname = oname.lower()
        c = cfg[oname]
        if oname == "odom-feat-fc":
            prev = feat
            for i, h in enumerate(c.get("hidden-size", [256, 128])):
                out["odom_feat_net.layers.%d.weight" % i] = (h, prev)
                out["odom_feat_net.layers.%d.bias" % i] = (h,)
                prev = h
            feat = prev
        else:
            kind = "gru" if c["type"].lower() == "gru" else "lstm"
            _rnn_shapes("odom_feat_net.rnn.", kind, feat, c["hidden-size"], c.get("num-layers", 2),
                        c.get("bidirectional", False), out)
            feat = c["hidden-size"]
    out["fc_pos.weight"] = (3, feat)
    out["fc_pos.bias"] = (3,)
    out["fc_ori.weight"] = (3, feat)
    out["fc_ori.bias"] = (3,)
    return out


def _key_seed(seed, key):
    h = 1469598103934665603
    for ch in ("%d|%s" % (seed, key)).encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFFFFFFFFFF


def synthetic_state(cfg, seed=0):
    """Deterministic weights for parity runs, independent of module construction order:
    every tensor is drawn from its own generator seeded by hash(seed, key).
    weights ~ U(+-sqrt(3/fan_in)) (unit-gain), biases U(+-0.1), BN gamma U(0.5,1.5), beta U(+-0.2),
    running_mean U(+-0.1), running_var U(0.5,1.5)."""
    shapes = state_shapes(cfg)
    bn_prefixes = {k[: -len("running_mean")] for k in shapes if k.endswith(".running_mean")}
    sd = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed(_key_seed(seed, key))
        prefix, leaf = key.rsplit(".", 1)
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros((), dtype=torch.long)
            continue
        u = torch.rand(shape, generator=g, dtype=torch.float32)
        if (prefix + ".") in bn_prefixes:
            if leaf in ("weight", "running_var"):
                sd[key] = 0.5 + u
            elif leaf == "bias":
                sd[key] = (u - 0.5) * 0.4
            else:
                sd[key] = (u - 0.5) * 0.2
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[key] = (u * 2 - 1) * math.sqrt(3.0 / fan_in)
        else:
            sd[key] = (u * 2 - 1) * 0.1
    return sd


IMG_STD = (0.1269, 0.0951, 0.0108, 0.3436, 0.4445, 0.5664)  # config.yaml:26-27 (x,y,z,nx,ny,nz)


def synthetic_batch(B, S, H, W, T_imu, seed=0, zero_frac=0.15):
    """Synthetic frame-pair batch: xyz / normals [B,S,2,3,H,W] with the dataset's
    per-channel std (mean-subtracted => zero mean), a fraction of empty returns, IMU ~ N(0,1)."""
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    frames = torch.randn(B, S + 1, 6, H, W, generator=g)
    frames *= torch.tensor(IMG_STD).view(1, 1, 6, 1, 1)
    if zero_frac > 0:
        keep = (torch.rand(B, S + 1, 1, H, W, generator=g) >= zero_frac).float()
        frames *= keep
    idx = torch.tensor([[i, i + 1] for i in range(S)])
    pairs = frames[:, idx]  # [B,S,2,6,H,W]   (misc.py:65-69)
    xyz = pairs[:, :, :, 0:3]
    normals = pairs[:, :, :, 3:].contiguous()
    imus = torch.randn(B, S, T_imu, 6, generator=g)
    return xyz, normals, imus
