"""Caller-side glue on the device: pose chaining, ground-truth pairing, the full HWS / LWS
losses, the fused finite check and the pairing gather -- against tests/golden/pose_glue.pt (outputs of the reference's
own Trainer.se3_to_SE3, DataCombiCreater.process_ground_turth, HWSLoss / LWSLoss, oracle/make_golden_pose.py) and
the oracle (oracle/pose_oracle.py; liegroups' SO(3) maps are restated there: parity unpinned for that class)."""
import argparse
import os

import pytest
import torch

from oracle import deeplio_oracle as O
from oracle import pose_oracle as P
from oracle.configs import make_cfg
from tests.helpers import GOLDEN_DIR, quat_tol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def rec():
    return torch.load(os.path.join(GOLDEN_DIR, "pose_glue.pt"), weights_only=False)


def test_se3_chain_forward_and_backward(rec):
    from deeplio_b200 import pose
    c = rec["chain"]
    x, w = c["x"].to(DEV).requires_grad_(True), c["w"].to(DEV).requires_grad_(True)
    fx, fq = pose.se3_to_SE3(x, w)          # check=True: reads the status word, raises like the reference
    assert torch.allclose(fx.detach().cpu(), c["f2g_x"], rtol=1e-6, atol=2e-6)
    assert ((fq.detach().cpu() - c["f2g_q"]).abs() <= quat_tol(c["f2g_q"])).all()
    ((fx * c["gx"].to(DEV)).sum() + (fq * c["gq"].to(DEV)).sum()).backward()
    well = (c["f2g_q"][:, :, 0].abs().min(dim=1).values > 0.05)
    assert torch.allclose(x.grad.cpu()[well], c["dx"][well], rtol=1e-5, atol=1e-5)
    assert torch.allclose(w.grad.cpu()[well], c["dw"][well], rtol=1e-4, atol=1e-4)
    # every sample (incl. the near-half-turn ones) against the oracle in fp64 at a conditioning-aware bar
    x64, w64 = c["x"].double().requires_grad_(True), c["w"].double().requires_grad_(True)
    ox, oq, _ = P.se3_to_SE3(x64, w64)
    ((ox * c["gx"].double()).sum() + (oq * c["gq"].double()).sum()).backward()
    amp = 1.0 / (4.0 * c["f2g_q"][:, :, 0].abs().clamp_min(1e-3)).min(dim=1).values ** 2
    for b in range(x.shape[0]):
        bar = 2e-5 * (1.0 + float(amp[b])) * (1.0 + float(w64.grad[b].abs().max()))
        assert (w.grad[b].cpu().double() - w64.grad[b]).abs().max().item() <= bar, b
        assert (x.grad[b].cpu().double() - x64.grad[b]).abs().max().item() <= 1e-5 * (1.0 + float(x64.grad[b].abs().max()))


def test_se3_chain_status_flags():
    from deeplio_b200 import pose
    x = torch.zeros(2, 3, 3, device=DEV)
    w = torch.zeros(2, 3, 3, device=DEV)
    w[1, 1, 0] = float("inf")
    with pytest.raises(ValueError):
        pose.se3_to_SE3(x, w)
    assert int(pose.status_word(DEV).item()) == 0          # cleared by the raise
    fx, fq = pose.se3_to_SE3(x, torch.zeros_like(w))
    assert torch.equal(fq.cpu(), torch.tensor([1.0, 0, 0, 0]).expand(2, 3, 4))
    assert float(fx.abs().max()) == 0.0


def test_ground_truth_pairing(rec):
    from deeplio_b200 import data, pose
    g = rec["gt"]
    f2f, f2g = data.ground_truth(g["gts"].to(DEV), g["combinations"])
    assert pose.raise_for_status(DEV) == 0
    tmax = g["gts"][:, :, 0:3].abs().max().item()
    # translations: the reference forms R^T t_j - R^T t_i in fp32, the kernel R^T (t_j - t_i)
    assert (f2f.cpu()[:, :, 0:3] - g["f2f"][:, :, 0:3]).abs().max().item() <= 4e-7 * tmax + 1e-6
    assert (f2g.cpu()[:, :, 0:3] - g["f2g"][:, :, 0:3]).abs().max().item() <= 4e-7 * tmax + 1e-6
    assert torch.allclose(f2f.cpu()[:, :, 3:], g["f2f"][:, :, 3:], rtol=1e-5, atol=2e-6)
    assert torch.allclose(f2g.cpu()[:, :, 3:], g["f2g"][:, :, 3:], rtol=1e-5, atol=2e-6)
    bad = g["gts"].clone()
    bad[1, 2, 3:12] *= 1.01                                # no longer a rotation: the reference raises (misc.py:104)
    data.ground_truth(bad.to(DEV), g["combinations"])
    with pytest.raises(ValueError):
        pose.raise_for_status(DEV)


@pytest.mark.parametrize("name", ["hws_both", "hws_local", "hws_global", "lws_both"])
def test_losses_as_the_trainer_calls_them(rec, name):
    """trainer.py:246-263: chain the predictions, detach one side for one-sided losses, slice the global terms to
    1 .. max_glob_seq, call the criterion -- with deeplio_b200.pose / deeplio_b200.losses in place of the reference's."""
    from deeplio_b200 import losses, pose
    lo, case = rec["loss"], rec["loss"]["cases"][name]
    lt = case["loss_types"]
    if "sx" in case:
        crit = losses.HWSLoss(sx=float(case["sx"]), sq=float(case["sq"]), learn_hyper_params=True, device=DEV, loss_Types=lt)
    else:
        crit = losses.LWSLoss(beta=case["beta"], loss_Types=lt)
    pt, pw = lo["pred_t"].to(DEV).requires_grad_(True), lo["pred_w"].to(DEV).requires_grad_(True)
    p, q = pose.se3_to_SE3(pt, pw)
    a, b = pt, pw
    if lt[0] and not lt[1]:
        p, q = p.detach(), q.detach()
    elif lt[1] and not lt[0]:
        a, b = pt.detach(), pw.detach()
    gf, gg = lo["gt_f2f"].to(DEV), lo["gt_f2g"].to(DEV)
    g0, g1 = lo["g0"], lo["g1"]
    loss = crit(a, b, p[:, g0:g1, :], q[:, g0:g1, :], gf[:, :, 0:3], gf[:, :, 3:], gg[:, g0:g1, 0:3], gg[:, g0:g1, 3:7])
    (2.0 * loss).backward()                                 # a non-trivial upstream gradient
    assert torch.allclose(loss.detach().cpu(), case["loss"], rtol=2e-6, atol=1e-6)
    dt = pt.grad.cpu() if pt.grad is not None else torch.zeros_like(lo["pred_t"])
    dw = pw.grad.cpu() if pw.grad is not None else torch.zeros_like(lo["pred_w"])
    assert torch.allclose(dt, 2.0 * case["dt"], rtol=1e-5, atol=1e-6 * (1 + case["dt"].abs().max().item()))
    assert torch.allclose(dw, 2.0 * case["dw"], rtol=1e-4, atol=1e-5 * (1 + case["dw"].abs().max().item()))
    if "sx" in case:
        assert torch.allclose(crit.sx.grad.cpu(), 2.0 * case["dsx"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(crit.sq.grad.cpu(), 2.0 * case["dsq"], rtol=1e-5, atol=1e-6)
        assert [n for n, _ in crit.named_parameters()] == ["sx", "sq"]


def test_loss_factory_mirrors_the_reference():
    from deeplio_b200 import losses
    cfg = make_cfg(height=16, width=64)
    crit = losses.get_loss_function(cfg, DEV)
    assert isinstance(crit, losses.HWSLoss) and crit.loss_Types == [True, True]
    assert float(crit.sq) == -3.0 and crit.sx.requires_grad
    cfg["losses"]["active"] = "lwsloss"
    cfg["losses"]["loss-type"] = "local"
    crit = losses.get_loss_function(cfg, DEV)
    assert isinstance(crit, losses.LWSLoss) and crit.beta == 1125.0 and crit.loss_Types == [True, False]
    cfg["losses"]["loss-type"] = "sideways"
    with pytest.raises(ValueError):
        losses.get_loss_function(cfg, DEV)
    cfg["losses"]["loss-type"] = "global"
    cfg["losses"]["active"] = "nope"
    with pytest.raises(ValueError):
        losses.get_loss_function(cfg, DEV)


def test_fused_finite_check():
    from deeplio_b200 import pose
    g = torch.Generator().manual_seed(0)
    a = torch.randn(3, 1001, generator=g).to(DEV)            # odd size: scalar tail
    b = torch.randn(1 << 20, generator=g).to(DEV)
    c = torch.randn(7, generator=g).to(DEV)[1:]              # unaligned pointer
    assert int(pose.check_finite([("a", a), ("b", b), ("c", c)]).item()) == 0
    b[123457] = float("nan")
    with pytest.raises(ValueError, match="b"):
        pose.check_finite([("a", a), ("b", b), ("c", c)])
    c[5] = float("-inf")
    flags = pose.check_finite([("a", a), ("b", b), ("c", c)], check=False)
    assert int(flags.item()) == 0b110


def test_paired_frames_equal_the_materialised_pairs():
    """DataCombiCreater on the device: the model fed PairedFrames handles (frames paired inside dlio_pair_gather)
    gives bit-identical outputs and gradients to the model fed the reference's [B,S,2,C,H,W] tensors."""
    from deeplio_b200 import data, nets
    from deeplio_b200.config import build_config_container
    B, S, H, W, T = 2, 3, 16, 64, 5
    cfg = make_cfg(height=H, width=W, seq=S, odom_hidden=32)
    build_config_container(cfg, argparse.Namespace(device=DEV, batch_size=B))
    model = nets.get_model((3, H, W), cfg, DEV)
    model.load_state_dict(O.synthetic_state(cfg, seed=2))
    model.train()
    g = torch.Generator().manual_seed(3)
    frames = torch.randn(B, S + 1, 6, H, W, generator=g)
    batch = {"images": frames, "untrans-images": frames.clone(), "imus": torch.randn(B, S, T, 6, generator=g),
             "gts": P.synthetic_gts(B, S + 1, seed=4)}
    creater = data.DataCombiCreater(cfg["datasets"]["combinations"], device=DEV, check=True)
    creater(batch)
    assert tuple(creater.res_imgs.shape) == (B, S, 2, 3, H, W) and tuple(creater.res_gt_f2f.shape) == (B, S, 6)
    f2f, f2g = P.ground_truth(batch["gts"], cfg["datasets"]["combinations"])
    assert torch.allclose(creater.res_gt_f2f.cpu()[:, :, 3:], f2f[:, :, 3:], rtol=1e-5, atol=2e-6)
    assert torch.allclose(creater.res_gt_f2g.cpu()[:, :, 3:], f2g[:, :, 3:], rtol=1e-5, atol=2e-6)
    from deeplio_b200 import engine as E
    idx = torch.tensor(cfg["datasets"]["combinations"])
    pairs = frames[:, idx].to(DEV)                           # misc.py:65-69
    assert torch.equal(creater.res_normals.materialize(), pairs[:, :, :, 3:])

    def step(inputs):
        model.zero_grad(set_to_none=True)
        pos, ori = model(inputs)
        ((pos ** 2).sum() + (ori ** 2).sum()).backward()
        return pos.detach(), ori.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}
    materialised = [[pairs[:, :, :, 0:3], pairs[:, :, :, 3:].contiguous()], batch["imus"].to(DEV)]
    handles = [[creater.res_imgs, creater.res_normals], creater.res_imu]
    pos2, ori2, grads2 = step(materialised)
    # (a) same first-layer kernels on both sides (3xTF32 from fp32 planes): the gather itself changes nothing
    first_f16, E.FIRST_F16 = E.FIRST_F16, False
    try:
        pos, ori, grads = step(handles)
    finally:
        E.FIRST_F16 = first_f16
    assert torch.equal(pos, pos2) and torch.equal(ori, ori2)
    for k in grads:
        scale = grads2[k].abs().max().item() + 1e-12
        assert (grads[k] - grads2[k]).abs().max().item() <= 1e-4 * scale, k      # atomics order only
    # (b) the default: PairedFrames feed the first layer packed fp16 planes (folded-split operands) -- other
    # kernels, same fp32-level arithmetic
    pos, ori, grads = step(handles)
    assert (pos - pos2).abs().max().item() <= 2e-5 * pos2.abs().max().item()
    assert (ori - ori2).abs().max().item() <= 2e-5 * ori2.abs().max().item()
    for k in grads:
        scale = grads2[k].abs().max().item() + 1e-12
        # two different kernels for the first layer: a ReLU / arg-max decision within round-off of a tie may fall the
        # other way, and in these 16 x 64 images one flip moves a gradient by ~1 / sqrt(terms) ~ 1e-2 (the arithmetic
        # itself is held to 2e-4 at imposed decisions by tests/test_gpu_model.py / test_gpu_fullsize.py)
        assert (grads[k] - grads2[k]).abs().max().item() <= 3e-2 * scale, k
    # a NaN in the frames is reported by the gather itself
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    fr = frames.clone()
    fr[1, 2, 4, 3, 7] = float("nan")
    data.pair_gather(torch.device(DEV), data.PairedFrames(fr.to(DEV), cfg["datasets"]["combinations"], 3, 3), 8, 2, 4,
                     True, flags, 4)
    assert int(flags.item()) == 4


def test_cat_and_stack_through_the_strided_copy_equal_torch():
    """functional.cat_last / stack_mid (dlio_copy2d) against torch.cat / torch.stack, values and gradients; the stack
    takes the row-strided views the IMU net hands it (out[:, -1, :H] of a [B, T, 2H] tensor)."""
    from deeplio_b200 import functional as Fn
    g = torch.Generator().manual_seed(9)
    a = torch.randn(3, 4, 5, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(3, 4, 7, generator=g).to(DEV).requires_grad_(True)
    w = torch.randn(3, 4, 12, generator=g).to(DEV)
    (Fn.cat_last(a, b) * w).sum().backward()
    ga, gb = a.grad.clone(), b.grad.clone()
    a.grad = b.grad = None
    ref = torch.cat((a, b), dim=2)
    assert torch.equal(Fn.cat_last(a, b), ref)
    (ref * w).sum().backward()
    assert torch.equal(ga, a.grad) and torch.equal(gb, b.grad)
    outs = [torch.randn(4, 6, 10, generator=g).to(DEV).requires_grad_(True) for _ in range(3)]
    feats = [o[:, -1, :5] for o in outs]
    ws = torch.randn(4, 3, 5, generator=g).to(DEV)
    got = Fn.stack_mid(feats)
    assert torch.equal(got, torch.stack(feats, dim=1))
    (got * ws).sum().backward()
    grads = [o.grad.clone() for o in outs]
    for o in outs:
        o.grad = None
    (torch.stack([o[:, -1, :5] for o in outs], dim=1) * ws).sum().backward()
    assert all(torch.equal(x, o.grad) for x, o in zip(grads, outs))


def test_pair_gather_fp16_planes_equal_packing_the_materialised_pairs():
    """dlio_pair_gather(dst_h2): the packed fp16 planes of the first layer, written straight from the un-paired frames,
    are bit for bit what dlio_pack_f16-style splitting of the materialised pairs gives at the same bound; pads are zero."""
    from deeplio_b200 import data
    g = torch.Generator().manual_seed(12)
    B, F_, H, W = 2, 3, 6, 16
    frames = (torch.randn(B, F_, 6, H, W, generator=g) * torch.tensor([0.1, 0.1, 0.01, 0.4, 0.4, 0.5]).view(1, 1, 6, 1, 1)).to(DEV)
    combos = [[0, 1], [1, 2]]
    pf = data.PairedFrames(frames, combos, 3, 3)
    h2, bound = data.pair_gather(torch.device(DEV), pf, 8, 2, 4, False, f16=True)
    assert abs(bound.item() - frames.abs().max().item()) <= 1e-6 * frames.abs().max().item()
    pairs = pf.materialize().reshape(B * 2, 6, H, W)                      # channels (t0: c0..c2, t1: c0..c2)
    s = 2.0 ** (14 - torch.tensor(bound.item()).frexp().exponent.item())
    v = torch.zeros(B * 2, H + 4, W + 8, 8, device=DEV)
    v[:, 2:2 + H, 4:4 + W, :6] = pairs.permute(0, 2, 3, 1)
    hi = (v * s).half()
    lo = ((v * s - hi.float()) * 2048.0).half()
    assert torch.equal(h2[..., 0, :].view(torch.int16), hi.view(torch.int16))
    assert torch.equal(h2[..., 1, :].view(torch.int16), lo.view(torch.int16))


def test_nvtx_switch_is_harmless():
    """dlio_set_option("nvtx", 1): ranges are pushed / popped around the launches of an entry point; results unchanged."""
    from deeplio_b200 import _lib as L
    x = torch.randn(1000, device=DEV)
    b0, b1 = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    L.absmax(x.data_ptr(), x.numel(), b0.data_ptr(), torch.cuda.current_stream().cuda_stream)
    L.set_option(b"nvtx", 1)
    try:
        L.absmax(x.data_ptr(), x.numel(), b1.data_ptr(), torch.cuda.current_stream().cuda_stream)
    finally:
        L.set_option(b"nvtx", 0)
    assert b0.item() == b1.item() == x.abs().max().item()
