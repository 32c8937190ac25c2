"""The optimizer / training-state protocol the reference's Trainer relies on (trainer.py:56-58,108,161,269-281;
misc.py:131-165), driven the way the Trainer drives it, on the B200 path."""
import argparse

import pytest
import torch

from oracle import deeplio_oracle as O
from oracle.configs import make_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def small_model(seed=4, **kw):
    from deeplio_b200 import nets
    from deeplio_b200.config import build_config_container
    cfg = make_cfg(height=16, width=64, seq=2, odom_hidden=32, **kw)
    build_config_container(cfg, argparse.Namespace(device=DEV, batch_size=2))
    model = nets.get_model((3, 16, 64), cfg, DEV)
    model.load_state_dict(O.synthetic_state(cfg, seed=seed))
    model.train()
    xyz, normals, imus = O.synthetic_batch(2, 2, 16, 64, 5, seed=seed)
    return model, [[xyz.to(DEV), normals.to(DEV)], imus.to(DEV)]


def loss_of(model, batch, scale=1.0):
    pos, ori = model(batch)
    return scale * ((pos ** 2).sum() + (ori ** 2).sum())


def test_flat_adam_follows_the_reference_call_pattern():
    """trainer.py:56-58: ``Adam([{'params': model.parameters()}, {'params': criterion.parameters()}], lr, wd)``, a
    scheduler that rewrites ``param_groups[i]['lr']``, ``state_dict()`` saved into the checkpoint and loaded on
    resume -- here against torch.optim.Adam fed the same gradients."""
    from deeplio_b200.optim import create_optimizer, FlatAdam
    model, batch = small_model()
    sx = torch.nn.Parameter(torch.tensor(0.0, device=DEV))
    sq = torch.nn.Parameter(torch.tensor(-3.0, device=DEV))
    groups = [{"params": model.parameters()}, {"params": [sx, sq]}]
    opt = create_optimizer(groups, {"optimizer": "adam"}, argparse.Namespace(lr=1e-3, weight_decay=1e-4))
    assert isinstance(opt, FlatAdam) and len(opt.param_groups) == 2
    assert opt.param_groups[0]["lr"] == 1e-3 and opt.param_groups[1]["weight_decay"] == 1e-4
    ours = list(model.parameters()) + [sx, sq]
    twins = [p.detach().clone().requires_grad_(True) for p in ours]
    ref = torch.optim.Adam([{"params": twins[:-2]}, {"params": twins[-2:]}], lr=1e-3, weight_decay=1e-4)
    for step in range(3):
        for g_, r_ in zip(opt.param_groups, ref.param_groups):      # what PolynomialLRDecay does (misc.py:157-163)
            g_["lr"] = r_["lr"] = 1e-3 * (1.0 - 0.2 * step)
        opt.zero_grad()
        (loss_of(model, batch) + (sx - 1.0) ** 2 + (sq * sq)).backward()
        for p, t in zip(ours, twins):
            t.grad = p.grad.detach().clone()
        opt.step()
        ref.step()
    for p, t in zip(ours, twins):
        assert (p.detach() - t.detach()).abs().max().item() < 2e-6
    # checkpoint round trip in torch.optim.Adam's layout: our state loads into torch's Adam and back
    sd = opt.state_dict()
    ref2 = torch.optim.Adam([{"params": twins[:-2]}, {"params": twins[-2:]}], lr=0.5)
    ref2.load_state_dict(sd)
    assert ref2.param_groups[0]["lr"] == opt.param_groups[0]["lr"]
    i = len(ours) - 3
    assert torch.equal(ref2.state[twins[i]]["exp_avg"], sd["state"][i]["exp_avg"])
    assert float(ref2.state[twins[i]]["step"]) == 3.0
    opt2 = FlatAdam([{"params": [p.detach().clone().requires_grad_(True) for p in ours[:-2]]},
                     {"params": [p.detach().clone().requires_grad_(True) for p in ours[-2:]]}], lr=7.0)
    opt2.load_state_dict(ref.state_dict())
    assert opt2.step_count == 3 and opt2.param_groups[0]["lr"] == ref.param_groups[0]["lr"]
    assert torch.allclose(opt2.exp_avg, opt.exp_avg, rtol=1e-5, atol=1e-9)
    assert torch.allclose(opt2.exp_avg_sq, opt.exp_avg_sq, rtol=1e-5, atol=1e-12)


def test_gradients_that_left_the_arena_are_rehomed_and_missing_ones_skipped():
    """``model.zero_grad()`` (set_to_none) detaches every .grad from the arena; the next backward creates fresh
    tensors.  step() must consume THOSE gradients (not an all-zero arena), and must leave a parameter without a
    gradient untouched (torch.optim.Adam semantics: no decay, no moment update)."""
    from deeplio_b200.optim import FlatAdam
    model, batch = small_model()
    opt = FlatAdam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    params = dict(model.named_parameters())
    before = {k: p.detach().clone() for k, p in params.items()}
    model.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in params.values())
    loss_of(model, batch).backward()
    frozen = "fc_ori.bias"
    params[frozen].grad = None
    grads = {k: p.grad.detach().clone() for k, p in params.items() if p.grad is not None}
    opt.step()
    twins = {k: before[k].clone().requires_grad_(True) for k in params}
    ref = torch.optim.Adam(list(twins.values()), lr=1e-3, weight_decay=1e-4)
    for k, g in grads.items():
        twins[k].grad = g
    ref.step()
    for k, p in params.items():
        assert (p.detach() - twins[k].detach()).abs().max().item() < 2e-6, k
    assert torch.equal(params[frozen].detach(), before[frozen])
    # the views are back in place
    opt.zero_grad()
    assert all(p.grad is not None and p.grad.data_ptr() == opt.flat_grad.data_ptr() + 4 * off
               for p, off in zip(opt.params, opt.offsets))


def test_two_backwards_accumulate_like_autograd():
    """Gradient accumulation: two backward passes per zero_grad().  Encoder parameters (written in place by the
    backward kernels) and head / RNN / fc parameters must all hold the SUM of the two micro-batch gradients."""
    from deeplio_b200.optim import FlatAdam
    model, batch = small_model(seed=9)
    xyz2, normals2, imus2 = O.synthetic_batch(2, 2, 16, 64, 5, seed=10)
    batch2 = [[xyz2.to(DEV), normals2.to(DEV)], imus2.to(DEV)]
    # reference: plain autograd accumulation without the arena
    model.zero_grad(set_to_none=True)
    loss_of(model, batch).backward()
    loss_of(model, batch2, 0.5).backward()
    want = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt = FlatAdam(model.parameters(), lr=1e-3)
    opt.zero_grad()
    loss_of(model, batch).backward()
    loss_of(model, batch2, 0.5).backward()
    for k, p in model.named_parameters():
        scale = want[k].abs().max().item() + 1e-12
        assert (p.grad - want[k]).abs().max().item() <= 2e-5 * scale + 1e-9, k
    # and zero_grad() starts over
    opt.zero_grad()
    loss_of(model, batch).backward()
    one = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    loss_of(model, batch).backward()
    for k, p in model.named_parameters():
        scale = one[k].abs().max().item() + 1e-12
        assert (p.grad - one[k]).abs().max().item() <= 2e-5 * scale + 1e-9, k


@pytest.mark.parametrize("momentum", [0.3, None])
def test_batchnorm_momentum_comes_from_the_module(momentum):
    """PointSeg builds nn.BatchNorm2d(momentum=bn_d) (pointseg_modules.py:98-106); momentum=None is torch's
    cumulative moving average.  Two train-mode steps against F.batch_norm semantics on the oracle's statistics."""
    model, batch = small_model(lidar="lidar-feat-pointseg")
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = momentum
    for e in (model.lidar_feat_net.encoder1, model.lidar_feat_net.encoder2):
        e.__dict__.pop("_bn_cfg", None)
    name = "lidar_feat_net.encoder1.conv1a.1"
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        model(batch)
        model(batch)
    sd2 = model.state_dict()
    assert int(sd2[name + ".num_batches_tracked"]) == int(sd0[name + ".num_batches_tracked"]) + 2
    # batch statistics of conv1a's output from the oracle's conv
    xyz = batch[0][0].cpu()
    x = xyz.reshape(-1, 6, 16, 64)
    w, b = sd0["lidar_feat_net.encoder1.conv1a.0.weight"].cpu(), sd0["lidar_feat_net.encoder1.conv1a.0.bias"].cpu()
    y = torch.nn.functional.conv2d(x, w, b, (1, 2), (1, 2))
    rm, rv = sd0[name + ".running_mean"].cpu().clone(), sd0[name + ".running_var"].cpu().clone()
    nbt = int(sd0[name + ".num_batches_tracked"])
    for _ in range(2):
        nbt += 1
        f = momentum if momentum is not None else 1.0 / nbt
        torch.nn.functional.batch_norm(y, rm, rv, None, None, True, f, 1e-5)
    assert torch.allclose(sd2[name + ".running_mean"].cpu(), rm, rtol=1e-5, atol=1e-6)
    assert torch.allclose(sd2[name + ".running_var"].cpu(), rv, rtol=1e-5, atol=1e-6)


def test_graph_construction_leaves_the_training_state_alone():
    """Building the CUDA-graph step runs warm-up passes; BatchNorm running statistics, num_batches_tracked and the
    dropout call counter must come out as they went in."""
    from deeplio_b200 import functional as Fn
    from deeplio_b200.graph import GraphedTrainStep
    from deeplio_b200.optim import FlatAdam
    with torch.cuda.stream(torch.cuda.Stream()):
        model, batch = small_model(seed=6, no_dropout=False)
        opt = FlatAdam(model.parameters(), lr=1e-3)
        d = {"xyz": batch[0][0], "normals": batch[0][1], "imus": batch[1]}
        before = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k or "num_batches" in k}
        drop0 = Fn._drop_counter[0]
        step = GraphedTrainStep(lambda t: loss_of(model, [[t["xyz"], t["normals"]], t["imus"]]), d, opt.zero_grad,
                                model=model)
        torch.cuda.synchronize()
        after = model.state_dict()
        for k, v in before.items():
            assert torch.equal(after[k], v), k
        assert Fn._drop_counter[0] == drop0
        step(d)
        torch.cuda.synchronize()
        k = next(k for k in before if k.endswith("num_batches_tracked"))
        assert int(model.state_dict()[k]) == int(before[k]) + 1
