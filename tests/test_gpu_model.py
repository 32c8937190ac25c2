"""Parity of the B200 path with the CPU oracle and with the golden vectors generated from the unmodified
reference (tests/golden/*.pt, oracle/make_golden.py): forward (pos, ori), every parameter gradient, the BN
running statistics after one train-mode step, and an eval-mode forward.

Bars: north_star asks for pose outputs within 1e-4 relative of the reference; the tests use 2e-5 for the
forward and 2e-4 of the largest gradient for gradients (same bar the oracle is held to against the goldens).
"""
import argparse

import pytest
import torch

from oracle import deeplio_oracle as O
from tests.helpers import (GOLDEN_CASES, case_setup, count_relu_flips, diag, forced_oracle_step, grad_rows, load_golden,
                           oracle_train_step, rel_err)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FWD_TOL = 2e-5
GRAD_TOL = 2e-4


def build_b200_model(cfg, H, W, B, sd):
    from deeplio_b200 import nets
    from deeplio_b200.config import build_config_container
    build_config_container(cfg, argparse.Namespace(device=DEV, batch_size=B))
    model = nets.get_model((3, H, W), cfg, DEV)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model


def to_dev(inputs):
    xyz, normals, imus = inputs
    return [[xyz.to(DEV), normals.to(DEV)], imus.to(DEV)]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_model_matches_oracle_and_golden(name):
    rec = load_golden(name)
    cfg, sd, inputs = case_setup(rec)
    model = build_b200_model(cfg, rec["H"], rec["W"], rec["B"], sd)
    model.train()
    from deeplio_b200 import engine as E
    E.MASK_TRACE = {}
    try:
        pos, ori = model(to_dev(inputs))
        mtrace = E.MASK_TRACE
    finally:
        E.MASK_TRACE = None
    loss = (pos ** 2).sum() + (ori ** 2).sum()
    loss.backward()
    torch.cuda.synchronize()
    # forward vs the reference-generated golden and vs the oracle
    assert rel_err(pos.detach().cpu(), rec["pos"]) < FWD_TOL
    assert rel_err(ori.detach().cpu(), rec["ori"]) < FWD_TOL
    opos, oori, ograds, osd = oracle_train_step(cfg, sd, inputs)
    assert rel_err(pos.detach().cpu(), opos) < FWD_TOL
    assert rel_err(ori.detach().cpu(), oori) < FWD_TOL
    # gradients: EVERY parameter at the plain bar GRAD_TOL (2e-4 of the tensor's largest entry; tensors whose exact
    # gradient is zero -- a convolution bias in front of a train-mode BatchNorm -- are held to 1e-5 of the model's
    # largest entry), against an fp64 evaluation of the oracle WITH THE B200 PATH'S DISCRETE DECISIONS IMPOSED (ReLU
    # masks, max-pool arg-max: tests.helpers.forced_oracle_step).  A ReLU input within round-off of zero that lands on
    # the other side moves a gradient -- a sum of thousands to 10^8 terms of random sign -- by ~1 / sqrt(terms), far
    # more than any arithmetic error, in any two fp32 implementations (round 1 needed sensitivity-scaled bars for
    # that); the differing decisions are counted and bounded instead (measured: 0 .. 12 of 2 .. 6 million).
    _, _, g64, own = forced_oracle_step(cfg, sd, inputs, mtrace)
    flips = count_relu_flips(mtrace, own)
    n_flips, n_dec = sum(f for f, _ in flips.values()), sum(t for _, t in flips.values())
    assert n_flips <= 2e-5 * n_dec + 2, (n_flips, n_dec)
    gmax = max(float(n) for n, _ in rec["grads"].values())
    params = dict(model.named_parameters())
    assert set(params) == set(ograds)
    ours = {k: (p.grad.cpu() if p.grad is not None else torch.zeros_like(ograds[k])) for k, p in params.items()}
    rows = grad_rows(ours, g64)
    worst = max((r for r in rows if r[2] > 1e-3 * gmax), key=lambda r: r[1] / r[2])
    diag({"test": "golden", "case": name, "tensors": len(rows), "flips": n_flips, "decisions": n_dec,
          "worst": (worst[0], worst[1] / worst[2])})
    for k, e_ours, scale, _, _ in rows:
        assert e_ours <= GRAD_TOL * scale + 1e-5 * gmax, (k, e_ours / (scale + 1e-30), scale, gmax)
        # the reference-generated golden norms (each side with its own decisions): the oracle's own distance from the
        # imposed-decision gradients bounds what the flips may move
        e_nat = (ograds[k].double() - g64[k]).abs().max().item()
        norm, head = rec["grads"][k]
        bar = GRAD_TOL + 4 * e_nat / (scale + 1e-30)
        assert abs(ours[k].double().norm().item() - float(norm)) <= bar * float(norm) + 1e-5 * gmax, k
    # dead-direction RNN parameters still get (zero) gradient tensors, so Adam + L2 decay updates them
    for k, p in params.items():
        if "_l1_reverse" in k:
            assert p.grad is not None and p.grad.abs().max().item() == 0.0, k
    # BN running statistics after the step
    msd = model.state_dict()
    for k, s in rec["running"].items():
        assert abs(msd[k].double().sum().item() - float(s)) <= 2e-5 * max(1.0, abs(float(s))), k
    for k, v in msd.items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == 1
    # eval-mode forward (running statistics, no dropout)
    model.eval()
    with torch.no_grad():
        epos, eori = model(to_dev(inputs))
    assert rel_err(epos.cpu(), rec["eval_pos"]) < FWD_TOL
    assert rel_err(eori.cpu(), rec["eval_ori"]) < FWD_TOL


def test_strided_xyz_view_needs_no_copy():
    """The trainer hands xyz as a channel-slice view of a [B,S,2,6,H,W] tensor (misc.py:65-69)."""
    rec = load_golden("lidar_only_simple1")
    cfg, sd, (xyz, normals, imus) = case_setup(rec)
    model = build_b200_model(cfg, rec["H"], rec["W"], rec["B"], sd).eval()
    pairs = torch.cat([xyz, normals], dim=3).to(DEV)            # [B,S,2,6,H,W]
    view = pairs[:, :, :, 0:3]
    assert not view.is_contiguous()
    with torch.no_grad():
        a = model([[view, pairs[:, :, :, 3:].contiguous()], imus.to(DEV)])
        b = model([[xyz.to(DEV).contiguous(), normals.to(DEV)], imus.to(DEV)])
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_dropout_active_in_train_mode_and_off_in_eval():
    from oracle.configs import make_cfg
    cfg = make_cfg(seq=2, height=16, width=64, no_dropout=False, odom_hidden=64)
    sd = O.synthetic_state(cfg, seed=3)
    model = build_b200_model(cfg, 16, 64, 2, sd)
    xyz, normals, imus = O.synthetic_batch(2, 2, 16, 64, 5, seed=3)
    model.train()
    torch.manual_seed(1)
    a = model(to_dev((xyz, normals, imus)))[0]
    b = model(to_dev((xyz, normals, imus)))[0]
    assert not torch.equal(a, b)           # different masks on consecutive calls
    model.eval()
    with torch.no_grad():
        c = model(to_dev((xyz, normals, imus)))[0]
        d = model(to_dev((xyz, normals, imus)))[0]
    assert torch.equal(c, d)


def test_factory_errors_match_reference():
    from deeplio_b200 import nets
    from deeplio_b200.config import build_config_container
    from oracle.configs import make_cfg
    cfg = make_cfg(height=16, width=64)
    build_config_container(cfg, argparse.Namespace(device=DEV, batch_size=1))
    cfg["deeplio"]["lidar-feat-net"]["name"] = "lidar-feat-nope"
    with pytest.raises(ValueError):
        nets.get_model((3, 16, 64), cfg, DEV)
    with pytest.raises(RuntimeError):
        nets.get_model((3, 16, 64), make_cfg(height=16, width=64), "cpu")
