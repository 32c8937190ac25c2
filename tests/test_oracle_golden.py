"""The oracle restatement reproduces the golden vectors generated from the unmodified reference
(oracle/make_golden.py).  CPU only; this is the pin for oracle/deeplio_oracle.py."""
import pytest
import torch

from oracle import deeplio_oracle as O
from tests.helpers import GOLDEN_CASES, case_setup, load_golden, oracle_train_step, rel_err

FWD_TOL = 2e-6    # fp32 CPU vs fp32 CPU, different op order only
GRAD_TOL = 2e-4   # relative to the largest gradient norm of the model (BN-cancelled grads are ~0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    rec = load_golden(name)
    cfg, sd, inputs = case_setup(rec)
    pos, ori, grads, sd_after = oracle_train_step(cfg, sd, inputs)
    assert rel_err(pos, rec["pos"]) < FWD_TOL
    assert rel_err(ori, rec["ori"]) < FWD_TOL
    gmax = max(float(n) for n, _ in rec["grads"].values())
    assert set(grads) == set(rec["grads"])
    for k, (norm, head) in rec["grads"].items():
        g = grads[k]
        assert abs(g.double().norm().item() - float(norm)) <= GRAD_TOL * float(norm) + 1e-5 * gmax, k
        hd = g.flatten()[: head.numel()]
        assert (hd - head).abs().max().item() <= GRAD_TOL * g.abs().max().item() + 1e-5 * gmax, k
    for k, s in rec["running"].items():
        got = sd_after[k].double().sum().item()
        assert abs(got - float(s)) <= 1e-5 * max(1.0, abs(float(s))), k
    with torch.no_grad():
        epos, eori = O.deeplio_forward(sd_after, cfg, *inputs, training=False)
    assert rel_err(epos, rec["eval_pos"]) < FWD_TOL
    assert rel_err(eori, rec["eval_ori"]) < FWD_TOL


def test_pool_out_matches_torch():
    import torch.nn.functional as F
    for n in list(range(3, 40)) + [64, 65, 129, 257, 513, 1024, 2048]:
        for s in (1, 2):
            for ceil in (False, True):
                ref = F.max_pool2d(torch.zeros(1, 1, n, 8), 3, (s, 1), 1, ceil_mode=ceil).shape[2]
                assert O.pool_out(n, s, ceil) == ref, (n, s, ceil)


# ----------------------------------------------------------------------------- caller-side glue
def test_pose_oracle_matches_reference_glue_goldens():
    """oracle/pose_oracle.py against tests/golden/pose_glue.pt, the outputs of the reference's own
    Trainer.se3_to_SE3, DataCombiCreater.process_ground_turth and HWSLoss / LWSLoss (oracle/make_golden_pose.py)."""
    import torch
    from oracle import pose_oracle as P
    from tests.helpers import GOLDEN_DIR
    import os
    rec = torch.load(os.path.join(GOLDEN_DIR, "pose_glue.pt"), weights_only=False)
    c = rec["chain"]
    x, w = c["x"].clone().requires_grad_(True), c["w"].clone().requires_grad_(True)
    fx, fq, status = P.se3_to_SE3(x, w)
    assert status == 0
    assert torch.allclose(fx, c["f2g_x"], rtol=1e-6, atol=1e-6)
    # matrix -> quaternion divides by 4 qw: near a half turn (qw -> 0) round-off of R (and whether the SVD projection
    # ran: |det - 1| < 1e-6 is decided at round-off level) is amplified by 1 / (4 qw)
    from tests.helpers import quat_tol
    assert ((fq - c["f2g_q"]).abs() <= quat_tol(c["f2g_q"])).all()
    ((fx * c["gx"]).sum() + (fq * c["gq"]).sum()).backward()
    well = (c["f2g_q"][:, :, 0].abs().min(dim=1).values > 0.05)      # samples without a near-half-turn pose
    assert torch.allclose(x.grad[well], c["dx"][well], rtol=1e-5, atol=1e-5)
    assert torch.allclose(w.grad[well], c["dw"][well], rtol=1e-4, atol=1e-4)
    g = rec["gt"]
    f2f, f2g = P.ground_truth(g["gts"], g["combinations"])
    assert torch.allclose(f2f, g["f2f"], rtol=1e-5, atol=2e-6)
    assert torch.allclose(f2g, g["f2g"], rtol=1e-5, atol=2e-6)
    lo = rec["loss"]
    for name, case in lo["cases"].items():
        pt, pw = lo["pred_t"].clone().requires_grad_(True), lo["pred_w"].clone().requires_grad_(True)
        p, q, _ = P.se3_to_SE3(pt, pw)
        lt = case["loss_types"]
        a, b = (pt, pw) if lt[0] else (pt.detach(), pw.detach())
        if not lt[1]:
            p, q = p.detach(), q.detach()
        kw = {}
        if "sx" in case:
            kw = dict(sx=case["sx"].clone().requires_grad_(True), sq=case["sq"].clone().requires_grad_(True))
        else:
            kw = dict(beta=case["beta"])
        loss = P.pose_loss(a, b, p[:, lo["g0"]:lo["g1"]], q[:, lo["g0"]:lo["g1"]], lo["gt_f2f"][:, :, 0:3],
                           lo["gt_f2f"][:, :, 3:], lo["gt_f2g"][:, lo["g0"]:lo["g1"], 0:3],
                           lo["gt_f2g"][:, lo["g0"]:lo["g1"], 3:7], loss_types=lt, **kw)
        loss.backward()
        assert torch.allclose(loss.detach(), case["loss"], rtol=1e-6, atol=1e-6), name
        assert torch.allclose(pt.grad if pt.grad is not None else torch.zeros_like(pt), case["dt"], rtol=1e-5, atol=1e-6), name
        assert torch.allclose(pw.grad if pw.grad is not None else torch.zeros_like(pw), case["dw"], rtol=1e-4, atol=1e-5), name
        if "sx" in case:
            assert torch.allclose(kw["sx"].grad, case["dsx"], rtol=1e-6, atol=1e-6), name
            assert torch.allclose(kw["sq"].grad, case["dsq"], rtol=1e-6, atol=1e-6), name
