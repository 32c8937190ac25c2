"""Generate tests/golden/pose_glue.pt by EXECUTING the reference's own caller-side glue (container only).

TEST INFRASTRUCTURE.  ``python -m oracle.make_golden_pose`` from the repo root, where /root/reference exists.
The functions run where they lie in the reference tree, unmodified:

* ``Trainer.se3_to_SE3``            deeplio/models/trainer.py:324-351
* ``DataCombiCreater.process_ground_turth``   deeplio/models/misc.py:83-125
* ``HWSLoss`` / ``LWSLoss``         deeplio/losses/losses.py:11-96

``liegroups`` is absent from this image, so ``liegroups.torch.SO3`` -- what the first two call -- is
``oracle.pose_oracle.SO3``, the restatement of that library (parity UNPINNED for the SO(3) maps themselves, pinned
for everything the reference does around them: chaining order, indexing, inverse, slicing, loss weighting).
Inputs are stored with the outputs (they are small).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pose_oracle as P  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "pose_glue.pt")


def reference_modules():
    ref_loader._install_stubs()
    import liegroups.torch as lt
    lt.SO3 = P.SO3
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    with ref_loader._in_tmp_cwd():
        import deeplio.losses as losses
        import deeplio.models.misc as misc
        import deeplio.models.trainer as trainer
    misc.SO3 = trainer.SO3 = P.SO3     # they bound the name at import time, possibly to an earlier empty stub
    return trainer, misc, losses


def chain_inputs(B=5, S=4, seed=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, S, 3, generator=g) * 0.5
    w = torch.randn(B, S, 3, generator=g) * 0.05
    w[0, 1] = torch.tensor([2e-7, -3e-7, 1e-7])             # first-order branch of exp (|w| < 1e-6)
    w[1, 0] = torch.tensor([3.1, 0.2, -0.1])                # near pi: quaternion w close to zero afterwards
    w[1, 1] = torch.tensor([0.02, 0.01, 0.0])
    w[2, 2] = torch.tensor([0.0, 3.14159, 0.0])             # rotation by ~pi about y: qw ~ 0 -> near-zero branch
    w[2, 0] = w[2, 1] = torch.zeros(3)
    w[3] = torch.randn(S, 3, generator=g) * 1.0             # large rotations
    return x, w


def main():
    trainer, misc, losses = reference_modules()
    rec = {}
    # ---- se3_to_SE3
    x, w = chain_inputs()
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    me = types.SimpleNamespace(device="cpu")
    f2g_x, f2g_q = trainer.Trainer.se3_to_SE3(me, xr, wr)
    gx, gq = torch.randn(f2g_x.shape, generator=torch.Generator().manual_seed(1)), \
        torch.randn(f2g_q.shape, generator=torch.Generator().manual_seed(2))
    ((f2g_x * gx).sum() + (f2g_q * gq).sum()).backward()
    rec["chain"] = dict(x=x, w=w, f2g_x=f2g_x.detach(), f2g_q=f2g_q.detach(), gx=gx, gq=gq, dx=xr.grad.clone(),
                        dw=wr.grad.clone())
    # ---- ground truth
    combos = [[0, 1], [1, 2], [2, 3]]
    gts = P.synthetic_gts(4, 4, seed=5)
    me = types.SimpleNamespace(combinations=combos, device="cpu")
    f2f, f2g = [], []
    for b in range(gts.shape[0]):
        a, c = misc.DataCombiCreater.process_ground_turth(me, gts[b])
        f2f.append(a)
        f2g.append(c)
    rec["gt"] = dict(gts=gts, combinations=combos, f2f=torch.stack(f2f), f2g=torch.stack(f2g))
    # ---- losses, the way trainer.py:246-263 slices and calls them (max_glob_seq = 2)
    g = torch.Generator().manual_seed(9)
    B, S = 3, 3
    pt, pw = torch.randn(B, S, 3, generator=g) * 0.3, torch.randn(B, S, 3, generator=g) * 0.02
    gt_f2f = torch.cat([torch.randn(B, S, 3, generator=g) * 0.3, torch.randn(B, S, 3, generator=g) * 0.02], 2)
    gt_f2g = torch.cat([torch.randn(B, S, 3, generator=g), torch.nn.functional.normalize(torch.randn(B, S, 4, generator=g), dim=2)], 2)
    cases = {}
    for name, crit in (("hws_both", losses.HWSLoss(sx=0.0, sq=-3.0, learn_hyper_params=True, loss_Types=[True, True])),
                       ("hws_local", losses.HWSLoss(sx=0.3, sq=-2.5, learn_hyper_params=True, loss_Types=[True, False])),
                       ("hws_global", losses.HWSLoss(sx=-0.2, sq=-1.0, learn_hyper_params=True, loss_Types=[False, True])),
                       ("lws_both", losses.LWSLoss(beta=1125.0, loss_Types=[True, True]))):
        ptr, pwr = pt.clone().requires_grad_(True), pw.clone().requires_grad_(True)
        me = types.SimpleNamespace(device="cpu")
        p, q = trainer.Trainer.se3_to_SE3(me, ptr, pwr)
        a, b_ = ptr, pwr
        if crit.loss_Types[0] and not crit.loss_Types[1]:
            p, q = p.detach(), q.detach()
        elif crit.loss_Types[1] and not crit.loss_Types[0]:
            a, b_ = ptr.detach(), pwr.detach()
        loss = crit(a, b_, p[:, 1:3, :], q[:, 1:3, :], gt_f2f[:, :, 0:3], gt_f2f[:, :, 3:], gt_f2g[:, 1:3, 0:3],
                    gt_f2g[:, 1:3, 3:7])
        loss.backward()
        out = dict(loss=loss.detach(), dt=ptr.grad.clone() if ptr.grad is not None else torch.zeros_like(pt),
                   dw=pwr.grad.clone() if pwr.grad is not None else torch.zeros_like(pw),
                   loss_types=list(crit.loss_Types))
        if hasattr(crit, "sx"):
            out.update(sx=crit.sx.detach().clone(), sq=crit.sq.detach().clone(), dsx=crit.sx.grad.clone(),
                       dsq=crit.sq.grad.clone())
        else:
            out.update(beta=crit.beta)
        cases[name] = out
    rec["loss"] = dict(pred_t=pt, pred_w=pw, gt_f2f=gt_f2f, gt_f2g=gt_f2g, g0=1, g1=3, cases=cases)
    torch.save(rec, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
