"""Per-parameter parity report of the B200 path against the CPU oracle for one golden case (GPU box).
usage: python scripts/diag_parity.py <golden-case> [...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import case_setup, load_golden, oracle_train_step, rel_err  # noqa: E402
from tests.test_gpu_model import build_b200_model, to_dev  # noqa: E402

for name in sys.argv[1:]:
    rec = load_golden(name)
    cfg, sd, inputs = case_setup(rec)
    model = build_b200_model(cfg, rec["H"], rec["W"], rec["B"], sd)
    model.train()
    pos, ori = model(to_dev(inputs))
    ((pos ** 2).sum() + (ori ** 2).sum()).backward()
    opos, oori, ograds, _ = oracle_train_step(cfg, sd, inputs)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    p64, o64, g64, _ = oracle_train_step(cfg, sd64, tuple(t.double() for t in inputs))
    print("== %s  vs fp64: ours pos %.2e ori %.2e | oracle32 pos %.2e ori %.2e" % (
        name, rel_err(pos.detach().cpu().double(), p64), rel_err(ori.detach().cpu().double(), o64),
        rel_err(opos.double(), p64), rel_err(oori.double(), o64)))
    print("== %s  pos %.2e  ori %.2e" % (name, rel_err(pos.detach().cpu(), opos), rel_err(ori.detach().cpu(), oori)))
    gmax = max(g.abs().max().item() for g in ograds.values())
    for k, p in model.named_parameters():
        g = p.grad.cpu()
        o = ograds[k]
        t = g64[k]
        print("%-60s max|g| %.3e  ours-vs-oracle32 %.2e  ours-vs-fp64 %.2e  oracle32-vs-fp64 %.2e" % (
            k, o.abs().max().item(), rel_err(g, o), rel_err(g.double(), t), rel_err(o.double(), t)))
