"""Shared test helpers (oracle side)."""
import os

import torch

from oracle import deeplio_oracle as O
from oracle.configs import make_cfg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt"))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def case_setup(rec):
    """Rebuild (cfg, state, inputs) of a golden record from its seeds."""
    cfg = make_cfg(height=rec["H"], width=rec["W"], seq=rec["S"], **rec["kwargs"])
    sd = O.synthetic_state(cfg, seed=rec["seed"])
    xyz, normals, imus = O.synthetic_batch(rec["B"], rec["S"], rec["H"], rec["W"], rec["T"], seed=rec["seed"])
    chk = torch.stack([xyz.double().sum(), normals.double().sum(), imus.double().sum()]).float()
    assert torch.allclose(chk, rec["input_checksum"], rtol=1e-6, atol=1e-6), "RNG drift: inputs differ from golden"
    return cfg, sd, (xyz, normals, imus)


def oracle_train_step(cfg, sd, inputs):
    """Forward (train mode) + backward of sum(pos^2)+sum(ori^2) on the oracle; returns pos, ori, grads, sd."""
    xyz, normals, imus = inputs
    sd = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running_" not in k) else v.clone())
          for k, v in sd.items()}
    pos, ori = O.deeplio_forward(sd, cfg, xyz, normals, imus, training=True)
    loss = (pos ** 2).sum() + (ori ** 2).sum()
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items() if v.requires_grad}
    return pos.detach(), ori.detach(), grads, sd


def rel_err(a, b, floor=0.0):
    """max|a-b| / (max|b| + floor)."""
    return (a - b).abs().max().item() / (b.abs().max().item() + floor + 1e-30)
