"""Shared test helpers (oracle side)."""
import os

import torch

from oracle import deeplio_oracle as O
from oracle.configs import make_cfg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt") and not f.startswith("pose_"))


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


# ----------------------------------------------------------------------------- gradient-parity bookkeeping
def diag(record):
    """Appends one JSON line to gpurun_out/parity_diag.jsonl (the per-tensor numbers behind the parity assertions
    travel back from the GPU box with the rest of gpurun_out/)."""
    import json
    root = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_diag.jsonl"), "a") as f:
            f.write(json.dumps(record) + "\n")
    except OSError:
        pass


def f64_state(sd, inputs):
    return ({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()},
            tuple(t.double() for t in inputs))


def perturbed_grads(cfg, sd64, in64, draws):
    """fp64 oracle gradients under relative perturbations of every weight and input: [(seed, amplitude), ...]."""
    out = []
    for seed, amp in draws:
        gen = torch.Generator().manual_seed(seed)

        def jitter(t):
            return t * (1.0 + amp * torch.randn(t.shape, generator=gen, dtype=torch.float64)) if t.is_floating_point() else t
        out.append(oracle_train_step(cfg, {k: jitter(v) for k, v in sd64.items()}, tuple(jitter(t) for t in in64))[2])
    return out


def count_relu_flips(ours, oracle, prefix="lidar_feat_net."):
    """Discrete decisions on which the B200 path (engine.MASK_TRACE: ReLU masks, bool NCHW; max-pool arg-max, uint8
    NCHW under '<conv>#pool') and the oracle (oracle.TRACE, same form) disagree: {layer: (differing, total)}."""
    flips = {}
    for k, m in ours.items():
        ref = oracle.get(prefix + k)
        assert ref is not None and tuple(ref.shape) == tuple(m.shape), ("decision trace mismatch", k)
        flips[k] = (int((m != ref).sum()), m.numel())
    return flips


def forced_oracle_step(cfg, sd, inputs, decisions, dtype=torch.float64, prefix="lidar_feat_net."):
    """One train-mode forward + backward of the oracle in ``dtype`` with the B200 path's discrete decisions (ReLU
    masks and max-pool arg-max, engine.MASK_TRACE) imposed on it.  Returns (pos, ori, grads, the oracle's OWN
    decisions for ``count_relu_flips``).  Why: a parameter gradient is a sum of 10^5 .. 10^8 terms of random sign, so
    one ReLU input (or one pair of pooling candidates) within round-off of a tie, decided the other way, moves it by
    ~1 / sqrt(terms) -- 1e-3 .. 1e-4, above the 2e-4 bar -- in ANY two fp32 implementations (the fp32 oracle against
    the fp64 oracle included).  With the decisions imposed, what is compared is the arithmetic."""
    def cast(t):
        return t.to(dtype) if t.is_floating_point() else t
    O.TRACE = {}
    O.FORCE_MASKS = {prefix + k: v for k, v in decisions.items()}
    try:
        pos, ori, grads, _ = oracle_train_step(cfg, {k: cast(v) for k, v in sd.items()}, tuple(cast(t) for t in inputs))
        own = O.TRACE
    finally:
        O.TRACE = None
        O.FORCE_MASKS = None
    return pos, ori, grads, own


def grad_rows(grads, g64, g32=None, gperts=()):
    """Per-tensor parity numbers: [(name, err, scale, e_ref, e_pert)], all max-abs against the fp64 gradient."""
    rows = []
    for k, g in grads.items():
        ref = g64[k]
        scale = ref.abs().max().item()
        e = (g.double() - ref).abs().max().item()
        e_ref = (g32[k].double() - ref).abs().max().item() if g32 is not None else 0.0
        e_pert = max(((gp[k] - ref).abs().max().item() for gp in gperts), default=0.0)
        rows.append((k, e, scale, e_ref, e_pert))
    return rows


def quat_tol(q_ref, base=5e-6):
    """Per-entry tolerance for quaternions [.., 4] (wxyz) obtained from rotation matrices: ``base`` amplified by
    1 / (4 |qw|) where the conversion divides by 4 qw (capped at the near-zero branch, |qw| < 1e-6 -> other formulas)."""
    qw = q_ref[..., 0:1].abs()
    return base * (1.0 + 1.0 / (4.0 * qw.clamp_min(1e-3))).expand_as(q_ref)
