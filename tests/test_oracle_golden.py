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
