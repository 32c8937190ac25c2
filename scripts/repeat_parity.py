"""Repeats the train-step parity check of one golden case N times in one process (GPU box) and prints, per
iteration, the worst parameter-gradient error against the fp64 oracle -- a nondeterminism detector.  With
DLIO_TRACE=1 it also diffs every conv_bn backward tensor (dout, dz, sums, dy, y, BN vectors) against the previous
iteration, which is how the knife-edge ReLU elements of the small fixtures were found: y and dout agree to 1e-6
while dz differs by 30 % (one mask flip among 64 samples per channel).
usage: [DLIO_TRACE=1] python scripts/repeat_parity.py <golden-case> [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import case_setup, load_golden, oracle_train_step, rel_err  # noqa: E402
from tests.test_gpu_model import build_b200_model, to_dev  # noqa: E402

name = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rec = load_golden(name)
cfg, sd, inputs = case_setup(rec)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
p64, o64, g64, _ = oracle_train_step(cfg, sd64, tuple(t.double() for t in inputs))
junk = []
from deeplio_b200 import engine as E  # noqa: E402
prev = None
TRACE = bool(int(os.environ.get("DLIO_TRACE", "0")))
for it in range(N):
    E.DEBUG_TRACE = {} if TRACE else None
    model = build_b200_model(cfg, rec["H"], rec["W"], rec["B"], sd)
    model.train()
    pos, ori = model(to_dev(inputs))
    ((pos ** 2).sum() + (ori ** 2).sum()).backward()
    worst, wk = 0.0, None
    for k, p in model.named_parameters():
        e = rel_err(p.grad.cpu().double(), g64[k]) if g64[k].abs().max() > 0 else 0.0
        if e > worst:
            worst, wk = e, k
    print("iter %2d  pos %.2e ori %.2e  worst grad %.2e (%s)" % (
        it, rel_err(pos.detach().cpu().double(), p64), rel_err(ori.detach().cpu().double(), o64), worst, wk), flush=True)
    if TRACE:
        cur = [(k[1], v) for k, v in E.DEBUG_TRACE.items()]      # insertion order = backward order, enc1 then enc2
        if prev is not None:
            for (n1, a), (n2, b) in zip(prev, cur):
                d = {t: (a[t].double() - b[t].double()).abs().max().item() / (a[t].double().abs().max().item() + 1e-30) for t in a}
                if max(d.values()) > 1e-4:
                    print("   vs prev iter: %-12s " % n1 + " ".join("%s %.1e" % kv for kv in d.items()))
        prev = cur
    # perturb the caching allocator's state between iterations
    junk.append(torch.full((1 + 257 * it,), float("nan"), device="cuda"))
    if it % 3 == 2:
        junk.clear()
        torch.cuda.empty_cache()
