"""A/B asked for by the round-1 review (item 4): what do single-pass fp16 dgrad / wgrad cost in gradient accuracy?
Runs one train-mode step of the headline workload (Simple-1 + bi-LSTM IMU + soft fusion + LSTM odometry, 64 x 2048,
batch 8 x S 2) twice on the same weights and inputs -- the shipped three-product scheme and the "bwd_single_pass"
switch -- and prints, per parameter tensor, max|g1 - g3| / max|g3|.  The three-product gradients are the ones the
parity tests hold to 2e-4 of the fp64 oracle, so this difference IS the single-pass error at that resolution.
Step times of the two modes: bench.py with DLIO_BWD_SINGLE=1.
usage (GPU box): python scripts/ab_single_pass.py > gpurun_out/ab_single_pass.json"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import H, W, WORKLOAD, G0, G1, LR, WD, synthetic_host_batch  # noqa: E402
from deeplio_b200 import _lib, data, losses, nets, pose  # noqa: E402
from deeplio_b200.config import build_config_container  # noqa: E402
from deeplio_b200.optim import FlatAdam  # noqa: E402
from deeplio_b200.workloads import workload_config  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg, B, S, T = workload_config(WORKLOAD, H, W)
    for sec in ("lidar-feat-simple-1", "imu-feat-rnn", "odom-feat-rnn", "deeplio", "fusion-layer-soft"):
        if isinstance(cfg.get(sec), dict) and "dropout" in cfg[sec]:
            cfg[sec]["dropout"] = 0.0                      # the two runs must see the same masks
    combos = cfg["datasets"]["combinations"]
    build_config_container(cfg, argparse.Namespace(device=str(dev), batch_size=B))
    torch.manual_seed(1234)
    model = nets.get_model((3, H, W), cfg, str(dev))
    criterion = losses.get_loss_function(cfg, str(dev))
    model.train()
    opt = FlatAdam([{"params": model.parameters()}, {"params": criterion.parameters()}], lr=LR, weight_decay=WD)
    d = {k: v.to(dev) for k, v in synthetic_host_batch(B, S, T, seed=100).items()}
    bn0 = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}

    def grads(single):
        _lib.set_option(b"bwd_single_pass", 1 if single else 0)
        model.load_state_dict(bn0, strict=False)
        opt.zero_grad()
        f2f, f2g = data.ground_truth(d["gts"], combos)
        pos, ori = model([[data.PairedFrames(d["frames"], combos, 0, 3), data.PairedFrames(d["frames"], combos, 3, 3)],
                          d["imus"]])
        p, q = pose.se3_to_SE3(pos, ori, check=False)
        loss = criterion(pos, ori, p[:, G0:G1], q[:, G0:G1], f2f[:, :, 0:3], f2f[:, :, 3:], f2g[:, G0:G1, 0:3],
                         f2g[:, G0:G1, 3:7])
        loss.backward()
        torch.cuda.synchronize()
        _lib.set_option(b"bwd_single_pass", 0)
        return float(loss), opt.flat_grad.clone()

    l3, g3 = grads(False)
    l3b, g3b = grads(False)        # run-to-run noise of the shipped path (atomics order, mask flips)
    l1, g1 = grads(True)
    names = {id(p): k for k, p in list(model.named_parameters()) + [("criterion." + k, p) for k, p in criterion.named_parameters()]}
    rows = []
    for p, off in zip(opt.params, opt.offsets):
        a, b, c = (g[off:off + p.numel()] for g in (g3, g3b, g1))
        scale = a.abs().max().item()
        if scale == 0.0:
            continue
        rows.append({"param": names[id(p)], "single_vs_three": (c - a).abs().max().item() / scale,
                     "three_vs_three": (b - a).abs().max().item() / scale,
                     "rms_single_vs_three": ((c - a).double().pow(2).mean().sqrt() / a.double().pow(2).mean().sqrt()).item()})
    enc = [r for r in rows if "encoder" in r["param"]]
    out = {"workload": WORKLOAD, "batch": B, "loss_three": l3, "loss_single": l1,
           "encoder_tensors": len(enc),
           "max_err_single_vs_three": max(r["single_vs_three"] for r in enc),
           "median_err_single_vs_three": sorted(r["single_vs_three"] for r in enc)[len(enc) // 2],
           "max_err_three_vs_three": max(r["three_vs_three"] for r in enc),
           "median_rms_single_vs_three": sorted(r["rms_single_vs_three"] for r in enc)[len(enc) // 2],
           "tensors_within_2e-4": sum(r["single_vs_three"] < 2e-4 for r in enc),
           "worst": sorted(enc, key=lambda r: -r["single_vs_three"])[:8],
           "non_encoder_max": max((r["single_vs_three"] for r in rows if "encoder" not in r["param"]), default=None)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
