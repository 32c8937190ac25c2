"""Generate tests/golden/*.pt from the UNMODIFIED reference (container only).

TEST INFRASTRUCTURE.  Run as ``python -m oracle.make_golden`` from the repo root, in the
build container where /root/reference exists.  For each case the reference model is built
through its own factory (nets.get_model), loaded with ``synthetic_state(cfg, seed)``, run in
train mode (BatchNorm batch statistics, dropout p = 0) on ``synthetic_batch`` and the outputs
are stored: (pos, ori), per-parameter gradient L2 norms and leading entries for the loss
sum(pos^2)+sum(ori^2), the updated BN running statistics' sums, and an eval-mode forward.
Weights and inputs are NOT stored (they are regenerated from the seeds; an input checksum is
stored to detect RNG drift), which keeps the fixtures a few KB each.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import deeplio_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.configs import make_cfg  # noqa: E402

# name -> (make_cfg kwargs, B, S, H, W, T_imu, seed)
CASES = {
    "simple1_fc_fc": (dict(lidar="lidar-feat-simple-1", imu="imu-feat-fc", odom="odom-feat-fc"), 2, 2, 16, 64, 5, 11),
    "simple1_lstm_rnn": (dict(lidar="lidar-feat-simple-1", imu="imu-feat-rnn", odom="odom-feat-rnn", odom_hidden=64), 2, 3, 16, 64, 6, 12),
    "pointseg_lstm_rnn": (dict(lidar="lidar-feat-pointseg", imu="imu-feat-rnn", odom="odom-feat-rnn", odom_hidden=64), 2, 2, 16, 128, 5, 13),
    "flownet_sub_gru_gru": (dict(lidar="lidar-feat-flownet", imu="imu-feat-rnn", rnn_type="gru", odom="odom-feat-rnn",
                                 odom_rnn_type="gru", odom_hidden=64, lidar_fusion="sub"), 2, 2, 16, 128, 5, 14),
    "resnet_lstm_rnn": (dict(lidar="lidar-feat-resnet", imu="imu-feat-rnn", odom="odom-feat-rnn", odom_hidden=64), 2, 2, 16, 128, 7, 15),
    "imu_only_lstm": (dict(lidar=None, imu="imu-feat-rnn", odom="odom-feat-rnn", odom_hidden=64), 3, 3, 16, 64, 9, 16),
    "lidar_only_simple1": (dict(lidar="lidar-feat-simple-1", imu=None, odom="odom-feat-fc"), 2, 2, 16, 64, 5, 17),
    # BASELINE.json configs[4] shape: FlowNet + bi-LSTM over a 50-step IMU window
    "flownet_add_lstm_t50": (dict(lidar="lidar-feat-flownet", imu="imu-feat-rnn", odom="odom-feat-rnn", odom_hidden=64,
                                  lidar_fusion="add"), 2, 2, 16, 128, 50, 18),
    # odd image height, odd batch, widths that turn odd inside the net (ceil-mode pools with overhanging windows)
    "simple1_gru_fc_19x76": (dict(lidar="lidar-feat-simple-1", imu="imu-feat-rnn", rnn_type="gru", odom="odom-feat-fc"),
                             3, 2, 19, 76, 7, 19),
}
GRAD_HEAD = 6


def run_case(name):
    kw, B, S, H, W, T, seed = CASES[name]
    cfg = make_cfg(height=H, width=W, seq=S, **kw)
    model = ref_loader.build_reference_model(cfg, H, W, batch_size=B)
    sd = O.synthetic_state(cfg, seed=seed)
    model.load_state_dict(sd)
    ref_loader.apply_patch_p4(model)
    xyz, normals, imus = O.synthetic_batch(B, S, H, W, T, seed=seed)
    model.train()
    pos, ori = model([[xyz.clone(), normals.clone()], imus.clone()])
    loss = (pos ** 2).sum() + (ori ** 2).sum()
    loss.backward()
    grads = {}
    for k, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        grads[k] = (g.double().norm().float(), g.flatten()[:GRAD_HEAD].clone())
    running = {k: v.double().sum().float() for k, v in model.state_dict().items() if "running_" in k}
    model.eval()
    with torch.no_grad():
        epos, eori = model([[xyz.clone(), normals.clone()], imus.clone()])
    return {
        "case": name, "kwargs": kw, "B": B, "S": S, "H": H, "W": W, "T": T, "seed": seed,
        "input_checksum": torch.stack([xyz.double().sum(), normals.double().sum(), imus.double().sum()]).float(),
        "pos": pos.detach(), "ori": ori.detach(), "loss": loss.detach(),
        "grads": grads, "running": running, "eval_pos": epos, "eval_ori": eori,
        "torch": torch.__version__,
    }


def main():
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]          # optional: case names to (re)generate; default all
    for name in (only or CASES):
        rec = run_case(name)
        path = os.path.join(out_dir, name + ".pt")
        torch.save(rec, path)
        print(name, "pos[0,0]=", rec["pos"][0, 0].tolist(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
