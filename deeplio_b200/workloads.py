"""Named workloads: config.yaml-shaped dicts for the BASELINE.json configurations (reference config.yaml:1-132
for the structure and default hyper-parameters).  Used by bench.py and the examples; tests check that these
agree with the oracle-side copies."""
import copy
import math

DEFAULT_CONFIG = {
    "datasets": {
        "sequence-size": 3,
        "combinations": [[0, 1], [1, 2]],
        "kitti": {"image-width": 2048, "image-height": 64,
                  "mean-image": [-0.0014, 0.0043, -0.011, 0.2258, -0.0024, 0.0037, 0.3793, 0.1115],
                  "std-image": [0.1269, 0.0951, 0.0108, 0.1758, 0.3436, 0.4445, 0.5664, 0.0884]},
    },
    "deeplio": {
        "dropout": 0.25, "pretrained": False, "model-path": "",
        "lidar-feat-net": {"name": "lidar-feat-simple-1", "pretrained": False, "model-path": "", "requires-grad": True},
        "imu-feat-net": {"name": "imu-feat-rnn", "pretrained": False, "model-path": "", "requires-grad": True},
        "odom-feat-net": {"name": "odom-feat-rnn", "pretrained": False, "model-path": "", "requires-grad": True},
        "fusion-net": {"name": "fusion-layer-soft", "requires-grad": True, "pretrained": False},
    },
    "lidar-feat-pointseg": {"dropout": 0.1, "classes": ["unknown", "object"], "bypass": "simple",
                            "fusion": "add", "part": "encoder"},
    "lidar-feat-flownet": {"dropout": 0.0, "fusion": "add"},
    "lidar-feat-resnet": {"dropout": 0.25, "fusion": "add"},
    "lidar-feat-simple-1": {"dropout": 0.25, "fusion": "add", "bypass": False},
    "imu-feat-fc": {"input-size": 6, "hidden-size": [128, 256, 512, 512, 256, 128], "dropout": 0.0},
    "imu-feat-rnn": {"type": "lstm", "input-size": 6, "hidden-size": 128, "num-layers": 2,
                     "bidirectional": True, "dropout": 0.1},
    "fusion-layer-cat": {"type": "cat"},
    "fusion-layer-soft": {"type": "soft"},
    "odom-feat-fc": {"size": [1024, 512, 256], "dropout": 0.0},
    "odom-feat-rnn": {"type": "lstm", "hidden-size": 1024, "num-layers": 2, "bidirectional": True, "dropout": 0.0},
    "losses": {"active": "hwsloss", "hwsloss": {"params": {"learn": True, "sx": 0.0, "sq": -3.0}},
               "lwsloss": {"params": {"beta": 1125.0}}, "loss-type": "local+global"},
    "current-dataset": "kitti",
    "channels": [0, 1, 2, 4, 5, 6],
    "optimizer": "adam",
}

# name -> (lidar net, lidar fusion, imu net, imu rnn type, odom net, per-sample pairs S, IMU window, batch)
WORKLOADS = {
    "cfg0_simple1_fc_b1": ("lidar-feat-simple-1", "add", "imu-feat-fc", "lstm", "odom-feat-fc", 2, 15, 1),
    "cfg1_simple1_lstm_b8": ("lidar-feat-simple-1", "add", "imu-feat-rnn", "lstm", "odom-feat-rnn", 2, 15, 8),
    "cfg2_pointseg_lstm_b32": ("lidar-feat-pointseg", "add", "imu-feat-rnn", "lstm", "odom-feat-rnn", 2, 15, 32),
    "cfg3_resnet_gru_b64": ("lidar-feat-resnet", "cat", "imu-feat-rnn", "gru", "odom-feat-rnn", 2, 15, 64),
    "cfg4_flownet_lstm_t50_b16": ("lidar-feat-flownet", "add", "imu-feat-rnn", "lstm", "odom-feat-rnn", 2, 50, 16),
}


def workload_config(name, height=64, width=2048):
    """Returns (cfg dict, batch, S, T_imu) for a BASELINE.json workload, dropout as the reference ships it."""
    lidar, lfusion, imu, rnn_type, odom, seq, t_imu, batch = WORKLOADS[name]
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["datasets"]["sequence-size"] = seq + 1
    cfg["datasets"]["combinations"] = [[i, i + 1] for i in range(seq)]
    cfg["datasets"]["kitti"]["image-width"] = width
    cfg["datasets"]["kitti"]["image-height"] = height
    a = cfg["deeplio"]
    a["lidar-feat-net"]["name"] = lidar
    a["imu-feat-net"]["name"] = imu
    a["odom-feat-net"]["name"] = odom
    cfg[lidar]["fusion"] = lfusion
    cfg["imu-feat-rnn"]["type"] = rnn_type
    return cfg, batch, seq, t_imu


def synthetic_gts(batch, frames, seed=0):
    """Synthetic ground-truth poses [batch, frames, 15] = (t 3, R 9 row-major, v 3) per frame, the layout of
    ``Kitti.load_ground_truth`` (deeplio/datasets/kitti.py:292-301): smooth yaw-dominated forward motion from a
    random start pose (float64 arithmetic, returned as float32 torch tensor)."""
    import numpy as np
    import torch

    def rodrigues(w):
        th = float(np.linalg.norm(w))
        if th < 1e-12:
            return np.eye(3)
        a = w / th
        K = np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])
        return np.eye(3) + math.sin(th) * K + (1.0 - math.cos(th)) * (K @ K)
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, frames, 15))
    for b in range(batch):
        R = rodrigues(rng.standard_normal(3) * 0.5)
        t = rng.standard_normal(3) * 10.0
        for f in range(frames):
            w = np.array([0.002, 0.004, 0.03]) * (1.0 + rng.standard_normal(3))
            v = np.array([1.2, 0.02, 0.01]) * (1.0 + 0.3 * rng.standard_normal(3))
            out[b, f, 0:3], out[b, f, 3:12], out[b, f, 12:15] = t, R.reshape(9), v
            t = t + R @ v
            R = R @ rodrigues(w)
    return torch.from_numpy(out).float()
