"""Config dicts with the structure of the reference's config.yaml (config.yaml:1-132).

TEST INFRASTRUCTURE.  ``make_cfg`` builds the dict a user would get from
``yaml.safe_load(open('config.yaml'))`` with the net names / hyper-parameters swapped;
the BASELINE.json configs are named here so tests and bench agree on them.
"""
import copy

_BASE = {
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


def make_cfg(lidar="lidar-feat-simple-1", imu="imu-feat-rnn", fusion="fusion-layer-soft", odom="odom-feat-rnn",
             seq=2, no_dropout=True, lidar_fusion=None, rnn_type=None, odom_hidden=None, odom_rnn_type=None,
             height=64, width=2048):
    cfg = copy.deepcopy(_BASE)
    cfg["datasets"]["sequence-size"] = seq + 1
    cfg["datasets"]["combinations"] = [[i, i + 1] for i in range(seq)]
    cfg["datasets"]["kitti"]["image-width"] = width
    cfg["datasets"]["kitti"]["image-height"] = height
    a = cfg["deeplio"]
    a["lidar-feat-net"]["name"] = lidar
    a["imu-feat-net"]["name"] = imu
    a["fusion-net"]["name"] = fusion
    a["odom-feat-net"]["name"] = odom
    if lidar_fusion and lidar:
        cfg[lidar]["fusion"] = lidar_fusion
    if rnn_type:
        cfg["imu-feat-rnn"]["type"] = rnn_type
    if odom_rnn_type:
        cfg["odom-feat-rnn"]["type"] = odom_rnn_type
    if odom_hidden:
        cfg["odom-feat-rnn"]["hidden-size"] = odom_hidden
    if no_dropout:  # parity runs: dropout RNG cannot match across implementations
        a["dropout"] = 0.0
        for k in ("lidar-feat-pointseg", "lidar-feat-flownet", "lidar-feat-resnet", "lidar-feat-simple-1",
                  "imu-feat-fc", "imu-feat-rnn", "odom-feat-fc", "odom-feat-rnn"):
            cfg[k]["dropout"] = 0.0
    return cfg


# BASELINE.json configs[0..4] as (cfg kwargs, B, S, T_imu)
BASELINE_CONFIGS = {
    "cfg0_simple1_fc_b1": (dict(lidar="lidar-feat-simple-1", imu="imu-feat-fc", odom="odom-feat-fc", seq=2), 1, 2, 15),
    "cfg1_simple1_lstm_b8": (dict(lidar="lidar-feat-simple-1", imu="imu-feat-rnn", odom="odom-feat-rnn", seq=2), 8, 2, 15),
    "cfg2_pointseg_lstm_b32": (dict(lidar="lidar-feat-pointseg", imu="imu-feat-rnn", odom="odom-feat-rnn", seq=2), 32, 2, 15),
    "cfg3_resnet_gru_b64": (dict(lidar="lidar-feat-resnet", imu="imu-feat-rnn", rnn_type="gru", odom="odom-feat-rnn",
                                 lidar_fusion="cat", seq=2), 64, 2, 15),
    "cfg4_flownet_lstm_t50_b16": (dict(lidar="lidar-feat-flownet", imu="imu-feat-rnn", odom="odom-feat-rnn", seq=2), 16, 2, 50),
}
