"""Model factory: the drop-in for ``deeplio.models.nets`` (reference nets/__init__.py:16-238).

Same entry points (``get_model``, ``create_deeplio_arch``, ``create_{lidar,imu,fusion,odometry}_feat_net``,
``load_state_dict``, ``disable_grad``), same registry names, same ``ValueError`` for unknown names, same
``pretrained`` / ``requires-grad`` handling.  Differences: a missing ``fusion-net.pretrained`` key is read as
false (the shipped config.yaml lacks it and the reference raises KeyError), and
``fusion-layer-cat`` works (P2).
"""
import logging
import os

import torch

from .. import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .deeplio import DeepLIO
from .fusion import DeepLIOFusionCat, DeepLIOFusionSoft
from .imu import ImuFeatFC, ImufeatRNN0
from .lidar import LidarFlowNetFeat, LidarPointSegFeat, LidarResNetFeat, LidarSimpleFeat1
from .odom import OdomFeatFC, OdomFeatRNN

net_logger = logging.getLogger("deeplio_b200")

LIDAR_NETS = {"lidar-feat-pointseg": LidarPointSegFeat, "lidar-feat-flownet": LidarFlowNetFeat,
              "lidar-feat-resnet": LidarResNetFeat, "lidar-feat-simple-1": LidarSimpleFeat1}
IMU_NETS = {"imu-feat-fc": ImuFeatFC, "imu-feat-rnn": ImufeatRNN0}
FUSION_NETS = {"fusion-layer-cat": DeepLIOFusionCat, "fusion-layer-soft": DeepLIOFusionSoft}
ODOM_NETS = {"odom-feat-fc": OdomFeatFC, "odom-feat-rnn": OdomFeatRNN}


def get_model(input_shape, cfg, device):
    return create_deeplio_arch(input_shape, cfg, device)


def create_deeplio_arch(input_shape, cfg, device):
    if torch.device(device).type != "cuda":
        raise RuntimeError("deeplio_b200 runs on CUDA devices only (got device=%r); there is no CPU fallback" % (device,))
    arch_cfg = cfg["deeplio"]
    net = DeepLIO(input_shape, cfg)
    lidar = create_lidar_feat_net(input_shape, cfg, arch_cfg, device)
    imu = create_imu_feat_net(cfg, arch_cfg, device)
    fusion = None
    if lidar is not None and imu is not None:
        fusion = create_fusion_net([lidar.get_output_shape(), imu.get_output_shape()], cfg, arch_cfg, device)
    if fusion is not None:
        odom_in = fusion.get_output_shape()
    elif lidar is not None:
        odom_in = lidar.get_output_shape()
    elif imu is not None:
        odom_in = imu.get_output_shape()
    else:
        raise ValueError("No input-shape for odometry network is defined, please check you configuration!")
    odom = create_odometry_feat_net(odom_in, cfg, arch_cfg, device)
    net.lidar_feat_net, net.imu_feat_net, net.fusion_net, net.odom_feat_net = lidar, imu, fusion, odom
    net.initialize()
    net.to(device=device)
    if arch_cfg["pretrained"]:
        load_state_dict(net, arch_cfg["model-path"])
        net.pretrained = True
    return net


def _finish(feat_net, feat_cfg, device, loader=None):
    feat_net.to(device)
    if feat_cfg.get("pretrained", False):
        (loader or load_state_dict)(feat_net, feat_cfg["model-path"])
        feat_net.pretrained = True
    if not feat_cfg.get("requires-grad", True):
        disable_grad(feat_net)
    return feat_net


def _lookup(registry, feat_cfg, err):
    name = feat_cfg.get("name", None) if feat_cfg else None
    if name is None:
        return None, None
    name = name.lower()
    if name not in registry:
        raise ValueError(err.format(name))
    return name, registry[name]


def create_lidar_feat_net(input_shape, cfg, arch_cfg, device):
    feat_cfg = arch_cfg["lidar-feat-net"]
    name, cls = _lookup(LIDAR_NETS, feat_cfg, "Wrong feature network {}")
    if cls is None:
        return None
    net_logger.info("creating deeplio lidar feature net (%s).", name)

    def loader(net, path):  # PointSeg encoder-only checkpoints (nets/__init__.py:108-120)
        if name == "lidar-feat-pointseg" and "encoder" in path:
            load_state_dict(net.encoder1, path)
        else:
            load_state_dict(net, path)
    return _finish(cls(input_shape, cfg[name]), feat_cfg, device, loader)


def create_imu_feat_net(cfg, arch_cfg, device):
    feat_cfg = arch_cfg["imu-feat-net"]
    name, cls = _lookup(IMU_NETS, feat_cfg, "Wrong feature network {}")
    if cls is None:
        return None
    net_logger.info("creating deeplio imu feature net (%s).", name)
    return _finish(cls(cfg[name]), feat_cfg, device)


def create_fusion_net(input_shape, cfg, arch_cfg, device):
    feat_cfg = arch_cfg.get("fusion-net", None)
    name, cls = _lookup(FUSION_NETS, feat_cfg, "Wrong feature network {}")
    if cls is None:
        return None
    net_logger.info("creating deeplio fusion layer (%s).", name)
    return _finish(cls(input_shape, cfg[name]), feat_cfg, device)


def create_odometry_feat_net(input_shape, cfg, arch_cfg, device):
    feat_cfg = arch_cfg["odom-feat-net"]
    name, cls = _lookup(ODOM_NETS, feat_cfg, "Wrong odometry feature network {}")
    if cls is None:
        return None
    net_logger.info("creating deeplio odom feature net (%s).", name)
    feat_net = cls(input_shape[2], cfg[name])
    feat_net.to(device)
    if not feat_cfg.get("requires-grad", True):
        disable_grad(feat_net)
    return feat_net


def load_state_dict(module, model_path):
    net_logger.info("loading %s's state dict (%s).", getattr(module, "name", type(module).__name__), model_path)
    if not os.path.isfile(model_path):
        net_logger.error("%s: No model found (%s)!", getattr(module, "name", type(module).__name__), model_path)
    # a parameter-free module (DeepLIOFusionCat) has no tensor to take the device from
    first = next(iter(module.parameters()), None)
    if first is None:
        first = next(iter(module.buffers()), None)
    dev = first.device if first is not None else getattr(module, "device", "cpu")
    state = torch.load(model_path, map_location=dev)
    module.load_state_dict(state["state_dict"])


def disable_grad(module):
    for param in module.parameters():
        param.requires_grad = False
