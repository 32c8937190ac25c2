"""Top-level module: drop-in for ``DeepLIO`` (deeplio_nets.py:26-100)."""
import logging

from torch import nn

from .. import functional as Fn
from ..config import get_config_container
from .base import BaseNet


class DeepLIO(BaseNet):
    """lidar_feat_net -> imu_feat_net -> fusion_net -> odom_feat_net -> dropout -> fc_pos / fc_ori."""

    def __init__(self, input_shape, cfg, bn_d=0.1):
        super().__init__()
        self.cfg_container = get_config_container()
        self.seq_size = self.cfg_container.seq_size
        self.combinations = self.cfg_container.combinations
        self.logger = logging.getLogger("deeplio_b200")
        self.cfg = cfg["deeplio"]
        self.p = self.cfg.get("dropout", 0.0)
        self.input_shape = input_shape
        self.lidar_feat_net = None
        self.imu_feat_net = None
        self.fusion_net = None
        self.odom_feat_net = None
        self.drop = None
        self.fc_pos = None
        self.fc_ori = None
        # optional callable, invoked during backward when the gradients of everything downstream of the feature nets
        # (fusion, odometry net, heads) are complete (deeplio_b200.parallel.OverlappedGradReducer)
        self.on_head_grads_ready = None

    def initialize(self):
        last = next(n for n in (self.odom_feat_net, self.fusion_net, self.imu_feat_net, self.lidar_feat_net)
                    if n is not None)
        width = last.get_output_shape()[2]  # [B, S, N]
        if self.p > 0:
            self.drop = nn.Dropout(self.p)  # kept for module-tree parity; the mask comes from the C ABI
        self.fc_pos = nn.Linear(width, 3)
        self.fc_ori = nn.Linear(width, 3)

    def forward(self, x):
        lidar_imgs, imu_meas = x[0], x[1]
        last = lidar = imu = None
        if self.lidar_feat_net is not None:
            last = lidar = self.lidar_feat_net(lidar_imgs)
        if self.imu_feat_net is not None:
            last = imu = self.imu_feat_net(imu_meas)
        feat = lidar if lidar is not None else imu
        if self.on_head_grads_ready is not None and feat is not None and feat.requires_grad:
            feat.register_hook(self._head_grads_hook)
        if self.fusion_net is not None:
            last = self.fusion_net([lidar, imu])
        if self.odom_feat_net is not None:
            last = self.odom_feat_net(last)
        last = Fn.dropout(last, self.p, self.training)
        return (Fn.linear(last, self.fc_pos.weight, self.fc_pos.bias),
                Fn.linear(last, self.fc_ori.weight, self.fc_ori.bias))

    def _head_grads_hook(self, grad):
        if self.on_head_grads_ready is not None:
            self.on_head_grads_ready()
        return None

    def get_feat_networks(self):
        nets = []
        for net in (self.odom_feat_net, self.fusion_net, self.imu_feat_net, self.lidar_feat_net):
            if net is not None and isinstance(net, nn.Module):
                nets.extend(net.get_modules())
        return nets
