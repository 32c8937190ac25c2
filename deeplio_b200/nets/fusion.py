"""Feature fusion: drop-ins for ``DeepLIOFusionCat`` (fusion_nets.py:9-37) and ``DeepLIOFusionSoft`` (:40-78)."""
import torch
from torch import nn

from .. import functional as Fn
from ..config import get_config_container
from .base import BaseNet


class DeepLIOFusionCat(BaseNet):
    """Plain concatenation.  The reference class is not an nn.Module, so its own factory fails on ``.to(device)``
    (nets/__init__.py:180); this one is a (parameter-free) module and works there."""

    def __init__(self, input_shapes, cfg):
        super().__init__()
        self.cfg_container = get_config_container()
        self.seq_size = self.cfg_container.seq_size
        self.combinations = self.cfg_container.combinations
        self.type = cfg.get("type", "cat").lower()
        self.input_shapes = input_shapes
        self.output_shape = [1, self.seq_size, sum(s[-1] for s in input_shapes)]

    @property
    def device(self):
        return torch.device(self.cfg_container.device)

    def forward(self, x):
        if self.type != "cat":
            raise NotImplementedError()
        return Fn.cat_last(x[0], x[1])


class DeepLIOFusionSoft(BaseNet):
    """Soft gating: each modality is scaled by sigmoid(Linear(cat(lidar, imu))) and the results concatenated.
    Computed out of place (the reference multiplies in place, which breaks autograd for ResNet / FlowNet,
    SURVEY.md 8c P4); values are identical."""

    def __init__(self, input_shapes, cfg):
        super().__init__()
        self.cfg_container = get_config_container()
        self.seq_size = self.cfg_container.seq_size
        self.combinations = self.cfg_container.combinations
        self.input_shapes = input_shapes
        self.s1_feat = None
        self.s2_feat = None
        total = sum(s[-1] for s in input_shapes)
        self.layers = nn.ModuleList([nn.Linear(total, s[-1]) for s in input_shapes])
        self.output_shape = [1, self.seq_size, total]

    def forward(self, x):
        lidar_feat, imu_feat = x[0], x[1]
        cat_feat = Fn.cat_last(lidar_feat, imu_feat)
        self.s1_feat = Fn.linear(cat_feat, self.layers[0].weight, self.layers[0].bias, "sigmoid")
        self.s2_feat = Fn.linear(cat_feat, self.layers[1].weight, self.layers[1].bias, "sigmoid")
        return Fn.cat_last(Fn.mul(lidar_feat.contiguous(), self.s1_feat), Fn.mul(imu_feat.contiguous(), self.s2_feat))
