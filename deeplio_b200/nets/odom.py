"""Odometry feature nets: drop-ins for ``OdomFeatFC`` (odom_feat_nets.py:8-45) and ``OdomFeatRNN`` (:48-86)."""
from torch import nn

from .. import functional as Fn
from ..config import get_config_container
from .base import BaseNet, require_cuda


class OdomFeatFC(BaseNet):
    def __init__(self, in_features, cfg):
        super().__init__()
        self.input_size = in_features
        # the reference reads 'hidden-size' although config.yaml spells the key 'size' (odom_feat_nets.py:12 vs
        # config.yaml:105-107), so the default [256, 128] is what actually runs; kept for parity
        self.hidden_size = cfg.get("hidden-size", [256, 128])
        self.p = cfg.get("dropout", 0.0)
        self.cfg_container = get_config_container()
        self.seq_size = self.cfg_container.seq_size
        self.combinations = self.cfg_container.combinations
        sizes = [self.input_size] + list(self.hidden_size)
        self.layers = nn.ModuleList([nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(self.hidden_size))])

    def forward(self, x):
        require_cuda(x, "odometry-net input")
        b, s, n = x.shape
        y = x.reshape(b * s, n)
        for layer in self.layers:
            y = Fn.linear(y, layer.weight, layer.bias, "leaky_relu")
        y = Fn.dropout(y, self.p, self.training)
        return y.view(b, s, -1)

    def get_output_shape(self):
        return [1, 1, self.hidden_size[-1]]


class OdomFeatRNN(BaseNet):
    def __init__(self, in_features, cfg):
        super().__init__()
        self.rnn_type = "gru" if cfg["type"].lower() == "gru" else "lstm"
        self.num_layers = cfg.get("num-layers", 2)
        self.hidden_size = cfg.get("hidden-size", 6)
        self.p = cfg.get("dropout", 0.0)
        self.bidirectional = cfg.get("bidirectional", False)
        self.input_size = in_features
        self.cfg_container = get_config_container()
        self.seq_size = self.cfg_container.seq_size
        self.combinations = self.cfg_container.combinations
        cls = nn.GRU if self.rnn_type == "gru" else nn.LSTM
        self.rnn = cls(input_size=self.input_size, hidden_size=self.hidden_size, num_layers=self.num_layers,
                       bidirectional=self.bidirectional, batch_first=True, dropout=self.p)  # parameters only
        self.num_dir = 2 if self.bidirectional else 1

    def forward(self, x):
        require_cuda(x, "odometry-net input")
        out, _ = Fn.rnn(x, None, self.rnn_type, self.num_layers, self.bidirectional, self.hidden_size,
                        list(self.rnn._flat_weights), self.p, self.training)
        return out[:, :, :self.hidden_size]  # forward direction of the last layer

    def get_output_shape(self):
        return [1, 1, self.hidden_size]
