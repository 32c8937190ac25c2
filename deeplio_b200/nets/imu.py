"""IMU feature nets: drop-ins for ``ImuFeatFC`` (imu_feat_nets.py:21-53) and ``ImufeatRNN0`` (:56-83)."""
import torch
from torch import nn

from .. import functional as Fn
from ..config import get_config_container
from .base import BaseNet, require_cuda


class BaseImuFeatNet(BaseNet):
    def __init__(self, cfg):
        super().__init__()
        self.p = cfg["dropout"]
        self.input_size = cfg["input-size"]
        self.num_layers = cfg.get("num-layers", 2)
        self.cfg_container = get_config_container()
        self.seq_size = self.cfg_container.seq_size
        self.combinations = self.cfg_container.combinations


class ImuFeatFC(BaseImuFeatNet):
    """Per IMU sample: a stack of leaky_relu(Linear) layers; the T samples of a window are summed.

    The reference loops over (b, s) in Python with 2*len(net) launches each (imu_feat_nets.py:43-50);
    here all B*S*T samples go through each layer in one launch."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.hidden_size = cfg.get("hidden-size", [6, 6])
        self.num_layers = len(self.hidden_size)
        sizes = [self.input_size] + list(self.hidden_size)
        self.net = nn.ModuleList([nn.Linear(sizes[i], sizes[i + 1]) for i in range(self.num_layers)])
        self.output_shape = [1, self.seq_size, self.hidden_size[-1]]

    def forward(self, x):
        if not torch.is_tensor(x):  # the reference also accepts nested lists of [T, 6] tensors
            x = torch.stack([torch.stack(list(xs)) for xs in x])
        require_cuda(x, "imu input")
        b, s, t, n = x.shape
        y = x.reshape(b * s * t, n)
        for m in self.net:
            y = Fn.linear(y, m.weight, m.bias, "leaky_relu")
        y = Fn.dropout(y, self.p, self.training)
        return Fn.sum_mid(y.view(b * s, t, -1)).view(b, self.seq_size, -1)


class ImufeatRNN0(BaseImuFeatNet):
    """(bi)LSTM / GRU over each IMU window; the hidden / cell state of every (layer, direction) slot is
    carried from window s to s+1; the feature is the last layer's forward-direction output at the last
    time step (imu_feat_nets.py:75-83)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.rnn_type = "gru" if cfg["type"].lower() == "gru" else "lstm"
        self.hidden_size = cfg.get("hidden-size", 6)
        self.bidirectional = cfg.get("bidirectional", False)
        cls = nn.GRU if self.rnn_type == "gru" else nn.LSTM
        self.rnn = cls(input_size=self.input_size, hidden_size=self.hidden_size, num_layers=self.num_layers,
                       bidirectional=self.bidirectional, dropout=self.p, batch_first=True)  # parameters only
        self.num_dir = 2 if self.bidirectional else 1
        self.output_shape = [1, self.seq_size, self.hidden_size]

    def forward(self, x):
        require_cuda(x, "imu input")
        b, s, t, n = x.shape
        weights = list(self.rnn._flat_weights)
        state = None
        feats = []
        for seq in range(s):
            out, state = Fn.rnn(x[:, seq], state, self.rnn_type, self.num_layers, self.bidirectional,
                                self.hidden_size, weights, self.p, self.training)
            feats.append(out[:, -1, :self.hidden_size])
        return Fn.stack_mid(feats)
