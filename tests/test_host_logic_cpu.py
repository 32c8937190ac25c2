"""Host-side decisions of the executor that need no GPU: which producer / consumer pairs skip fp32 copies or use the
pixel-pair view, and that the named workloads agree with the oracle-side configuration table."""
import types

import pytest


def _engine():
    from deeplio_b200 import engine as E
    return E


def _act(**kw):
    base = dict(h2=object(), group=2, c=64, w=128, pw=2, ph=1, h=16)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_pair_view_applicability():
    E = _engine()
    assert E.pair_ok(_act(), 64, 128, 3, 5, (1, 2))              # FlowNet conv2 / conv3
    assert E.pair_ok(_act(), 64, 128, 3, 3, (2, 2))              # H stride by row decimation
    assert E.pair_ok(_act(pw=0), 64, 128, 1, 1, (2, 2))          # 1x1 downsample needs no column pads
    assert not E.pair_ok(_act(group=1), 64, 128, 3, 3, (1, 2))   # planes not in the pixel-pair layout
    assert not E.pair_ok(_act(w=127), 64, 128, 3, 3, (1, 2))     # odd width: no [w/2, 2c] view
    assert not E.pair_ok(_act(pw=1), 64, 128, 3, 3, (1, 2))      # odd row pads
    assert not E.pair_ok(_act(), 64, 128, 3, 7, (1, 2))          # kw = 7 is the first layer's space-to-depth path
    assert not E.pair_ok(_act(), 64, 48, 3, 3, (1, 2))           # Cout % 64 (wgrad tile)
    assert not E.pair_ok(_act(), 64, 128, 3, 3, (1, 1))          # not strided
    assert not E.pair_ok(_act(c=24), 24, 128, 3, 3, (1, 2))      # 2 * Cin must fill 64-half K chunks


def test_consumer_reads_f16_only():
    E = _engine()
    assert E.consumer_reads_f16_only(128, 256, 3, (1, 1), 257)
    assert E.consumer_reads_f16_only(64, 128, (3, 5), (1, 2), 1024)
    assert not E.consumer_reads_f16_only(64, 128, (3, 5), (1, 2), 45)      # odd width -> CUDA-core path reads fp32
    assert not E.consumer_reads_f16_only(96, 128, 3, (1, 1), 64)           # Cin % 64
    assert not E.consumer_reads_f16_only(128, 80, 3, (1, 1), 64)           # Cout % 64 (wgrad on the CUDA cores)


@pytest.mark.parametrize("name", ["cfg0_simple1_fc_b1", "cfg1_simple1_lstm_b8", "cfg2_pointseg_lstm_b32",
                                  "cfg3_resnet_gru_b64", "cfg4_flownet_lstm_t50_b16"])
def test_workloads_agree_with_oracle_configs(name):
    from deeplio_b200.workloads import workload_config
    from oracle.configs import BASELINE_CONFIGS, make_cfg
    cfg, batch, seq, t_imu = workload_config(name)
    kw, obatch, oseq, ot = BASELINE_CONFIGS[name]
    ocfg = make_cfg(no_dropout=False, **kw)
    assert (batch, seq, t_imu) == (obatch, oseq, ot)
    for key in ("lidar-feat-net", "imu-feat-net", "odom-feat-net", "fusion-net"):
        assert cfg["deeplio"][key]["name"] == ocfg["deeplio"][key]["name"], key
    lidar = cfg["deeplio"]["lidar-feat-net"]["name"]
    assert cfg[lidar]["fusion"] == ocfg[lidar]["fusion"]
    assert cfg["imu-feat-rnn"]["type"] == ocfg["imu-feat-rnn"]["type"]
