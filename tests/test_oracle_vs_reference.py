"""Direct oracle-vs-reference check (container only: needs /root/reference; skipped on the GPU box)."""
import pytest
import torch

from oracle import deeplio_oracle as O
from oracle import ref_loader
from oracle.configs import make_cfg
from tests.helpers import oracle_train_step, rel_err

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")

CASES = [
    (dict(lidar="lidar-feat-simple-1", imu="imu-feat-rnn", odom="odom-feat-rnn", odom_hidden=32), 2, 2, 24, 96, 4),
    (dict(lidar="lidar-feat-pointseg", imu="imu-feat-fc", odom="odom-feat-fc", lidar_fusion="sub"), 1, 2, 16, 128, 3),
    (dict(lidar="lidar-feat-resnet", imu="imu-feat-rnn", rnn_type="gru", odom="odom-feat-fc"), 2, 1, 16, 64, 3),
    (dict(lidar="lidar-feat-flownet", imu="imu-feat-rnn", odom="odom-feat-rnn", odom_hidden=32), 2, 2, 16, 128, 3),
]


@pytest.mark.parametrize("kw,B,S,H,W,T", CASES)
def test_oracle_equals_reference(kw, B, S, H, W, T, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    cfg = make_cfg(height=H, width=W, seq=S, **kw)
    model = ref_loader.build_reference_model(cfg, H, W, batch_size=B)
    shapes = O.state_shapes(cfg)
    ref_sd = model.state_dict()
    assert set(shapes) == set(ref_sd)
    for k, shp in shapes.items():
        assert tuple(ref_sd[k].shape) == tuple(shp), k
    sd = O.synthetic_state(cfg, seed=5)
    model.load_state_dict(sd)
    ref_loader.apply_patch_p4(model)
    model.train()
    inputs = O.synthetic_batch(B, S, H, W, T, seed=6)
    pos, ori = model([[inputs[0].clone(), inputs[1].clone()], inputs[2].clone()])
    ((pos ** 2).sum() + (ori ** 2).sum()).backward()
    opos, oori, grads, sd_after = oracle_train_step(cfg, sd, inputs)
    assert rel_err(opos, pos.detach()) < 2e-6
    assert rel_err(oori, ori.detach()) < 2e-6
    gmax = max(p.grad.abs().max().item() for p in model.parameters() if p.grad is not None)
    for k, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert (g - grads[k]).abs().max().item() <= 2e-4 * g.abs().max().item() + 1e-5 * gmax, k
    for k, v in model.state_dict().items():
        if "running_" in k:
            assert torch.allclose(v, sd_after[k], rtol=1e-5, atol=1e-6), k


def test_install_into_the_real_reference_factory(tmp_path, monkeypatch):
    """``deeplio_b200.install()`` rebinds the class names the UNMODIFIED reference factory resolves from its module
    globals (nets/__init__.py:6-10,95-102,141-144,173-176,206-209); the reference's own ``nets.get_model`` then
    builds deeplio_b200 modules whose state_dict keys and shapes are the reference's (checkpoint compatibility)."""
    import argparse
    monkeypatch.chdir(tmp_path)
    import deeplio_b200
    from deeplio_b200 import nets as ours
    from deeplio_b200.dropin import _NAMES
    ref_nets, ref_misc = ref_loader.import_reference()
    cfg = ref_loader.patch_cfg(make_cfg(height=16, width=64, seq=2, lidar="lidar-feat-simple-1", imu="imu-feat-rnn",
                                        odom="odom-feat-rnn", odom_hidden=32))
    ref_keys = {k: tuple(v.shape) for k, v in ref_loader.build_reference_model(cfg, 16, 64).state_dict().items()}
    saved = {n: getattr(ref_nets, n) for n in _NAMES if hasattr(ref_nets, n)}
    try:
        done = deeplio_b200.install()
        assert set(done) == set(saved)
        ref_misc.build_config_container(cfg, argparse.Namespace(device="cpu", batch_size=1))
        model = ref_nets.get_model(input_shape=(3, 16, 64), cfg=cfg, device="cpu")    # the reference's factory code
        assert type(model) is ours.DeepLIO
        assert type(model.lidar_feat_net) is ours.LidarSimpleFeat1
        assert type(model.imu_feat_net) is ours.ImufeatRNN0
        assert type(model.fusion_net) is ours.DeepLIOFusionSoft
        assert type(model.odom_feat_net) is ours.OdomFeatRNN
        got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert got == ref_keys
        # what trainer.py:60-61,156,165-170 reads
        assert model.name == "deeplio" and [n.name for n in model.get_feat_networks()] == [
            "odomfeatrnn", "deepliofusionsoft", "imufeatrnn0", "lidarsimplefeat1"]
        # a reference checkpoint loads into it
        model.load_state_dict(O.synthetic_state(cfg, seed=3))
    finally:
        for n, cls in saved.items():
            setattr(ref_nets, n, cls)
