"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/deeplio_b200.h declares, the host-side modules expose the reference's state_dict keys / shapes, and
the factory mirrors the reference's registry.  No compute calls (no GPU here)."""
import argparse
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from deeplio_b200.build import build_library
    build_library()
    from deeplio_b200 import _lib
    return _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "deeplio_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dlio_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    decl = declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if s not in exported]
    assert not missing, missing
    assert sorted(lib.EXPORTS) == decl          # the ctypes binding covers the whole header
    header = open(os.path.join(ROOT, "include", "deeplio_b200.h")).read()
    assert lib.abi_version() == lib.ABI_VERSION == int(re.search(r"#define DLIO_ABI_VERSION (\d+)", header).group(1))


def test_workspace_bytes_is_the_scratch_the_entry_points_take(lib):
    """Host-only query (no device call): the sizes the engine allocates for each op's scratch."""
    assert lib.workspace_bytes("conv2d_fwd", 128) == 2 * 128 * 8
    assert lib.workspace_bytes("bn_bwd", 64) == (2 * 64 + 1) * 8
    dims = (0, 2, 2, 8, 15, 6, 128)           # kind, layers, directions, B, T, I, H
    assert lib.workspace_bytes("rnn_fwd", *dims) == 4 * lib.rnn_reserve_floats(*dims) > 0
    assert lib.workspace_bytes("rnn_bwd", *dims) == 4 * lib.rnn_bwd_scratch_floats(*dims) > 0
    assert lib.workspace_bytes("scan_project", 64, 2048) == lib.scan_scratch_bytes(64, 2048) > 0
    with pytest.raises(lib.DlioError, match="unknown op"):
        lib.workspace_bytes("fire", 1)
    with pytest.raises(lib.DlioError, match="takes 2 dims"):
        lib.workspace_bytes("scan_project", 64)


def test_library_is_sm100a_with_no_other_arch(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.parametrize("kw", [dict(), dict(lidar="lidar-feat-pointseg"), dict(lidar="lidar-feat-resnet", rnn_type="gru"),
                                dict(lidar="lidar-feat-flownet", lidar_fusion="sub"), dict(imu="imu-feat-fc", odom="odom-feat-fc"),
                                dict(lidar="lidar-feat-resnet", lidar_fusion="cat"), dict(fusion="fusion-layer-cat"),
                                dict(lidar=None), dict(imu=None, odom="odom-feat-fc")])
def test_state_dict_keys_and_shapes_match_reference_layout(lib, kw):
    """Module construction is device-independent; only forward needs CUDA."""
    from deeplio_b200 import nets
    from deeplio_b200.config import build_config_container
    from oracle import deeplio_oracle as O
    from oracle.configs import make_cfg
    cfg = make_cfg(**kw)
    build_config_container(cfg, argparse.Namespace(device="cuda:0", batch_size=2))
    a = cfg["deeplio"]
    net = nets.DeepLIO((3, 64, 2048), cfg)
    lidar = nets.create_lidar_feat_net((3, 64, 2048), cfg, a, "cpu")
    imu = nets.create_imu_feat_net(cfg, a, "cpu")
    fusion = None
    if lidar is not None and imu is not None:
        fusion = nets.create_fusion_net([lidar.get_output_shape(), imu.get_output_shape()], cfg, a, "cpu")
    odom_in = (fusion or lidar or imu).get_output_shape()
    odom = nets.create_odometry_feat_net(odom_in, cfg, a, "cpu")
    net.lidar_feat_net, net.imu_feat_net, net.fusion_net, net.odom_feat_net = lidar, imu, fusion, odom
    net.initialize()
    sd = net.state_dict()
    shapes = O.state_shapes(cfg)            # pinned to the reference by tests/test_oracle_vs_reference.py
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    assert net.name == "deeplio"
    for sub in net.get_feat_networks():
        assert sub.name == type(sub).__name__.lower()
    if lidar is not None:
        assert list(lidar.get_output_shape()) == [1, 2, 128]


def test_pool_out_matches_torch():
    import torch.nn.functional as F
    from deeplio_b200.engine import pool_out
    for n in list(range(3, 40)) + [64, 65, 129, 257, 513, 1024, 2048]:
        for s in (1, 2):
            for ceil in (False, True):
                ref = F.max_pool2d(torch.zeros(1, 1, n, 8), 3, (s, 1), 1, ceil_mode=ceil).shape[2]
                assert pool_out(n, s, ceil) == ref


def test_install_rebinds_reference_registry():
    """The reference factory resolves class names from its module globals; install() swaps them."""
    import sys
    import types
    fake = types.ModuleType("fake_ref_nets")
    for n in ("DeepLIO", "LidarSimpleFeat1", "OdomFeatRNN"):
        setattr(fake, n, object)
    sys.modules["fake_ref_nets"] = fake
    from deeplio_b200 import nets
    from deeplio_b200.dropin import install
    done = install("fake_ref_nets")
    assert set(done) == {"DeepLIO", "LidarSimpleFeat1", "OdomFeatRNN"}
    assert fake.LidarSimpleFeat1 is nets.LidarSimpleFeat1


def test_config_container_required():
    from deeplio_b200 import config
    old = config._container
    config._container = None
    try:
        with pytest.raises(ValueError):
            config.get_config_container()
    finally:
        config._container = old
