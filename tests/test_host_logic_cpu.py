"""Host-side decisions of the executor that need no GPU: which producer / consumer pairs skip fp32 copies or use the
pixel-pair view, and that the named workloads agree with the oracle-side configuration table."""
import os
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


# ----------------------------------------------------------------------------- pose kernels' arithmetic on the host
def _pose_host_binary(tmp_path_factory):
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("pose_host") / "pose_host")
    src = os.path.join(ROOT, "tests", "native", "pose_host.cu")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", out, src], check=True,
                   capture_output=True)
    return out


@pytest.fixture(scope="module")
def pose_host(tmp_path_factory):
    return _pose_host_binary(tmp_path_factory)


def _run(binary, *chunks):
    import subprocess
    import numpy as np
    data = b"".join(np.ascontiguousarray(c).tobytes() for c in chunks)
    return subprocess.run([binary], input=data, capture_output=True, check=True).stdout


def test_pose_chain_math_matches_reference_glue(pose_host):
    """The per-sample functions the CUDA pose kernels are built from (csrc/pose_math.cuh), compiled for the host,
    against the reference-generated fixture (Trainer.se3_to_SE3 and its autograd gradients) and the oracle."""
    import numpy as np
    import torch
    from oracle import pose_oracle as P
    from tests.helpers import GOLDEN_DIR, quat_tol
    c = torch.load(os.path.join(GOLDEN_DIR, "pose_glue.pt"), weights_only=False)["chain"]
    B, S, _ = c["x"].shape
    out = _run(pose_host, np.int32([0]), np.int32([B, S]), c["x"].numpy(), c["w"].numpy(), c["gx"].numpy(), c["gq"].numpy())
    f = np.frombuffer(out[:-4], dtype=np.float32)
    status = int(np.frombuffer(out[-4:], dtype=np.int32)[0])
    n3, n4 = B * S * 3, B * S * 4
    ox, oq = torch.from_numpy(f[:n3].copy()).view(B, S, 3), torch.from_numpy(f[n3:n3 + n4].copy()).view(B, S, 4)
    dx = torch.from_numpy(f[n3 + n4:2 * n3 + n4].copy()).view(B, S, 3)
    dw = torch.from_numpy(f[2 * n3 + n4:].copy()).view(B, S, 3)
    assert status == 0
    assert torch.allclose(ox, c["f2g_x"], rtol=1e-6, atol=2e-6)
    assert ((oq - c["f2g_q"]).abs() <= quat_tol(c["f2g_q"])).all()
    well = (c["f2g_q"][:, :, 0].abs().min(dim=1).values > 0.05)      # samples without a near-half-turn pose
    assert well.sum() >= 2
    assert torch.allclose(dx[well], c["dx"][well], rtol=1e-5, atol=1e-5)
    assert torch.allclose(dw[well], c["dw"][well], rtol=1e-4, atol=1e-4)
    # the ill-conditioned samples against an fp64 run of the oracle (same branches), at a conditioning-aware bar
    x64, w64 = c["x"].double().requires_grad_(True), c["w"].double().requires_grad_(True)
    fx, fq, _ = P.se3_to_SE3(x64, w64)
    ((fx * c["gx"].double()).sum() + (fq * c["gq"].double()).sum()).backward()
    amp = 1.0 / (4.0 * c["f2g_q"][:, :, 0].abs().clamp_min(1e-3)).min(dim=1).values ** 2     # d(q)/dR ~ 1 / (4 qw)^2
    for b in range(B):
        bar = 2e-5 * (1.0 + float(amp[b])) * (1.0 + float(w64.grad[b].abs().max()))
        assert (dw[b].double() - w64.grad[b]).abs().max().item() <= bar, (b, bar)
        assert (dx[b].double() - x64.grad[b]).abs().max().item() <= 1e-5 * (1.0 + float(x64.grad[b].abs().max()))


def test_pose_chain_math_flags_and_long_chains(pose_host):
    """Status flags instead of the reference's exceptions; a 64-pair chain stays a rotation (projection kicks in)."""
    import numpy as np
    import torch
    from oracle import pose_oracle as P
    g = torch.Generator().manual_seed(1)
    B, S = 3, 64
    x, w = torch.randn(B, S, 3, generator=g) * 0.3, torch.randn(B, S, 3, generator=g) * 0.4
    z3, z4 = np.zeros((B, S, 3), np.float32), np.zeros((B, S, 4), np.float32)
    out = _run(pose_host, np.int32([0]), np.int32([B, S]), x.numpy(), w.numpy(), z3, z4)
    f = np.frombuffer(out[:-4], dtype=np.float32)
    oq = torch.from_numpy(f[B * S * 3:B * S * 7].copy()).view(B, S, 4)
    assert int(np.frombuffer(out[-4:], dtype=np.int32)[0]) == 0
    fx, fq, _ = P.se3_to_SE3(x.double(), w.double())
    sign = torch.sign((oq.double() * fq).sum(dim=2, keepdim=True))          # q and -q are the same rotation
    assert ((oq.double() * sign - fq).abs() <= 4 * quat_tol_f64(fq)).all()
    x[1, 5, 0] = float("nan")
    out = _run(pose_host, np.int32([0]), np.int32([B, S]), x.numpy(), w.numpy(), z3, z4)
    assert int(np.frombuffer(out[-4:], dtype=np.int32)[0]) & 1


def quat_tol_f64(q):
    from tests.helpers import quat_tol
    return quat_tol(q.float()).double() * 8       # 64 chained fp32 products against fp64


def test_ground_truth_math_matches_reference_glue(pose_host):
    import numpy as np
    import torch
    from tests.helpers import GOLDEN_DIR
    g = torch.load(os.path.join(GOLDEN_DIR, "pose_glue.pt"), weights_only=False)["gt"]
    gts, comb = g["gts"], np.int32(g["combinations"])
    B, F, _ = gts.shape
    S = comb.shape[0]
    out = _run(pose_host, np.int32([1]), np.int32([B, F, S]), comb, gts.numpy())
    f = np.frombuffer(out[:-4], dtype=np.float32)
    f2f = torch.from_numpy(f[:B * S * 6].copy()).view(B, S, 6)
    f2g = torch.from_numpy(f[B * S * 6:].copy()).view(B, S, 7)
    assert int(np.frombuffer(out[-4:], dtype=np.int32)[0]) == 0
    # translations: the reference forms R^T t_j - R^T t_i in fp32 (spatial.py:904-923 + a 4x4 matmul), the kernel
    # R^T (t_j - t_i): both within |t| * 2^-22 of the exact value
    tmax = gts[:, :, 0:3].abs().max().item()
    assert (f2f[:, :, 0:3] - g["f2f"][:, :, 0:3]).abs().max().item() <= 4e-7 * tmax + 1e-6
    assert (f2g[:, :, 0:3] - g["f2g"][:, :, 0:3]).abs().max().item() <= 4e-7 * tmax + 1e-6
    assert torch.allclose(f2f[:, :, 3:], g["f2f"][:, :, 3:], rtol=1e-5, atol=2e-6)
    assert torch.allclose(f2g[:, :, 3:], g["f2g"][:, :, 3:], rtol=1e-5, atol=2e-6)
