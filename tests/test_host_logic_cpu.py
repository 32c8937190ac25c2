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


# ----------------------------------------------------------------------------- scan kernels' arithmetic on the host
@pytest.fixture(scope="module")
def scan_host(tmp_path_factory):
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("scan_host") / "scan_host")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", out,
                    os.path.join(ROOT, "tests", "native", "scan_host.cu")], check=True, capture_output=True)
    return out


def check_scan_image(img, idx, g, exact_pixels=0.999):
    """img [H,W,8], idx [H,W] (index into the unfiltered scan, -1 empty) against the reference-generated fixture.
    Pixel assignment is integer work done in float32: identical wherever atan2 / asin round identically (everywhere
    on the host; all but a few 1e-4 of the points on the GPU, whose libm differs in the last ulp)."""
    import numpy as np
    from oracle import scan_oracle as S
    ref, H, W = g["image"], int(g["H"]), int(g["W"])
    keep = S.depth_filter(g["scan"], float(g["min_depth"]), float(g["max_depth"]))
    orig = np.flatnonzero(keep)
    filled = g["proj_range"] > 0
    ref_idx = np.where(filled, orig[g["proj_idx"]], -1)
    same = ref_idx == idx
    assert same.mean() >= exact_pixels, same.mean()
    for c in (0, 1, 2, 3, 7):                      # xyz / max_depth, remission, range: copies of the winning point
        assert np.array_equal(img[..., c][same], ref[..., c][same]), c
    # normals: well-conditioned ones (the four weighted differences do not cancel) agree to 1e-4; the reference's
    # n / (|n| + 1e-8) is round-off noise where |n| ~ 1e-8 (isolated points), there only the magnitude is bounded
    nb = same.copy()
    nb[1:] &= same[:-1]; nb[:-1] &= same[1:]; nb[:, 1:] &= same[:, :-1]; nb[:, :-1] &= same[:, 1:]
    rn = np.linalg.norm(ref[..., 4:7], axis=2)
    well = nb & (rn > 0.9)
    assert well.sum() > 0.3 * H * W
    assert np.abs(img[..., 4:7] - ref[..., 4:7])[well].max() < 1e-4
    assert (np.linalg.norm(img[..., 4:7], axis=2) <= 1.0 + 1e-5).all()
    border = np.ones((H, W), bool)
    border[1:-1, 1:-1] = False
    assert np.abs(img[..., 4:7][border]).max() == 0.0
    return same


def test_scan_math_matches_reference_laserscan(scan_host):
    import subprocess
    import numpy as np
    from tests.helpers import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "scan_glue.npz"))
    scan, H, W = g["scan"], int(g["H"]), int(g["W"])
    data = (np.int32([len(scan), H, W]).tobytes() +
            np.float32([g["fov_up"], g["fov_down"], g["min_depth"], g["max_depth"]]).tobytes() + scan.tobytes())
    out = subprocess.run([scan_host], input=data, capture_output=True, check=True).stdout
    img = np.frombuffer(out[:H * W * 32], np.float32).reshape(H, W, 8)
    idx = np.frombuffer(out[H * W * 32:], np.int32).reshape(H, W)
    same = check_scan_image(img, idx, g, exact_pixels=1.0)
    assert same.all()


def test_scan_oracle_matches_reference_laserscan():
    import numpy as np
    from oracle import scan_oracle as S
    from tests.helpers import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "scan_glue.npz"))
    org, normed = S.scan_image(g["scan"], int(g["H"]), int(g["W"]), float(g["fov_up"]), float(g["fov_down"]),
                               float(g["min_depth"]), float(g["max_depth"]), np.arange(8) * 0.1, list(range(8)))
    assert np.array_equal(org.transpose(1, 2, 0), g["image"])
    assert np.allclose(normed, org - (np.arange(8, dtype=np.float32) * np.float32(0.1))[:, None, None])
    # IMU windows: hand-checked case (kitti.py:317-343): [t0, t1) membership, padding to T, truncation, empty window
    ts = np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.9])
    imu = np.arange(42, dtype=np.float32).reshape(7, 6)
    w, valid = S.imu_windows(ts, imu, np.array([0.1, 0.35, 0.6, 0.8, 1.0]), T=2)
    assert valid.tolist() == [True, True, False, True]
    assert np.array_equal(w[0], imu[1:3]) and np.array_equal(w[1], imu[4:6])        # truncated to T = 2, exact fit
    assert np.array_equal(w[2], np.zeros((2, 6))) and np.array_equal(w[3], np.stack([imu[6], np.zeros(6)]))
