"""LiDAR / IMU preprocessing kernels against the reference-generated fixture
(tests/golden/scan_glue.npz: the reference's own LaserScan executed by oracle/make_golden_scan.py) and the oracle
(oracle/scan_oracle.py) at the BASELINE image size."""
import os

import numpy as np
import pytest
import torch

from oracle import scan_oracle as S
from tests.helpers import GOLDEN_DIR
from tests.test_host_logic_cpu import check_scan_image

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEAN = [-0.0014, 0.0043, -0.011, 0.2258, -0.0024, 0.0037, 0.3793, 0.1115]      # config.yaml:26


def test_scan_projection_matches_reference_laserscan():
    from deeplio_b200 import scan
    g = np.load(os.path.join(GOLDEN_DIR, "scan_glue.npz"))
    pts = torch.from_numpy(g["scan"]).to(DEV)
    org, normed, idx = scan.project_scan(pts, int(g["H"]), int(g["W"]), float(g["fov_up"]), float(g["fov_down"]),
                                         float(g["min_depth"]), float(g["max_depth"]), channels=range(8),
                                         mean_image=MEAN, want_index=True)
    img = org.permute(1, 2, 0).cpu().numpy()
    same = check_scan_image(img, idx.cpu().numpy(), g, exact_pixels=0.995)
    assert np.allclose(normed.cpu().numpy(), org.cpu().numpy() - np.float32(MEAN)[:, None, None], atol=1e-7)
    assert same.mean() >= 0.995


def test_scan_projection_full_size_against_oracle():
    """64 x 2048 image from a 120 k-point cloud (a KITTI frame's size), the six channels the nets consume."""
    from deeplio_b200 import scan
    H, W = 64, 2048
    cloud = S.synthetic_scan(120000, seed=3)
    channels = [0, 1, 2, 4, 5, 6]
    org, normed = scan.project_scan(torch.from_numpy(cloud).to(DEV), H, W, channels=channels, mean_image=MEAN)
    ref_org, ref_norm = S.scan_image(cloud, H, W, 3.0, -25.0, 1.0, 80.0, MEAN, channels)
    got = org.cpu().numpy()
    # pixels whose xyz agrees exactly hold the same winning point (different libm: a few 1e-4 of the points round
    # into a neighbouring pixel)
    same = (got[0:3] == ref_org[0:3]).all(axis=0)
    assert same.mean() >= 0.995, same.mean()
    nb = same.copy()
    nb[1:] &= same[:-1]; nb[:-1] &= same[1:]; nb[:, 1:] &= same[:, :-1]; nb[:, :-1] &= same[:, 1:]
    well = nb & (np.linalg.norm(ref_org[3:6], axis=0) > 0.9)
    assert well.mean() > 0.3
    # normals: n / (|n| + 1e-8) of a sum of four weighted cross products -- where range differences are tens of metres
    # the weights exp(-0.8 |dr|) are ~1e-7 and the cross products cancel, so 1-ulp differences of expf are amplified:
    # 99.9 % of the well-defined normals within 1e-4, all within 1e-2
    err = np.abs(got[3:6] - ref_org[3:6]).max(axis=0)[well]
    assert (err < 1e-4).mean() > 0.999 and err.max() < 1e-2, ((err < 1e-4).mean(), err.max())
    assert np.abs(normed.cpu().numpy() - ref_norm)[:, well].max() < 1e-2
    assert np.abs(normed.cpu().numpy()[0:3] - ref_norm[0:3])[:, same].max() == 0.0
    # an empty cloud gives the empty image (zeros minus the mean)
    org0, norm0 = scan.project_scan(torch.zeros(0, 4, device=DEV), 16, 64, channels=channels, mean_image=MEAN)
    assert float(org0.abs().max()) == 0.0
    assert np.allclose(norm0.cpu().numpy(), -np.float32(MEAN)[channels][:, None, None] * np.ones((1, 16, 64), np.float32))


def test_imu_windows_match_oracle():
    from deeplio_b200 import scan
    rng = np.random.default_rng(0)
    ts = np.cumsum(rng.uniform(0.005, 0.015, 400))                 # ~100 Hz OXTS stream with jitter
    imu = rng.standard_normal((400, 6)).astype(np.float32)
    velo = np.concatenate([[ts[3] - 1e-4], ts[3] + np.cumsum(rng.uniform(0.05, 0.25, 12)), [ts[-1] + 1.0, ts[-1] + 2.0]])
    mean, std = rng.standard_normal(6), rng.uniform(0.5, 2.0, 6)
    for T in (15, 4):
        ref, valid = S.imu_windows(ts, imu, velo, T=T, mean=mean, std=std)
        out, v = scan.imu_windows(torch.from_numpy(ts).to(DEV), torch.from_numpy(imu).to(DEV),
                                  torch.from_numpy(velo).to(DEV), samples=T, mean=mean, std=std)
        assert v.cpu().numpy().tolist() == valid.tolist() and not valid[-1]
        assert np.allclose(out.cpu().numpy(), ref, rtol=1e-6, atol=1e-6)
    ref, _ = S.imu_windows(ts, imu, velo, T=15)
    out, _ = scan.imu_windows(torch.from_numpy(ts).to(DEV), torch.from_numpy(imu).to(DEV), torch.from_numpy(velo).to(DEV))
    assert np.array_equal(out.cpu().numpy(), ref)
