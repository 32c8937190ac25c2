"""LiDAR / IMU preprocessing on the device.

``project_scan`` replaces, per frame, ``LaserScan.open_scan`` + ``do_range_projection`` + ``do_normal_projection``
(deeplio/common/laserscan.py:68-98,122-191,215-248) and the image assembly of ``KittiRawData.get_velo_image`` /
``Kitti.transform_images`` (deeplio/datasets/kitti.py:83-97,345-364): raw velodyne points [N, 4] on the device ->
(untransformed, mean-subtracted) images [C, H, W], two kernel launches per frame (dlio_scan_project).
``imu_windows`` replaces ``Kitti.load_imus`` + ``transform_imus`` (kitti.py:317-343,366-368).  The reference runs
these in numpy inside DataLoader worker processes; on-disk formats (.bin scans, pickled OXTS) stay the caller's job.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def project_scan(points4, height, width, fov_up=3.0, fov_down=-25.0, min_depth=1.0, max_depth=80.0,
                 channels=(0, 1, 2, 4, 5, 6), mean_image=None, want_index=False):
    """points4: float32 CUDA tensor [N, 4] = (x, y, z, remission).  Returns (images_org, images_normalized[, index]):
    [len(channels), H, W] each; channel numbering of the reference's 8-channel image (xyz / max_depth, remission,
    normal, range); ``mean_image`` = the 8 per-channel means of config.yaml (``mean-image``)."""
    if not (points4.is_cuda and points4.dtype == torch.float32 and points4.dim() == 2 and points4.shape[1] == 4):
        raise RuntimeError("project_scan: points must be a float32 CUDA tensor [N, 4] (no CPU path)")
    points4 = points4.contiguous()
    dev = points4.device
    nsel = len(channels)
    org = torch.empty((nsel, height, width), device=dev, dtype=torch.float32)
    normed = torch.empty_like(org)
    idx = torch.empty((height, width), device=dev, dtype=torch.int32) if want_index else None
    scratch = torch.empty((L.scan_scratch_bytes(height, width) // 8,), device=dev, dtype=torch.int64)
    ch = (C.c_int * nsel)(*[int(c) for c in channels])
    mean = (C.c_float * 8)(*[float(m) for m in mean_image]) if mean_image is not None else None
    L.scan_project(ptr(points4), points4.shape[0], height, width, float(fov_up), float(fov_down), float(min_depth),
                   float(max_depth), ch, nsel, mean, ptr(scratch), ptr(org), ptr(normed), ptr(idx), _stream())
    return (org, normed, idx) if want_index else (org, normed)


def imu_windows(ts, imu, velo_ts, samples=15, mean=None, std=None):
    """ts [M] float64 (sorted seconds), imu [M, 6] float32, velo_ts [F] float64, all CUDA tensors ->
    (windows [F-1, samples, 6] float32, valid [F-1] bool)."""
    if not (ts.is_cuda and ts.dtype == torch.float64 and imu.dtype == torch.float32 and velo_ts.dtype == torch.float64):
        raise RuntimeError("imu_windows: ts / velo_ts float64 and imu float32 CUDA tensors expected (no CPU path)")
    ts, imu, velo_ts = ts.contiguous(), imu.contiguous(), velo_ts.contiguous()
    f = velo_ts.shape[0]
    out = torch.empty((f - 1, samples, 6), device=ts.device, dtype=torch.float32)
    valid = torch.empty((f - 1,), device=ts.device, dtype=torch.int32)
    m6 = (C.c_float * 6)(*[float(v) for v in mean]) if mean is not None else None
    s6 = (C.c_float * 6)(*[float(v) for v in std]) if std is not None else None
    L.imu_windows(ptr(ts), ptr(imu), ts.shape[0], ptr(velo_ts), f, samples, m6, s6, ptr(out), ptr(valid), _stream())
    return out, valid.bool()
