"""CPU restatement of the reference's per-frame preprocessing.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/ import this.

* ``range_projection``  -- ``LaserScan.open_scan`` depth filter + ``do_range_projection``
  (deeplio/common/laserscan.py:68-98,122-191): spherical projection of a point cloud, nearest point wins a pixel.
  The reference sorts by decreasing depth and scatters, so later (nearer) writes win; this file states it as a
  z-buffer (minimum over (depth, index)), which is the same image except when two points of one pixel have EXACTLY
  the same float32 depth (numpy's default argsort is not stable, so the reference itself is unspecified there; the
  z-buffer picks the lower index).
* ``normal_projection`` -- ``do_normal_projection`` (laserscan.py:215-248): weighted cross products of the four
  neighbour differences.
* ``scan_image``        -- ``KittiRawData.get_velo_image`` (deeplio/datasets/kitti.py:83-97) + ``Kitti.transform_images``
  (:345-364): 8 channels (xyz / max_depth, remission, normal, range), mean subtraction, channel selection.
* ``imu_windows``       -- ``Kitti.load_imus`` + ``transform_imus`` (kitti.py:317-343,366-368): OXTS samples between
  consecutive LiDAR timestamps, zero-padded / truncated to 15 rows, then normalised (padding rows included, as the
  reference does).  kitti.py cannot run under numpy 2 (``np.float`` / ``np.int``, :331,337-338), so this part is
  pinned by restatement only.

Pinned by ``oracle/make_golden_scan.py`` (executes the reference's own ``LaserScan`` from /root/reference on a
synthetic scan) -> tests/golden/scan_glue.npz; tests/test_oracle_golden.py holds this file to it.
"""
import numpy as np


def depth_filter(points, min_depth, max_depth):
    """laserscan.py:86-91: drop points with depth > max_depth or depth < min_depth.  Returns the kept mask."""
    depth = np.linalg.norm(points[:, 0:3], 2, axis=1)
    return ~((depth > max_depth) | (depth < min_depth))


def pixel_of(points, H, W, fov_up, fov_down):
    """laserscan.py:129-162, float32 arithmetic as numpy performs it on float32 arrays."""
    fov_up_r = fov_up / 180.0 * np.pi
    fov_down_r = fov_down / 180.0 * np.pi
    fov = abs(fov_down_r) + abs(fov_up_r)
    depth = np.linalg.norm(points, 2, axis=1)
    yaw = -np.arctan2(points[:, 1], points[:, 0])
    pitch = np.arcsin(points[:, 2] / depth)
    proj_x = 0.5 * (yaw / np.pi + 1.0)
    proj_y = 1.0 - (pitch + abs(fov_down_r)) / fov
    proj_x = proj_x * W
    proj_y = proj_y * H
    px = np.maximum(0, np.minimum(W - 1, np.floor(proj_x))).astype(np.int32)
    py = np.maximum(0, np.minimum(H - 1, np.floor(proj_y))).astype(np.int32)
    return px, py, depth


def range_projection(points, remissions, H, W, fov_up=3.0, fov_down=-25.0):
    """Points already depth-filtered ([M,3] float32).  Returns proj_xyz [H,W,3], proj_range [H,W],
    proj_remission [H,W], proj_idx [H,W] (index into ``points``, 0 where empty -- the reference's initial value)."""
    px, py, depth = pixel_of(points, H, W, fov_up, fov_down)
    pix = py.astype(np.int64) * W + px
    # nearest wins; ties -> lower index: lexicographic minimum of (depth, index) per pixel
    order = np.lexsort((np.arange(len(depth)), depth, pix))
    first = np.ones(len(order), dtype=bool)
    first[1:] = pix[order][1:] != pix[order][:-1]
    win = order[first]
    proj_xyz = np.zeros((H, W, 3), np.float32)
    proj_range = np.zeros((H, W), np.float32)
    proj_rem = np.zeros((H, W), np.float32)
    proj_idx = np.zeros((H, W), np.int32)
    proj_xyz.reshape(-1, 3)[pix[win]] = points[win]
    proj_range.reshape(-1)[pix[win]] = depth[win]
    proj_rem.reshape(-1)[pix[win]] = remissions[win]
    proj_idx.reshape(-1)[pix[win]] = win
    return proj_xyz, proj_range, proj_rem, proj_idx


def normal_projection(proj_xyz, proj_range):
    img = np.dstack((proj_xyz, proj_range)).astype(np.float32)
    c = img[1:-1, 1:-1]
    top, bottom = img[:-2, 1:-1] - c, img[2:, 1:-1] - c
    left, right = img[1:-1, :-2] - c, img[1:-1, 2:] - c
    wt = [np.exp(np.float32(-0.8) * np.abs(d[..., 3:4])) for d in (top, left, bottom, right)]
    t, l, b, r = (w * d[..., :3] for w, d in zip(wt, (top, left, bottom, right)))
    n = np.cross(t, l) + np.cross(l, b) + np.cross(b, r) + np.cross(r, t)
    n = n / (np.linalg.norm(n, axis=2, keepdims=True) + np.float32(1e-8))
    return np.pad(n, ((1, 1), (1, 1), (0, 0))).astype(np.float32)


def scan_image(points4, H, W, fov_up, fov_down, min_depth, max_depth, mean_image, channels):
    """[N,4] raw velodyne points (x, y, z, remission) -> (untransformed, mean-subtracted) images [len(channels), H, W]."""
    keep = depth_filter(points4, min_depth, max_depth)
    pts, rem = points4[keep, 0:3], points4[keep, 3]
    xyz, rng, prem, _ = range_projection(pts, rem, H, W, fov_up, fov_down)
    nrm = normal_projection(xyz, rng)
    image = np.dstack((xyz / max_depth, prem, nrm, rng)).astype(np.float32)
    org = image.transpose(2, 0, 1)
    normed = org - np.asarray(mean_image, np.float32)[:, None, None]
    return org[channels], normed[channels]


def imu_windows(ts, imu, velo_ts, T=15, mean=None, std=None):
    """ts [M] OXTS timestamps (sorted), imu [M,6] (ax, ay, az, wx, wy, wz), velo_ts [F] LiDAR timestamps ->
    (windows [F-1, T, 6] float32, valid [F-1] bool)."""
    out, valid = [], []
    for i in range(len(velo_ts) - 1):
        sel = np.argwhere((ts >= velo_ts[i]) & (ts < velo_ts[i + 1])).flatten()
        if len(sel) == 0:
            vals = np.zeros((T, 6), np.float64)
            valid.append(False)
        else:
            vals = imu[sel].astype(np.float64)
            vals = np.pad(vals, ((0, max(T - len(sel), 0)), (0, 0)))[:T]
            valid.append(True)
        out.append(vals)
    out = np.stack(out)
    if mean is not None:
        out = (out - np.asarray(mean, np.float64)) / np.asarray(std, np.float64)
    return out.astype(np.float32), np.asarray(valid)


def synthetic_scan(n, seed=0, fov_up=3.0, fov_down=-25.0):
    """A velodyne-like cloud: rings in pitch, full yaw sweep, ranges 0.3 .. 95 m (some outside the 1 .. 80 m filter),
    a few duplicated points (exact depth ties) and points on the +/- pi yaw seam.  [n,4] float32."""
    rng = np.random.default_rng(seed)
    yaw = rng.uniform(-np.pi, np.pi, n)
    pitch = np.deg2rad(rng.uniform(fov_down - 1.0, fov_up + 1.0, n))
    r = np.exp(rng.uniform(np.log(0.3), np.log(95.0), n))
    yaw[:8] = np.array([np.pi, -np.pi, np.pi - 1e-7, -np.pi + 1e-7, 0.0, 1e-9, np.pi / 2, -np.pi / 2])
    pts = np.stack([r * np.cos(pitch) * np.cos(yaw), r * np.cos(pitch) * np.sin(yaw), r * np.sin(pitch),
                    rng.uniform(0, 1, n)], 1).astype(np.float32)
    pts[100:110] = pts[90:100]          # exact duplicates
    return pts
