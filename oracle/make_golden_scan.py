"""Generate tests/golden/scan_glue.npz by EXECUTING the reference's own ``LaserScan`` (container only).

TEST INFRASTRUCTURE.  ``python -m oracle.make_golden_scan`` from the repo root, where /root/reference exists.
``deeplio/common/laserscan.py`` runs where it lies (its ``open3d`` import is an empty stub: only
``do_normal_projection1``, which DeepLIO does not call, uses it).  The depth filter of ``open_scan``
(laserscan.py:86-91) is applied here with the same numpy expressions, because ``open_scan`` itself reads a file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from oracle import scan_oracle as S  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "scan_glue.npz")
H, W, FOV_UP, FOV_DOWN, MIN_D, MAX_D = 16, 128, 3.0, -25.0, 1.0, 80.0


def main():
    ref_loader._install_stubs()
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    with ref_loader._in_tmp_cwd():
        from deeplio.common.laserscan import LaserScan
    scan4 = S.synthetic_scan(6000, seed=1)
    depth = np.linalg.norm(scan4[:, 0:3], 2, axis=1)
    idx = np.vstack((np.argwhere(depth < MIN_D), np.argwhere(depth > MAX_D)))
    kept = np.delete(scan4, idx, axis=0)                      # laserscan.py:86-91
    ls = LaserScan(project=False, H=H, W=W, fov_up=FOV_UP, fov_down=FOV_DOWN, min_depth=MIN_D, max_depth=MAX_D)
    ls.set_points(kept[:, 0:3], kept[:, 3])
    ls.do_range_projection()
    normal = ls.do_normal_projection()
    # kitti.py:83-97
    image = np.dstack((ls.proj_xyz / MAX_D, ls.proj_remission, normal, ls.proj_range))
    np.savez_compressed(OUT, scan=scan4, H=H, W=W, fov_up=FOV_UP, fov_down=FOV_DOWN, min_depth=MIN_D, max_depth=MAX_D,
                        proj_xyz=ls.proj_xyz, proj_range=ls.proj_range, proj_remission=ls.proj_remission,
                        proj_idx=ls.proj_idx, proj_normal=normal, image=image.astype(np.float32))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
