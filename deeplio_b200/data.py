"""On-device batch preparation: drop-in for ``DataCombiCreater`` (deeplio/models/misc.py:9-128).

The reference, per step: ``imgs[:, combinations]`` (a gather that materialises every frame pair, [B,S,2,C,H,W]),
a channel split with a ``.contiguous()`` copy of the normals, the encoders' ``reshape(b*s, t*c, h, w)`` copy of the
strided xyz view (lidar_feat_nets.py:216-218), and -- for the ground truth -- Python loops over batch and pairs with
numpy ``inv_SE3`` and ``SO3.log`` on the CPU.  Here:

* the frames go to the device once ([B, F, C, H, W], no pairing on the host: (S+1)/(2S) of the bytes);
* ``res_imgs`` / ``res_normals`` are ``PairedFrames`` handles; the LiDAR nets hand them to ``dlio_pair_gather``, which
  reads each frame of a pair where it lies and writes the padded-NHWC operand of the first convolution directly --
  no [B,S,2,C,H,W] tensor ever exists.  ``PairedFrames.materialize()`` gives the reference's tensor (a strided
  gather, only for callers that really want it, e.g. the trainer's isnan guards -- use ``pose.check_finite`` on
  ``.frames`` instead);
* ``res_gt_f2f`` / ``res_gt_f2g`` come from one launch of ``dlio_gt_relative`` on the device.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import pose
from ._lib import ptr
from .config import get_config_container


class PairedFrames:
    """Frame pairs as a view description: ``frames`` [B, F, Ctot, H, W] on the device, ``combinations`` [S, 2]
    frame indices, channels ``c0 .. c0 + C`` -- what ``imgs[:, combinations][:, :, :, c0:c0+C]`` would hold."""

    def __init__(self, frames, combinations, c0, c):
        if frames.dim() != 5:
            raise RuntimeError("PairedFrames: frames must be [B, F, C, H, W]")
        if frames.stride(4) != 1 or frames.stride(3) != frames.shape[4]:
            frames = frames.contiguous()
        self.frames = frames
        self.combinations = np.ascontiguousarray(np.asarray(combinations, dtype=np.int32).reshape(-1, 2))
        self.c0, self.c = int(c0), int(c)

    @property
    def shape(self):
        b, f, _, h, w = self.frames.shape
        return torch.Size((b, len(self.combinations), 2, self.c, h, w))

    @property
    def device(self):
        return self.frames.device

    @property
    def dtype(self):
        return self.frames.dtype

    @property
    def is_cuda(self):
        return self.frames.is_cuda

    def __len__(self):
        return self.frames.shape[0]

    def materialize(self):
        """The reference's [B, S, 2, C, H, W] tensor (misc.py:65-69)."""
        idx = torch.as_tensor(self.combinations.astype(np.int64), device=self.frames.device)
        return self.frames[:, idx][:, :, :, self.c0:self.c0 + self.c]


def frames_bound(paired):
    """Upper bound of max|frames| (one float in device memory), taken over the whole frames tensor -- all channels, so
    the views cut from one tensor (xyz, normals) can share it."""
    fr = paired.frames
    bound = torch.empty(1, device=fr.device, dtype=torch.float32)
    L.absmax(ptr(fr), fr.numel(), ptr(bound), torch.cuda.current_stream().cuda_stream)
    return bound


def pair_gather(run_device, paired, c_pad, ph, pw, split, flags=None, flag_bit=0, f16=False, bound=None):
    """``PairedFrames`` -> padded NHWC [B*S, H+2ph, W+2pw, c_pad] in one launch: fp32 (+ TF32 lo plane when
    ``split``) -> (hi, lo), or with ``f16`` the packed fp16 planes [.., 2, c_pad] and their bound -> (h2, bound)."""
    fr = paired.frames
    b, f, _, h, w = fr.shape
    s = len(paired.combinations)
    shape = (b * s, h + 2 * ph, w + 2 * pw, c_pad)
    comb = paired.combinations.ctypes.data_as(C.c_void_p)
    st = torch.cuda.current_stream().cuda_stream
    if f16:
        if not fr.is_contiguous():
            raise RuntimeError("pair_gather(f16=True): frames must be contiguous (their bound is taken over the storage)")
        bound = frames_bound(paired) if bound is None else bound
        h2 = torch.empty(shape[:3] + (2, c_pad), device=run_device, dtype=torch.float16)
        L.pair_gather(ptr(fr), fr.stride(0), fr.stride(1), fr.stride(2), b, f, comb, s, paired.c0, paired.c,
                      L.Tensor4(b * s, h, w, c_pad, ph, pw), None, None, ptr(h2), ptr(bound), ptr(flags), flag_bit, st)
        return h2, bound
    hi = torch.empty(shape, device=run_device, dtype=torch.float32)
    lo = torch.empty_like(hi) if split else None
    L.pair_gather(ptr(fr), fr.stride(0), fr.stride(1), fr.stride(2), b, f, comb, s, paired.c0, paired.c,
                  L.Tensor4(b * s, h, w, c_pad, ph, pw), ptr(hi), ptr(lo), None, None, ptr(flags), flag_bit, st)
    return hi, lo


def ground_truth(gts, combinations):
    """gts [B, F, 15] on the device -> (gt_f2f [B,S,6], gt_f2g [B,S,7]) (misc.py:83-125), one launch; flags go to
    ``pose.status_word`` (``pose.raise_for_status`` raises the reference's errors)."""
    if not (gts.is_cuda and gts.dtype == torch.float32):
        raise RuntimeError("ground_truth: gts must be a float32 CUDA tensor (no CPU path)")
    gts = gts.contiguous()
    comb = np.ascontiguousarray(np.asarray(combinations, dtype=np.int32).reshape(-1, 2))
    b, f, _ = gts.shape
    s = len(comb)
    f2f = torch.empty((b, s, 6), device=gts.device, dtype=torch.float32)
    f2g = torch.empty((b, s, 7), device=gts.device, dtype=torch.float32)
    L.gt_relative(ptr(gts), b, f, comb.ctypes.data_as(C.c_void_p), s, ptr(f2f), ptr(f2g),
                  ptr(pose.status_word(gts.device)), torch.cuda.current_stream().cuda_stream)
    return f2f, f2g


class DataCombiCreater:
    """Same interface as the reference class: ``creater(data)`` fills ``res_imgs``, ``res_normals``, ``res_imu``,
    ``res_gt_f2f``, ``res_gt_f2g``, ``res_gt_global`` (and the ``*_org`` twins) from a collated batch dict
    (datasets/misc.py:12-32).  ``check``: raise on invalid ground-truth rotations right away (one synchronisation),
    as the reference does (misc.py:104-107)."""

    def __init__(self, combinations, device="cuda", check=False):
        self.combinations = np.asarray(combinations)
        self.device = device
        self.seq_size = get_config_container().seq_size
        self.check = check
        self.res_imgs = self.res_img_org = self.res_normals = self.res_normals_org = None
        self.res_imu = self.res_gt_f2f = self.res_gt_f2g = self.res_gt_global = None

    def process(self, data):
        imgs, normals, imgs_org, normals_org, imus = [], [], [], [], []
        if "images" in data:
            imgs, normals = self.process_images(data["images"].to(self.device, non_blocking=True))
            if "untrans-images" in data:
                imgs_org, normals_org = self.process_images(data["untrans-images"].to(self.device, non_blocking=True))
        if "imus" in data:
            imus = data["imus"].to(self.device, non_blocking=True)
        gt_global = data["gts"].to(self.device, non_blocking=True)
        gt_f2f, gt_f2g = ground_truth(gt_global.float(), self.combinations)
        if self.check:
            pose.raise_for_status(gt_global.device)
        self.res_imgs, self.res_img_org = imgs, imgs_org
        self.res_normals, self.res_normals_org = normals, normals_org
        self.res_imu = imus
        self.res_gt_f2f, self.res_gt_f2g, self.res_gt_global = gt_f2f, gt_f2g, gt_global

    def process_images(self, imgs):
        c = imgs.shape[2] // 2
        return PairedFrames(imgs, self.combinations, 0, c), PairedFrames(imgs, self.combinations, c, imgs.shape[2] - c)

    def __call__(self, args):
        return self.process(args)
