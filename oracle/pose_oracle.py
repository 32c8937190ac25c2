"""CPU restatement of the caller-side glue around the model.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.

What is restated, and how it is pinned:

* ``Trainer.se3_to_SE3`` (deeplio/models/trainer.py:324-351), ``DataCombiCreater.process_ground_turth``
  (deeplio/models/misc.py:83-125) and ``HWSLoss`` / ``LWSLoss`` (deeplio/losses/losses.py:11-96) -- the reference's own
  code.  Pinned: ``oracle/make_golden_pose.py`` executes those reference functions themselves (imported from
  /root/reference) and stores their outputs in tests/golden/pose_glue.pt; tests/test_oracle_golden.py holds this
  file to them.
* ``liegroups.torch.SO3`` (utiasSTARS/liegroups, no version pinned by the reference -- there is no requirements
  file -- and ABSENT from this image): ``exp``, ``log``, ``from_matrix`` / ``is_valid_matrix`` / ``normalize``,
  ``to_quaternion`` restated below from the library's published algorithm (Rodrigues formula with a first-order
  branch below 1e-6 rad, log through acos of the trace, SVD projection onto SO(3), the four-branch matrix ->
  quaternion conversion in wxyz order).  **PARITY UNPINNED for this class**: no reference test or golden vector
  covers it and the library itself cannot be run here; when the reference functions above are executed for the
  goldens, THIS class is what they call.  Its call sites: trainer.py:339,349; misc.py:104,119.
"""

import torch

TOL = 1e-6      # liegroups.torch.utils.isclose default


def _isclose(a, b, tol=TOL):
    return (a - b).abs() < tol


def wedge(phi):
    """[B,3] -> [B,3,3] skew-symmetric matrices."""
    z = torch.zeros_like(phi[:, 0])
    return torch.stack([torch.stack([z, -phi[:, 2], phi[:, 1]], 1),
                        torch.stack([phi[:, 2], z, -phi[:, 0]], 1),
                        torch.stack([-phi[:, 1], phi[:, 0], z], 1)], 1)


def vee(m):
    return torch.stack([m[:, 2, 1], m[:, 0, 2], m[:, 1, 0]], 1)


class SO3:
    """Batched rotation matrices, the subset of ``liegroups.torch.SO3`` the reference calls."""
    dim = 3

    def __init__(self, mat):
        self.mat = mat

    def as_matrix(self):
        return self.mat

    @classmethod
    def exp(cls, phi):
        squeeze = phi.dim() < 2
        if squeeze:
            phi = phi.unsqueeze(0)
        angle = phi.norm(p=2, dim=1)
        small = _isclose(angle, 0.0)
        eye = torch.eye(3, dtype=phi.dtype, device=phi.device).expand(phi.shape[0], 3, 3)
        # first-order branch near zero, Rodrigues elsewhere (the division is guarded, the branch is selected after)
        safe = torch.where(small, torch.ones_like(angle), angle)
        axis = phi / safe.unsqueeze(1)
        s, c = safe.sin()[:, None, None], safe.cos()[:, None, None]
        large = c * eye + (1.0 - c) * axis.unsqueeze(2) * axis.unsqueeze(1) + s * wedge(axis)
        mat = torch.where(small[:, None, None], eye + wedge(phi), large)
        return cls(mat.squeeze(0) if squeeze else mat)

    @classmethod
    def is_valid_matrix(cls, mat):
        if mat.dim() < 3:
            mat = mat.unsqueeze(0)
        det_ok = _isclose(torch.linalg.det(mat), 1.0)
        eye = torch.eye(3, dtype=mat.dtype, device=mat.device)
        inv_ok = _isclose(mat.transpose(2, 1).bmm(mat), eye).flatten(1).all(dim=1)
        return det_ok & inv_ok

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        valid = cls.is_valid_matrix(mat)
        if not (valid.all() or normalize):
            raise ValueError("Invalid rotation matrix. Use normalize=True to handle rounding errors.")
        res = cls(mat)
        if normalize and not valid.all():
            res.normalize(~valid)
        return res

    def normalize(self, mask=None):
        """SVD projection onto SO(3): U diag(1, 1, det U det V) V^T, applied to the matrices selected by ``mask``."""
        squeeze = self.mat.dim() < 3
        m = self.mat.unsqueeze(0) if squeeze else self.mat
        u, _, vh = torch.linalg.svd(m)
        d = torch.linalg.det(u) * torch.linalg.det(vh)
        s = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], 1))
        proj = u.bmm(s).bmm(vh)
        if mask is None:
            mask = torch.ones(m.shape[0], dtype=torch.bool, device=m.device)
        out = torch.where(mask[:, None, None], proj, m)
        self.mat = out.squeeze(0) if squeeze else out

    def log(self):
        squeeze = self.mat.dim() < 3
        m = self.mat.unsqueeze(0) if squeeze else self.mat
        cos_angle = (0.5 * (m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]) - 0.5).clamp(-1.0, 1.0)
        angle = cos_angle.acos()
        small = _isclose(angle, 0.0)
        eye = torch.eye(3, dtype=m.dtype, device=m.device)
        safe = torch.where(small, torch.ones_like(angle), angle)
        large = vee((0.5 * safe / safe.sin())[:, None, None] * (m - m.transpose(2, 1)))
        phi = torch.where(small[:, None], vee(m - eye), large)
        return phi.squeeze(0) if squeeze else phi

    def to_quaternion(self, ordering="wxyz"):
        squeeze = self.mat.dim() < 3
        R = self.mat.unsqueeze(0) if squeeze else self.mat
        tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
        qw = 0.5 * torch.sqrt((1.0 + tr).clamp_min(0.0))
        near = _isclose(qw, 0.0)

        def branch(i):
            j, k = (i + 1) % 3, (i + 2) % 3
            d = 2.0 * torch.sqrt((1.0 + R[:, i, i] - R[:, j, j] - R[:, k, k]).clamp_min(1e-30))
            q = [None] * 4
            q[0] = (R[:, k, j] - R[:, j, k]) / d
            q[1 + i] = 0.25 * d
            q[1 + j] = (R[:, j, i] + R[:, i, j]) / d
            q[1 + k] = (R[:, i, k] + R[:, k, i]) / d
            return torch.stack(q, 1)
        safe_w = torch.where(near, torch.ones_like(qw), qw)
        d = 4.0 * safe_w
        far = torch.stack([qw, (R[:, 2, 1] - R[:, 1, 2]) / d, (R[:, 0, 2] - R[:, 2, 0]) / d,
                           (R[:, 1, 0] - R[:, 0, 1]) / d], 1)
        c1 = (R[:, 0, 0] > R[:, 1, 1]) & (R[:, 0, 0] > R[:, 2, 2])
        c2 = (~c1) & (R[:, 1, 1] > R[:, 2, 2])
        nz = torch.where(c1[:, None], branch(0), torch.where(c2[:, None], branch(1), branch(2)))
        q = torch.where(near[:, None], nz, far)
        if ordering == "xyzw":
            q = q[:, [1, 2, 3, 0]]
        return q.squeeze(0) if squeeze else q


# --------------------------------------------------------------------------- trainer.py:324-351
def se3_to_SE3(f2f_x, f2f_r):
    """Chains frame-to-frame increments into frame-to-start poses.  Returns (f2g_x [B,S,3], f2g_q [B,S,4] wxyz,
    status): status bit 1 = det(exp(w)) not close to 1, bit 2 = det(R_accumulated) not close to 1 (the reference
    raises ValueError on either, trainer.py:341-348; torch.isclose defaults rtol 1e-5, atol 1e-8)."""
    B, S, _ = f2f_x.shape
    dt = f2f_x.dtype
    R_prev = torch.eye(3, dtype=dt).expand(B, 3, 3)
    t_prev = torch.zeros(B, 3, dtype=dt)
    xs, qs, status = [], [], 0
    for s in range(S):
        R_cur = SO3.exp(f2f_r[:, s]).as_matrix()
        if not torch.isclose(torch.linalg.det(R_cur), torch.ones(B, dtype=dt)).all():
            status |= 1
        t_prev = torch.bmm(R_prev, f2f_x[:, s].unsqueeze(2)).squeeze(2) + t_prev
        R_prev = torch.bmm(R_prev, R_cur)
        if not torch.isclose(torch.linalg.det(R_prev), torch.ones(B, dtype=dt)).all():
            status |= 2
        qs.append(SO3.from_matrix(R_prev, normalize=True).to_quaternion())
        xs.append(t_prev)
    return torch.stack(xs, 1), torch.stack(qs, 1), status


# --------------------------------------------------------------------------- misc.py:83-125
def ground_truth(gts, combinations):
    """gts [B, F, 15] = (t 3, R 9 row-major, v 3) per frame (kitti.py:292-301) -> (f2f [B,S,6] = (dx, so(3) log),
    f2g [B,S,7] = (dx, quaternion wxyz) relative to frame 0)."""
    B = gts.shape[0]
    t = gts[:, :, 0:3]
    R = gts[:, :, 3:12].reshape(B, -1, 3, 3)
    f2f, f2g = [], []
    for i, j in combinations:
        Ri_t = R[:, i].transpose(2, 1)
        Rij = Ri_t.bmm(R[:, j])
        dx = Ri_t.bmm((t[:, j] - t[:, i]).unsqueeze(2)).squeeze(2)
        f2f.append(torch.cat([dx, SO3.from_matrix(Rij, normalize=False).log()], 1))
    R0_t = R[:, 0].transpose(2, 1)
    for i, j in combinations:
        R0j = R0_t.bmm(R[:, j])
        dx = R0_t.bmm((t[:, j] - t[:, 0]).unsqueeze(2)).squeeze(2)
        f2g.append(torch.cat([dx, SO3.from_matrix(R0j).to_quaternion()], 1))
    return torch.stack(f2f, 1), torch.stack(f2g, 1)


# --------------------------------------------------------------------------- losses.py:11-96
def pose_loss(pred_t, pred_w, pred_p, pred_q, gt_t, gt_w, gt_p, gt_q, sx=None, sq=None, beta=None,
              loss_types=(True, True)):
    """HWSLoss (``sx``, ``sq`` given: learned homoscedastic weights, losses.py:68-86) or LWSLoss (``beta`` given:
    losses.py:21-38).  The global tensors are the slices the trainer passes (trainer.py:260-263)."""
    mse = torch.nn.functional.mse_loss
    Lt = mse(pred_t, gt_t) if loss_types[0] else 0.0
    Lw = mse(pred_w, gt_w) if loss_types[0] else 0.0
    Lp = mse(pred_p, gt_p) if loss_types[1] else 0.0
    Lq = mse(pred_q, gt_q) if loss_types[1] else 0.0
    if beta is not None:
        return (Lp + Lt) + beta * (Lq + Lw)
    return (Lp + Lt) * torch.exp(-sx) + sx + (Lq + Lw) * torch.exp(-sq) + sq


def synthetic_gts(B, F, seed=0, dtype=torch.float32):
    """Smooth yaw / pitch / forward-motion trajectories: [B, F, 15] = (t, R row-major, v)."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(B, F, 15, dtype=torch.float64)
    for b in range(B):
        R = SO3.exp((torch.randn(3, generator=g, dtype=torch.float64) * 0.5)).as_matrix()
        t = torch.randn(3, generator=g, dtype=torch.float64) * 10.0
        for f in range(F):
            w = torch.tensor([0.002, 0.004, 0.03], dtype=torch.float64) * (1.0 + torch.randn(3, generator=g, dtype=torch.float64))
            v = torch.tensor([1.2, 0.02, 0.01], dtype=torch.float64) * (1.0 + 0.3 * torch.randn(3, generator=g, dtype=torch.float64))
            out[b, f, 0:3], out[b, f, 3:12], out[b, f, 12:15] = t, R.reshape(9), v
            t = t + R @ v
            R = R @ SO3.exp(w).as_matrix()
    return out.to(dtype)
