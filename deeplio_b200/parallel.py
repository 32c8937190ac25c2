"""Data-parallel plumbing: one process per GPU, NCCL over NVLink for the single exchange step of the path
(the gradient all-reduce).  The reference has no distributed code at all.

Frame-pair samples are independent, so the batch is sharded by sample; BatchNorm statistics stay per rank
(the reference has no SyncBN).  Parameters and BN buffers are broadcast from rank 0 once; every step the flat
gradient arena (deeplio_b200.optim.FlatAdam) is summed with ONE all-reduce and the 1/world factor is folded
into the Adam kernel (``grad_scale``).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of ``total`` samples for ``rank`` (first ``total % world`` ranks get one more)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_model(model, src=0):
    """Rank ``src``'s parameters and buffers (BN running statistics) to every rank."""
    if world_size() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


class OverlappedGradReducer:
    """Gradient exchange overlapped with the encoders' backward pass.

    The gradients of everything downstream of the feature nets -- fusion net, odometry net (35.7 M of the 45.8 M
    parameters of the benchmark model) and the two heads -- are final when the gradient of the LiDAR feature vector
    arrives, i.e. BEFORE the convolutional encoders (more than half of the step) run their backward.  ``DeepLIO``
    calls ``model.on_head_grads_ready`` at that moment (a tensor hook); their slices of the flat gradient arena are
    all-reduced asynchronously on NCCL's stream while the encoder backward runs, and ``finish()`` reduces the rest
    (feature-net gradients) after ``backward()`` and joins.  Same sums as one all-reduce of the whole arena."""

    def __init__(self, model, opt, extra_late=()):
        """``extra_late``: further modules / parameters whose gradients are complete at the same moment (the loss
        module's sx / sq; with a split backward pass -- ``lidar_feat_net.split_backward`` -- also the IMU net and the
        LiDAR net's fc1, i.e. everything but the two encoders)."""
        self.flat_grad = opt.flat_grad
        late = set()
        mods = [getattr(model, name, None) for name in ("fusion_net", "odom_feat_net", "fc_pos", "fc_ori")] + list(extra_late)
        for m in mods:
            if isinstance(m, torch.nn.Module):
                late.update(id(p) for p in m.parameters())
            elif isinstance(m, torch.nn.Parameter):
                late.add(id(m))
        self.late_ranges, self.early_ranges = [], []
        ends = opt.offsets[1:] + [opt.numel]
        for p, a, b in zip(opt.params, opt.offsets, ends):
            rs = self.late_ranges if id(p) in late else self.early_ranges
            if rs and rs[-1][1] == a:
                rs[-1][1] = b
            else:
                rs.append([a, b])
        self.work, self.fired = [], False
        self.disabled = False         # measurement only (bench.py DLIO_NO_EXCHANGE=1): issue no collective at all
        model.on_head_grads_ready = self._fire

    def fire(self):
        """The downstream gradients are complete: start reducing them (asynchronously, on NCCL's stream)."""
        self._fire()

    def _fire(self):
        if world_size() > 1 and not self.fired and not self.disabled:
            for a, b in self.late_ranges:
                self.work.append(dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, async_op=True))
        self.fired = True

    def finish(self):
        """Call after backward(): reduces what the hook did not, waits for everything.  Returns the 1/world factor
        for the optimizer's ``grad_scale``."""
        if world_size() > 1 and not self.disabled:
            if self.fired:
                for a, b in self.early_ranges:
                    dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM)
            else:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            for w in self.work:
                w.wait()
        self.work, self.fired = [], False
        return 1.0 / world_size()


def allreduce_grads(flat_grad):
    """Sum the flat gradient arena over all ranks (mean is applied by the optimizer's grad_scale)."""
    if world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world_size()
