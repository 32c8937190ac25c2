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


def allreduce_grads(flat_grad):
    """Sum the flat gradient arena over all ranks (mean is applied by the optimizer's grad_scale)."""
    if world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world_size()
