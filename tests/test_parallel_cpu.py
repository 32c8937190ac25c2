"""Host-side data-parallel logic on CPU: world_size-2 gloo processes (the GPU path uses the same functions over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeplio_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and parallel.world_size() == world
    # broadcast: every rank ends with rank 0's parameters and buffers
    torch.manual_seed(100 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7))
    model[1].running_mean.fill_(float(rank))
    parallel.broadcast_model(model)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7))
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7))
    same = all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), ref.state_dict().values()))
    # gradient exchange: the mean over ranks of per-shard gradients equals the full-batch gradient
    torch.manual_seed(7)
    x, y = torch.randn(8, 5), torch.randn(8, 7)
    lin = torch.nn.Linear(5, 7)
    parallel.broadcast_model(lin)
    lo, hi = parallel.shard_range(8, rank, world)
    loss = ((lin(x[lo:hi]) - y[lo:hi]) ** 2).sum() / 8 * world     # per-rank mean over its shard, scaled like DDP
    loss.backward()
    flat = torch.cat([p.grad.flatten() for p in lin.parameters()])
    scale = parallel.allreduce_grads(flat)
    lin_full = torch.nn.Linear(5, 7)
    lin_full.load_state_dict(lin.state_dict())
    (((lin_full(x) - y) ** 2).sum() / 8).backward()
    full = torch.cat([p.grad.flatten() for p in lin_full.parameters()])
    out[rank] = (same, float((flat * scale - full).abs().max()), scale)
    dist.destroy_process_group()


def test_gloo_world2_broadcast_and_gradient_allreduce():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        same, err, scale = out[rank]
        assert same, "broadcast_model did not replicate rank 0"
        assert err < 1e-6 and scale == 0.5


def test_shard_range_partitions_the_batch():
    for total in (1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = parallel.shard_range(total, r, world)
                seen.extend(range(lo, hi))
            assert seen == list(range(total))
