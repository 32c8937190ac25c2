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


class _ToyModel(torch.nn.Module):
    """Feature net -> (hook point) -> fusion / odometry net / heads, with the attribute names DeepLIO uses."""

    def __init__(self):
        super().__init__()
        self.lidar_feat_net = torch.nn.Linear(6, 5)
        self.fusion_net = torch.nn.Linear(5, 5)
        self.odom_feat_net = torch.nn.Linear(5, 4)
        self.fc_pos = torch.nn.Linear(4, 3)
        self.fc_ori = torch.nn.Linear(4, 3)
        self.on_head_grads_ready = None

    def forward(self, x):
        feat = torch.tanh(self.lidar_feat_net(x))
        if self.on_head_grads_ready is not None:
            feat.register_hook(lambda g: self.on_head_grads_ready())
        h = torch.tanh(self.odom_feat_net(torch.tanh(self.fusion_net(feat))))
        return self.fc_pos(h), self.fc_ori(h)


class _FlatOpt:
    """The part of FlatAdam the reducer reads (FlatAdam itself needs a CUDA device): params as views of one arena."""

    def __init__(self, params):
        self.params = list(params)
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat_grad = torch.zeros(total)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(3)
    model = _ToyModel()
    opt = _FlatOpt(model.parameters())
    red = parallel.OverlappedGradReducer(model, opt)
    n_late = sum(b - a for a, b in red.late_ranges)
    n_early = sum(b - a for a, b in red.early_ranges)
    torch.manual_seed(10 + rank)
    x = torch.randn(4, 6)
    pos, ori = model(x)
    (pos.pow(2).sum() + ori.sum()).backward()
    local = opt.flat_grad.clone()
    fired = red.fired
    scale = red.finish()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    out[rank] = (fired, n_late, n_early, opt.numel, float((opt.flat_grad - sum(gathered)).abs().max()), scale,
                 red.fired, len(red.work))
    dist.destroy_process_group()


def test_gloo_world2_overlapped_gradient_reducer():
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_overlap_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        fired, n_late, n_early, numel, err, scale, fired_after, pending = out[rank]
        assert fired, "the head-gradient hook did not fire during backward"
        assert n_late > 0 and n_early > 0 and n_late + n_early == numel
        assert err < 1e-6 and scale == 0.5
        assert not fired_after and pending == 0


def _extra_late_worker(rank, world, port, out):
    """The split-backward call pattern of bench.py on the CPU: no hook, ``fire()`` called explicitly after the
    downstream backward, ``extra_late`` holding a module (a criterion with sx / sq) and a bare Parameter."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(3)
    model = _ToyModel()
    crit = torch.nn.Module()
    crit.sx, crit.sq = torch.nn.Parameter(torch.tensor(0.0)), torch.nn.Parameter(torch.tensor(-3.0))
    extra = torch.nn.Parameter(torch.ones(3))
    opt = _FlatOpt(list(model.parameters()) + list(crit.parameters()) + [extra])
    red = parallel.OverlappedGradReducer(model, opt, extra_late=[crit, extra])
    model.on_head_grads_ready = None                      # fired explicitly, as with a split backward pass
    late = sum(b - a for a, b in red.late_ranges)
    base = sum(b - a for a, b in parallel.OverlappedGradReducer(model, opt).late_ranges)
    model.on_head_grads_ready = None
    torch.manual_seed(10 + rank)
    pos, ori = model(torch.randn(4, 6))
    loss = (pos.pow(2).sum() + ori.sum()) * torch.exp(-crit.sx) + crit.sx + crit.sq * extra.sum()
    loss.backward()
    local = opt.flat_grad.clone()
    assert not red.fired
    red.fire()
    pending = len(red.work)
    scale = red.finish()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    out[rank] = (late, base, pending, float((opt.flat_grad - sum(gathered)).abs().max()), scale)
    dist.destroy_process_group()


def test_gloo_world2_reducer_with_extra_late_parameters_fired_explicitly():
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_extra_late_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        late, base, pending, err, scale = out[rank]
        assert late == base + 4 + 4 + 4          # sx, sq (one padded slot each) and the 3-element parameter
        assert pending >= 1 and err < 1e-6 and scale == 0.5
