"""Whole-step CUDA graph (deeplio_b200/graph.py): a replayed step leaves the same gradients as an eager step, takes
new inputs through its static buffers, and draws fresh dropout masks on every replay."""
import argparse

import pytest
import torch

from oracle import deeplio_oracle as O
from oracle.configs import make_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(no_dropout):
    from deeplio_b200 import nets
    from deeplio_b200.config import build_config_container
    from deeplio_b200.optim import FlatAdam
    B, S, H, W, T = 2, 2, 16, 128, 6
    cfg = make_cfg(lidar="lidar-feat-simple-1", imu="imu-feat-rnn", odom="odom-feat-rnn", seq=S, height=H, width=W,
                   odom_hidden=64, no_dropout=no_dropout)
    build_config_container(cfg, argparse.Namespace(device=DEV, batch_size=B))
    sd = O.synthetic_state(cfg, seed=4)
    model = nets.get_model((3, H, W), cfg, DEV)
    model.load_state_dict(sd)
    model.train()
    opt = FlatAdam(model.parameters(), lr=1e-3)
    batches = []
    for seed in (1, 2):
        xyz, normals, imus = O.synthetic_batch(B, S, H, W, T, seed=seed)
        batches.append({"xyz": xyz.to(DEV), "normals": normals.to(DEV), "imus": imus.to(DEV)})

    def fwd_loss(d):
        pos, ori = model([[d["xyz"], d["normals"]], d["imus"]])
        return (pos ** 2).sum() + (ori ** 2).sum()
    return model, opt, batches, fwd_loss


@pytest.fixture(autouse=True)
def dedicated_stream():
    """The train loop (eager steps, capture, replays) runs on one non-default stream: GraphedTrainStep's docstring."""
    with torch.cuda.stream(torch.cuda.Stream()):
        yield
    torch.cuda.synchronize()


def test_graph_replay_matches_eager_gradients():
    from deeplio_b200 import _lib as L
    from deeplio_b200.graph import GraphedTrainStep
    model, opt, batches, fwd_loss = _setup(no_dropout=True)
    eager = []
    for d in batches:
        opt.zero_grad()
        loss = fwd_loss(d)
        loss.backward()
        eager.append((float(loss), opt.flat_grad.clone()))
    del loss
    step = GraphedTrainStep(fwd_loss, batches[0], opt.zero_grad)
    assert step.captured_launches > 50
    n0 = L.launch_count()
    for d, (eloss, egrad) in zip(batches, eager):
        loss = step(d)
        torch.cuda.synchronize()
        # same kernels on the same data; fp64 statistics atomics may sum in another order (1e-6-level noise)
        assert abs(float(loss) - eloss) <= 1e-5 * max(1.0, abs(eloss))
        assert (opt.flat_grad - egrad).abs().max().item() <= 2e-4 * egrad.abs().max().item()
    assert L.launch_count() == n0, "a replay must not go through the C ABI again"


def test_graph_replays_draw_fresh_dropout_masks():
    from deeplio_b200.graph import GraphedTrainStep
    model, opt, batches, fwd_loss = _setup(no_dropout=False)
    step = GraphedTrainStep(fwd_loss, batches[0], opt.zero_grad)
    losses = []
    for _ in range(3):
        losses.append(float(step(batches[0])))
        torch.cuda.synchronize()
    # BN running statistics do not enter a train-mode forward, so only the dropout masks can change the loss
    assert len({round(v, 9) for v in losses}) == 3, losses


def test_split_backward_two_graphs_equal_the_unsplit_step():
    """The N > 1 step of bench.py on one GPU: the backward pass cut at the encoders' feature vectors
    (``split_backward``), graph A = forward + loss + backward of everything downstream, graph B = the encoders'
    backward, a callback in between (where the downstream all-reduce is fired).  Same gradients as the eager,
    unsplit step; both input slots work."""
    from deeplio_b200.graph import GraphedTrainStep
    model, opt, batches, fwd_loss = _setup(no_dropout=True)
    eager = []
    for d in batches:
        opt.zero_grad()
        loss = fwd_loss(d)
        loss.backward()
        eager.append((float(loss), opt.flat_grad.clone()))
    del loss
    lidar = model.lidar_feat_net
    lidar.split_backward = True
    try:
        step = GraphedTrainStep(fwd_loss, batches[0], opt.zero_grad, model=model, second_backward=lidar.backward_encoders)
        assert len(step.graphs_b) == len(step.graphs) == 2
        calls = []
        for rep in range(2):
            for d, (eloss, egrad) in zip(batches, eager):
                loss = step(d, between=lambda: calls.append(opt.flat_grad.abs().sum().item()))
                torch.cuda.synchronize()
                assert abs(float(loss) - eloss) <= 1e-5 * max(1.0, abs(eloss))
                assert (opt.flat_grad - egrad).abs().max().item() <= 2e-5 * egrad.abs().max().item()
        # the callback ran between the two graphs: the encoders' gradients were still zero there
        enc = [p for k, p in model.named_parameters() if ".encoder" in k]
        assert len(calls) == 4 and all(c > 0 for c in calls) and all(p.grad.abs().max().item() > 0 for p in enc[:2])
    finally:
        lidar.split_backward = False
