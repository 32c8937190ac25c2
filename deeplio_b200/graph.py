"""Whole-step CUDA graph: forward, loss and backward of one train step captured once and replayed.

A train step of this path is several hundred to ~1400 kernel launches (PointSeg) issued from Python through ctypes;
at small per-GPU batches the host cannot issue them as fast as the GPU retires them.  Everything in the step is
static -- shapes, the straight-line encoder programs, the arena the gradients are written into -- so the step is
captured into ONE graph (torch.cuda.graph: the caching allocator gives the capture a private pool, so every
intermediate keeps its address across replays) and replayed with a single launch.  Two things in the step are not
static and are handled explicitly: dropout masks (their seed is mixed with a counter in device memory that the
graph itself increments, ``functional._seed_epoch``) and the optimizer / gradient exchange, which stay outside the
graph (Adam's bias correction depends on the step number; the all-reduce overlaps through a hook when eager).

The reference has no counterpart (it launches op by op from Python, trainer.py:238-272).
"""
import gc

import torch

from . import _lib as L
from . import functional as Fn


class GraphedTrainStep:
    """``step(inputs) -> loss`` with forward + loss + backward replayed from a CUDA graph.

    ``fwd_loss(inputs) -> loss`` runs the model and the loss on a dict of device tensors; ``example`` is one such dict
    (shapes and dtypes fix the graph).  ``zero_grad`` is called inside the graph before the forward pass (the
    FlatAdam arena memset).  After ``step`` the parameter gradients are in place; the caller all-reduces and calls
    the optimizer as usual.

    Streams (torch autograd): the AccumulateGrad nodes of the parameters remember the stream they were created on
    and can outlive the step that created them; if eager steps ran on the DEFAULT stream before the capture, their
    nodes run on the legacy stream during the captured backward, which invalidates the capture (measured: capture from
    a fresh model works, after two eager default-stream steps it fails).  So run the train loop -- eager steps, this
    constructor, the replays -- on ONE dedicated stream (``with torch.cuda.stream(torch.cuda.Stream())``): the capture
    then happens on the current stream.  When the current stream is the default stream, a side stream is used and no
    eager step may have run before."""

    def __init__(self, fwd_loss, example, zero_grad, warmup=3, model=None, slots=2, second_backward=None):
        """``model`` (optional): its buffers (BatchNorm running statistics and ``num_batches_tracked``) and the dropout
        call counter are snapshotted before the warm-up passes and restored before the capture, so that building the
        graph leaves the training state exactly as an eager run would find it.

        ``slots``: number of static input sets, one captured graph each (``input_slots``).  An input pipeline that
        writes batch i + 1 into one set while the step on batch i reads the other (pipeline.DevicePrefetcher with
        ``buffers=step.input_slots``) then feeds the graphs without a device-to-device copy of the inputs.  The
        graphs share one memory pool (they are replayed one at a time, on one stream), so the activations exist once.

        ``second_backward`` (optional callable): the step is captured as TWO graphs per slot -- (A) zero_grad, forward,
        loss, ``loss.backward()`` and (B) ``second_backward()`` -- for a model whose autograd graph is cut in two
        (``LidarFeatNet.split_backward``: (B) is the encoders' backward).  ``step(inputs, between=f)`` calls ``f()``
        between the two replays: a data-parallel step launches the all-reduce of the gradients (A) completed there,
        so that it runs under (B)."""
        dev = next(iter(example.values())).device
        self.input_slots = [{k: torch.empty_like(v) for k, v in example.items()} for _ in range(max(1, slots))]
        for slot in self.input_slots:
            for k, v in example.items():
                slot[k].copy_(v)
        self.static_in = self.input_slots[0]
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.graphs, self.graphs_b, self.static_losses = [], [], []
        self._next = 0
        gc.collect()       # drop autograd graphs of earlier eager steps that are only kept alive by reference cycles
        prev = Fn._seed_epoch[0]
        Fn._seed_epoch[0] = self.epoch
        try:
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev) if cur == torch.cuda.default_stream(dev) else cur
            side.wait_stream(cur)
            saved = [(b, b.clone()) for b in model.buffers()] if model is not None else []
            drop0 = Fn._drop_counter[0]
            with torch.cuda.stream(side):                      # warm-up off the default stream, as capture requires
                for _ in range(warmup):
                    zero_grad()
                    fwd_loss(self.static_in).backward()
                    if second_backward is not None:
                        second_backward()
                with torch.no_grad():
                    for b, v in saved:
                        b.copy_(v)
            del saved
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            pool = None
            for slot in self.input_slots:
                Fn._drop_counter[0] = drop0            # every graph draws the same mask streams (seed + device epoch)
                graph = torch.cuda.CUDAGraph()
                n0 = L.launch_count()
                # capture on the warm-up stream: autograd nodes that outlive an iteration (AccumulateGrad) stay on one stream
                with torch.cuda.graph(graph, stream=side, pool=pool):
                    self.epoch.add_(1)
                    zero_grad()
                    loss = fwd_loss(slot)
                    loss.backward()
                pool = graph.pool()
                if second_backward is not None:
                    graph_b = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph_b, stream=side, pool=pool):
                        second_backward()
                    self.graphs_b.append(graph_b)
                self.captured_launches = L.launch_count() - n0     # library kernels every replay launches
                self.graphs.append(graph)
                self.static_losses.append(loss)
            self.graph, self.static_loss = self.graphs[0], self.static_losses[0]
            Fn._drop_counter[0] = drop0                # the capture drew its seeds; no step has run yet
        finally:
            Fn._seed_epoch[0] = prev

    def __call__(self, inputs, between=None):
        """``inputs``: one of ``input_slots`` (replayed in place) or any dict of tensors with the captured shapes
        (copied into the next slot first).  ``between``: called between the two graphs of a split step."""
        k = next((i for i, s in enumerate(self.input_slots) if inputs is s), None)
        if k is None:
            k = self._next
            self._next = (k + 1) % len(self.input_slots)
            for name, v in inputs.items():
                if v.data_ptr() != self.input_slots[k][name].data_ptr():
                    self.input_slots[k][name].copy_(v, non_blocking=True)
        self.graphs[k].replay()
        if self.graphs_b:
            if between is not None:
                between()
            self.graphs_b[k].replay()
        return self.static_losses[k].detach()
