"""Input pipeline of the train loop: pinned-host batches are copied to the device on a copy stream one step
ahead of the compute stream, and per-step scalars (the loss) are read back one step late, so neither the
host->device copy nor the device->host read stalls the kernels of the step in flight.

The reference does both synchronously (``DataCombiCreater`` moves every batch with blocking ``.to(device)``,
deeplio/models/misc.py:38-44; ``loss.item()`` right after the optimizer step, trainer.py:281-283).
"""
import torch


class DevicePrefetcher:
    """Iterates over ``batches`` (an iterable of dicts name -> pinned CPU tensor, all with the same shapes) and
    yields dicts of device tensors.  Two sets of device buffers; the copy of batch i+1 runs on ``copy_stream``
    while the caller computes on batch i.  A yielded dict is valid until the caller asks for the batch after
    the next one (its buffers are then overwritten)."""

    def __init__(self, batches, device, depth=2, buffers=None):
        """``buffers`` (optional): a list of >= 2 dicts of device tensors to copy into instead of buffers of its own --
        e.g. ``graph.GraphedTrainStep.input_slots``, so that the host batch lands directly in a captured graph's
        static inputs and the yielded dict IS that slot."""
        self.it = iter(batches)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher: device must be a CUDA device (no CPU fallback)")
        self.copy_stream = torch.cuda.Stream(self.device)
        if buffers is not None:
            if len(buffers) < 2:
                raise ValueError("DevicePrefetcher: at least two buffer sets are needed to copy one batch ahead")
            depth = len(buffers)
        self.depth = depth
        self.bufs = list(buffers) if buffers is not None else [None] * depth
        if buffers is not None:
            # the sets may hold the previous user's data in flight on the current stream
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        self.ready = [torch.cuda.Event() for _ in range(depth)]      # copy of slot k finished
        self.free = [None] * depth                                   # compute on slot k finished
        self.bytes_per_batch = 0
        self.slot = 0
        self.pending = None
        self._issue()

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.pending = None
            return
        k = self.slot
        if self.bufs[k] is None:
            self.bufs[k] = {n: torch.empty(t.shape, dtype=t.dtype, device=self.device) for n, t in host.items()}
        if not self.bytes_per_batch:
            self.bytes_per_batch = sum(t.numel() * t.element_size() for t in host.values())
        with torch.cuda.stream(self.copy_stream):
            if self.free[k] is not None:
                self.copy_stream.wait_event(self.free[k])            # the step that read this slot is done
            for n, t in host.items():
                self.bufs[k][n].copy_(t, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.pending = k
        self.slot = (k + 1) % self.depth

    def __iter__(self):
        return self

    def __next__(self):
        if self.pending is None:
            raise StopIteration
        k = self.pending
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self.ready[k])
        # slot k may be overwritten once everything the caller enqueues before asking for batch k + depth is done;
        # the event is recorded when the NEXT batch is requested, i.e. after the caller's step on this one
        prev = (k - 1) % self.depth
        if self.bufs[prev] is not None:
            ev = torch.cuda.Event()
            ev.record(main)
            self.free[prev] = ev
        self._issue()
        return self.bufs[k]


class LaggedScalar:
    """Device scalar -> host float with one step of delay: ``push(t)`` enqueues an asynchronous copy of ``t`` into
    pinned memory and returns the value pushed one call earlier (None the first time); ``flush()`` returns the
    last one.  The host never waits for the step it has just launched."""

    def __init__(self):
        self.host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.n = 0

    def _read(self, i):
        self.events[i].synchronize()
        return float(self.host[i][0])

    def push(self, t):
        i = self.n & 1
        self.host[i].copy_(t.detach().reshape(1), non_blocking=True)
        self.events[i].record()
        self.n += 1
        return self._read(1 - i) if self.n > 1 else None

    def flush(self):
        return self._read((self.n - 1) & 1) if self.n else None
