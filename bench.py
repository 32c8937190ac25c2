#!/usr/bin/env python
"""Benchmark of the DeepLIO training hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one full train step of BASELINE.json configs[1] (Simple-1 LiDAR net + bi-LSTM IMU net + soft
fusion + LSTM odometry net, 64x2048 frame pairs, batch 8 per GPU, S = 2 pairs per sample): forward -> HWS
loss -> backward -> gradient all-reduce (N > 1) -> Adam, dropout active, BatchNorm in train mode.
Two timed passes of the same K steps: eager launches with the library's per-call CUDA events (the per-class
breakdown in ``roofline.classes``, ``config.eager_ms_per_step``), then forward + loss + backward replayed from one
CUDA graph (``value``; DLIO_GRAPH=0 keeps the eager number).  ``e2e`` feeds every step from pinned host memory
through ``deeplio_b200.pipeline`` (copy stream one step ahead, loss read one step late).
Prints ONE JSON line on rank 0.  ``--impl reference`` times the CPU restatement of the reference (the
reference is pure Python / PyTorch and does not travel to the GPU box) on the host cores.
"""
import argparse
import itertools
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frame-pairs/sec (train step)"
UNIT = "frame-pairs/s"
WORKLOAD = "cfg1_simple1_lstm_b8"     # BASELINE.json configs[1]
H, W = 64, 2048
LR, WD = 1e-3, 1e-4                   # reference defaults (deeplio/train.py:39,45)


# ----------------------------------------------------------------------------- workload description
def conv_flops_simple1(n_images):
    """Algorithmic conv FLOPs of ONE Simple-1 encoder over n_images 64x2048 images, split by
    the kernel class that runs each layer: {class: flops}.  Every layer runs on tcgen05: conv2..conv7 (stride 1,
    Cin % 64 == 0) as 3xF16, conv1 (Cin = 6, 5x7 stride (1,2)) as 3xTF32 through the space-to-depth view
    (csrc/conv_s2d.cu; its 2.3x extra MACs are NOT counted -- these are algorithmic FLOPs).  conv1 needs no dgrad
    (the input has no gradient)."""
    from deeplio_b200.engine import pool_out
    spec = [(6, 64, 5, 7, (1, 2), (1, 2)), (64, 128, 3, 5, (1, 1), (1, 2)), (128, 128, 3, 3, (1, 1), None),
            (128, 256, 3, 3, (1, 1), (2, 2)), (256, 256, 3, 3, (1, 1), None), (256, 512, 3, 3, (1, 1), (2, 2)),
            (512, 512, 3, 3, (1, 1), None)]
    h, w = H, W
    out = {k: 0.0 for k in ("conv_fwd_simt", "conv_dgrad_simt", "conv_wgrad_simt", "conv_fwd_tc", "conv_dgrad_tc",
                            "conv_wgrad_tc")}
    for i, (ci, co, kh, kw, (sh, sw), pool) in enumerate(spec):
        ho, wo = (h + 2 * ((kh - 1) // 2) - kh) // sh + 1, (w + 2 * ((kw - 1) // 2) - kw) // sw + 1
        f = 2.0 * co * ho * wo * ci * kh * kw * n_images
        tc = ((sh, sw) == (1, 1) and ci % 32 == 0) or i == 0
        out["conv_fwd_tc" if tc else "conv_fwd_simt"] += f
        out["conv_wgrad_tc" if (tc and (co % 128 == 0 or i == 0)) else "conv_wgrad_simt"] += f
        if i > 0:
            out["conv_dgrad_tc" if ((sh, sw) == (1, 1) and co % 32 == 0 and ci % 16 == 0) else "conv_dgrad_simt"] += f
        h, w = ho, wo
        if pool:
            h, w = pool_out(h, pool[0], True), pool_out(w, pool[1], True)
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}   # NVML clocks-event-reason bits
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_train_steps(steps, warmup, batch=1, seq=2, t_imu=15):
    """The reference's CPU path (oracle restatement, fp32 torch CPU, all host threads): full train steps of the
    same model on a bounded sample (batch ``batch``).  Returns (frame-pairs/s, seconds per step, cores)."""
    import torch
    from oracle import deeplio_oracle as O
    from oracle.configs import BASELINE_CONFIGS, make_cfg
    kw, _, _, _ = BASELINE_CONFIGS[WORKLOAD]
    cfg = make_cfg(no_dropout=False, height=H, width=W, **kw)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.synthetic_state(cfg, seed=0)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k}
    state = dict(sd)
    state.update(leaves)
    opt = torch.optim.Adam(list(leaves.values()), lr=LR, weight_decay=WD)
    xyz, normals, imus = O.synthetic_batch(batch, seq, H, W, t_imu, seed=0)
    g = torch.Generator().manual_seed(5)
    gt_pos, gt_ori = torch.randn(batch, seq, 3, generator=g) * 0.1, torch.randn(batch, seq, 3, generator=g) * 0.01
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        pos, ori = O.deeplio_forward(state, cfg, xyz, normals, imus, training=True)
        mse = torch.nn.functional.mse_loss
        loss = mse(pos, gt_pos) * 1.0 + mse(ori, gt_ori) * float(torch.exp(torch.tensor(3.0))) - 3.0
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch * seq / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    fps, sec, cores = cpu_train_steps(steps, warmup)
    sample = "batch 1 x S=2 (2 frame pairs of 64x2048) per step, %d timed steps after %d warm-up" % (steps, warmup)
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path (oracle/), torch CPU fp32"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from deeplio_b200 import _lib, functional as Fn, nets, parallel
    from deeplio_b200.config import build_config_container
    from deeplio_b200.optim import FlatAdam
    from deeplio_b200.pipeline import DevicePrefetcher, LaggedScalar
    from deeplio_b200.workloads import workload_config

    rank, local_rank, world = parallel.init_from_env()
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 through torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.device_check(local_rank)
    # the whole train loop runs on ONE dedicated stream: the CUDA-graph capture below needs every autograd node that
    # outlives a step to belong to the capturing stream, never to the legacy default stream (deeplio_b200/graph.py)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    workload = args.workload or WORKLOAD
    cfg, B, S, T = workload_config(workload, H, W)
    B = args.batch or B                       # per-GPU batch (weak scaling)
    build_config_container(cfg, argparse.Namespace(device=str(dev), batch_size=B))
    torch.manual_seed(1234)
    model = nets.get_model((3, H, W), cfg, str(dev))
    parallel.broadcast_model(model)
    model.train()
    opt = FlatAdam(model.parameters(), lr=LR, weight_decay=WD)
    n_params = sum(p.numel() for p in model.parameters())
    # N > 1: the all-reduce of the odometry-net / fusion / head gradients (78 % of the bytes) overlaps the encoders'
    # backward; DLIO_OVERLAP=0 falls back to one all-reduce after backward
    reducer = parallel.OverlappedGradReducer(model, opt) if os.environ.get("DLIO_OVERLAP", "1") != "0" else None

    # synthetic batch of this rank: pinned host copy + device-resident copy
    g = torch.Generator().manual_seed(100 + rank)
    std = torch.tensor([0.1269, 0.0951, 0.0108, 0.3436, 0.4445, 0.5664]).view(1, 1, 6, 1, 1)
    frames = torch.randn(B, S + 1, 6, H, W, generator=g) * std
    frames *= (torch.rand(B, S + 1, 1, H, W, generator=g) >= 0.15).float()
    idx = torch.tensor([[i, i + 1] for i in range(S)])
    host = {"pairs": frames[:, idx].contiguous().pin_memory(),            # [B,S,2,6,H,W]
            "imus": torch.randn(B, S, T, 6, generator=g).pin_memory(),
            "gt_pos": (torch.randn(B, S, 3, generator=g) * 0.1).pin_memory(),
            "gt_ori": (torch.randn(B, S, 3, generator=g) * 0.01).pin_memory()}
    h2d_bytes = sum(t.numel() * 4 for t in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}

    def split(d):
        pairs = d["pairs"]
        return [[pairs[:, :, :, 0:3], pairs[:, :, :, 3:].contiguous()], d["imus"]]   # as misc.py:65-69 does

    def fwd_loss(d):
        pos, ori = model(split(d))
        return Fn.hws_loss(pos, ori, d["gt_pos"], d["gt_ori"])

    def eager_step(d):
        opt.zero_grad()
        loss = fwd_loss(d)
        loss.backward()
        scale = reducer.finish() if reducer is not None else parallel.allreduce_grads(opt.flat_grad)
        opt.step(scale)
        return loss.detach()      # no reference to the autograd graph survives the step (deeplio_b200.graph)

    gstep = [None]

    def graph_step(d):
        loss = gstep[0](d)                       # forward + loss + backward: one graph launch
        opt.step(parallel.allreduce_grads(opt.flat_grad))
        return loss

    def train_step(d):
        return graph_step(d) if gstep[0] is not None else eager_step(d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """``fn(steps)`` enqueues ``steps`` train steps; device time between two events on the launching stream"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident_steps(steps):
        for _ in range(steps):
            train_step(resident)

    # pass 1 times every kernel class in isolation (one stream); the graph pass may put the two encoders on two streams
    from deeplio_b200 import engine as E
    enc_streams = E.ENC_STREAMS
    E.ENC_STREAMS = False
    for _ in range(args.warmup):
        train_step(resident)
    torch.cuda.reset_peak_memory_stats(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # pass 1, eager launches with the library's per-call CUDA events: the per-class breakdown (roofline.classes)
    _lib.profile_enable(1)
    n0 = _lib.launch_count()
    ms_eager = timed(resident_steps, args.steps)
    launches = _lib.launch_count() - n0
    prof = _lib.profile_read()
    _lib.profile_enable(0)
    ms_total, launch_mode = ms_eager, "eager (one C-ABI call per kernel)"
    # pass 2, the same step with forward + loss + backward replayed from ONE CUDA graph (deeplio_b200.graph); the
    # gradient exchange and the optimizer stay outside the graph.  `value` is this pass; DLIO_GRAPH=0 keeps pass 1.
    if os.environ.get("DLIO_GRAPH", "1") != "0":
        try:
            from deeplio_b200.graph import GraphedTrainStep
            if reducer is not None:
                model.on_head_grads_ready = None
                reducer = None
            E.ENC_STREAMS = enc_streams
            gstep[0] = GraphedTrainStep(fwd_loss, resident, opt.zero_grad, model=model)
            for _ in range(args.warmup):
                train_step(resident)
            n0 = _lib.launch_count()
            ms_total = timed(resident_steps, args.steps)
            launches = _lib.launch_count() - n0 + gstep[0].captured_launches * args.steps
            launch_mode = "cuda-graph (forward + loss + backward: %d library kernels per replay; all-reduce and Adam eager)" % gstep[0].captured_launches
        except Exception as e:      # capture is an optimisation: report the eager numbers and say why
            gstep[0] = None
            launch_mode = "eager (CUDA graph capture failed: %s)" % str(e).splitlines()[0][:160]
            print("bench.py: CUDA graph capture failed, eager timings stand: %r" % (e,), file=sys.stderr)

    # end to end through the public API: every step's inputs come from pinned host memory (copied on a copy stream
    # one step ahead, deeplio_b200.pipeline.DevicePrefetcher) and every step's loss is read back on the host (one
    # step late, LaggedScalar); all copies and reads happen inside the timed region
    losses = []

    def e2e_steps(steps):
        lag = LaggedScalar()
        for d in DevicePrefetcher(itertools.repeat(host, steps), dev):
            losses.append(lag.push(train_step(d)))
        losses.append(lag.flush())
    e2e_steps(2)
    ms_e2e = timed(e2e_steps, args.steps)
    assert all(v is None or v == v for v in losses), "non-finite loss in the end-to-end run"
    sampler.stop_flag = True
    sampler.join(timeout=2)
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30

    pairs_per_step = B * S * world
    value = pairs_per_step * args.steps / (ms_total / 1e3)
    e2e = pairs_per_step * args.steps / (ms_e2e / 1e3)

    # roofline of the dominant kernel class (largest share of the step among the profiled conv classes)
    # per-class FLOP accounting exists for the headline workload's encoder (Simple-1); other workloads (secondary
    # runs, --workload) report class times only
    flops = {k: 2 * v for k, v in conv_flops_simple1(B * S).items()} if "simple1" in workload else {}    # two encoders
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (dense bf16 cuBLAS, the only measured tensor peak)"
    if not bf16_peak:
        bf16_peak, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained bf16)"
    classes = {}
    for name, (ms, n) in prof.items():
        per_step_ms = ms / args.steps
        classes[name] = {"ms_per_step": per_step_ms, "launches_per_step": n / args.steps,
                         "share_of_step": per_step_ms / (ms_eager / args.steps)}
        if name in flops:
            classes[name]["gflop_per_step"] = flops[name] / 1e9
            classes[name]["tflops"] = flops[name] / (per_step_ms * 1e-3) / 1e12 if per_step_ms > 0 else None
    roofline = None
    if classes and flops:
        dom = max((k for k in classes if k in flops), key=lambda k: classes[k]["ms_per_step"])
        c = classes[dom]
        tc = dom.endswith("_tc")
        # the tcgen05 kernels issue 3 fp16 MMAs per fp32-accurate product (hi*hi, lo*hi, hi*lo), and fp16 runs at the
        # bf16 rate, so the tensor-pipe ceiling for fp32-equivalent algorithmic FLOPs is bf16_peak / 3
        # (frac_of_3xf16_ceiling); `frac` is against the measured bf16 peak itself, as the contract asks.
        # traffic: dram bytes per launch of this class from the committed ncu pass (profiles/r01_traffic.json)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": dom, "achieved": c["tflops"], "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": c["tflops"] / bf16_peak, "traffic": traffic, "peak_source": peak_src,
                    "frac_of_3xf16_ceiling": (c["tflops"] / (bf16_peak / 3.0)) if tc else None,
                    "math": ("3xF16 tcgen05: three kind::f16 MMAs per fp32-accurate product, fp32 accumulation "
                             "(fp32-equivalent algorithmic FLOPs)") if tc else "fp32 FMA (CUDA cores)",
                    "launch_ms": c["ms_per_step"] / c["launches_per_step"], "classes": classes}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and workload == WORKLOAD:
        fps, sec, cores = cpu_train_steps(3, 1)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "batch 1 x S=2 (2 frame pairs) per step, 3 timed steps after 1 warm-up, %.2f s/step" % sec}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "per_gpu_batch": B, "global_batch": B * world, "pairs_per_sample": S,
                           "image": "64x2048x6 x2 (xyz, normals)", "imu_window": T, "parallelism": "dp%d" % world,
                           "launch": launch_mode, "eager_ms_per_step": ms_eager / args.steps,
                           "encoder_streams": 2 if E.ENC_STREAMS else 1,
                           "params_M": n_params / 1e6, "optimizer": "adam lr 1e-3 wd 1e-4 (fused, flat arena)",
                           "l2": "working set per step (%.1f GB peak, activations) exceeds the 126 MB L2; no explicit flush" % peak_gb},
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu_baseline}
        if roofline is None:
            line["classes"] = classes
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's, 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="", help="secondary runs: another BASELINE.json workload "
                    "(deeplio_b200.workloads.WORKLOADS); the default is the headline workload")
    args = ap.parse_args()
    # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout must carry the one JSON line only
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    # stdout carries the ONE JSON line and nothing else: libraries (NCCL prints its version banner from C, whatever
    # NCCL_DEBUG says on some builds) get stderr as their fd 1 for the whole run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        sys.stdout.flush()


if __name__ == "__main__":
    main()
