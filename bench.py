#!/usr/bin/env python
"""Benchmark of the DeepLIO training hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--batch B]

One "step" = one full train step of BASELINE.json configs[1] (Simple-1 LiDAR net + bi-LSTM IMU net + soft fusion +
LSTM odometry net, 64x2048 frames, batch 8 per GPU, S = 2 frame pairs per sample) as the reference's Trainer runs it
(trainer.py:197-281): ground-truth pairing -> pairing of the frames -> forward -> pose chaining se(3) -> SE(3) -> HWS
loss (local + global, learnable sx / sq) -> backward -> gradient all-reduce (N > 1) -> Adam; dropout active,
BatchNorm in train mode.  Everything is deeplio_b200 code: data.ground_truth / PairedFrames, nets.get_model,
pose.se3_to_SE3, losses.HWSLoss, optim.FlatAdam.

Two timed passes of the same K steps: eager launches with the library's per-call CUDA events (the per-class breakdown
in ``roofline.classes``; one stream, so every class is timed in isolation), then forward + loss + backward replayed
from CUDA graphs (``value``; DLIO_GRAPH=0 keeps the eager number).  ``e2e`` feeds every step from pinned host memory
([B, S+1, 6, H, W] frames -- not pairs --, IMU windows, ground-truth poses) through ``deeplio_b200.pipeline`` and reads
the loss back: the pipeline is primed with one batch before the timed region (its steady state), inside the region every
one of the K steps issues one host->device copy of a full batch and one device->host read.  Prints ONE JSON line on rank 0.

``--impl reference``: the reference's CPU path -- the oracle restatement (oracle/; the reference is pure Python /
PyTorch, does not travel to the GPU box and has nothing to compile) -- on all host cores, SAME workload, batch and
step counts (steps are only cut when the run would exceed ~4 minutes; the line says so).
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
G0, G1 = 1, 3                         # global-loss slice 1 .. max_glob_seq (trainer.py:42,261)
STD = (0.1269, 0.0951, 0.0108, 0.3436, 0.4445, 0.5664)    # per-channel std of the mean-subtracted images (config.yaml:26-27)


def workload_dict(workload, B, S, T, world, n_params):
    """The workload-defining part of ``config``: identical in the B200 arm and the reference arm."""
    return {"workload": workload, "per_gpu_batch": B, "global_batch": B * world, "pairs_per_sample": S,
            "image": "64x2048x6 per frame (xyz, normals), %d frames per sample" % (S + 1), "imu_window": T,
            "loss": "HWSLoss local+global, learnable sx/sq", "optimizer": "adam lr 1e-3 wd 1e-4",
            "params_M": round(n_params / 1e6, 3)}


def synthetic_host_batch(B, S, T, seed):
    """per-channel N(0, sigma_c) frames with ~15 % empty pixels, N(0,1) IMU windows, smooth trajectories."""
    import torch
    from deeplio_b200.workloads import synthetic_gts
    g = torch.Generator().manual_seed(seed)
    std = torch.tensor(STD).view(1, 1, 6, 1, 1)
    frames = torch.randn(B, S + 1, 6, H, W, generator=g) * std
    frames *= (torch.rand(B, S + 1, 1, H, W, generator=g) >= 0.15).float()
    return {"frames": frames, "imus": torch.randn(B, S, T, 6, generator=g), "gts": synthetic_gts(B, S + 1, seed)}


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
def cpu_train_steps(workload, steps, warmup, batch, budget_s=None):
    """The reference's CPU path (oracle restatement, fp32 torch CPU, all host threads): full train steps of the same
    model and the same step definition (ground-truth pairing, forward, pose chaining, HWS loss, backward, Adam).
    Returns (frame-pairs/s, seconds per step, cores, steps actually timed, warm-up steps actually run)."""
    import torch
    from oracle import deeplio_oracle as O
    from oracle import pose_oracle as P
    from oracle.configs import BASELINE_CONFIGS, make_cfg
    kw, _, seq, t_imu = BASELINE_CONFIGS[workload]
    cfg = make_cfg(no_dropout=False, height=H, width=W, **kw)
    combos = cfg["datasets"]["combinations"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.synthetic_state(cfg, seed=0)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k}
    state = dict(sd)
    state.update(leaves)
    sx, sq = torch.tensor(0.0, requires_grad=True), torch.tensor(-3.0, requires_grad=True)
    opt = torch.optim.Adam([{"params": list(leaves.values())}, {"params": [sx, sq]}], lr=LR, weight_decay=WD)
    d = synthetic_host_batch(batch, seq, t_imu, seed=100)
    idx = torch.tensor(combos)
    times, done_warm = [], 0
    i = 0
    while len(times) < steps:
        t0 = time.perf_counter()
        opt.zero_grad()
        gt_f2f, gt_f2g = P.ground_truth(d["gts"], combos)
        pairs = d["frames"][:, idx]                                           # misc.py:65-69
        pos, ori = O.deeplio_forward(state, cfg, pairs[:, :, :, 0:3], pairs[:, :, :, 3:].contiguous(), d["imus"],
                                     training=True)
        p, q, _ = P.se3_to_SE3(pos, ori)
        loss = P.pose_loss(pos, ori, p[:, G0:G1], q[:, G0:G1], gt_f2f[:, :, 0:3], gt_f2f[:, :, 3:],
                           gt_f2g[:, G0:G1, 0:3], gt_f2g[:, G0:G1, 3:7], sx=sx, sq=sq)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        else:
            done_warm += 1
            if budget_s is not None and i == 0:
                # bound the whole run: keep >= 3 timed steps and >= 1 warm-up step
                fit = int(budget_s / max(dt, 1e-3))
                if fit < steps + warmup:
                    warmup = max(1, min(warmup, fit // 4))
                    steps = max(3, min(steps, fit - warmup))
        i += 1
    sec = sum(times) / len(times)
    return batch * seq / sec, sec, cores, len(times), done_warm, sum(p.numel() for p in leaves.values()) + 2


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.configs import BASELINE_CONFIGS
    workload = args.workload or WORKLOAD
    _, B, S, T = BASELINE_CONFIGS[workload]
    B = args.batch or (B if workload == WORKLOAD else max(1, B // 8))
    fps, sec, cores, steps, warm, n_params = cpu_train_steps(workload, args.steps, args.warmup, B, budget_s=240.0)
    cut = "" if (steps == args.steps and warm == args.warmup) else \
        " (cut from %d / %d so that the run stays within ~4 minutes)" % (args.steps, args.warmup)
    sample = "batch %d x S=%d (%d frame pairs of 64x2048) per step, %d timed steps after %d warm-up%s" % (
        B, S, B * S, steps, warm, cut)
    cfg = workload_dict(workload, B, S, T, 1, n_params)
    cfg["impl_note"] = "CPU restatement of the reference path (oracle/), torch CPU fp32, %d threads" % cores
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from deeplio_b200 import _lib, data, engine as E, losses, nets, parallel, pose
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
    # measurement-only switch (never the default): dgrad / wgrad with the hi*hi product alone
    single_bwd = os.environ.get("DLIO_BWD_SINGLE", "0") == "1"
    if single_bwd:
        _lib.set_option(b"bwd_single_pass", 1)
    workload = args.workload or WORKLOAD
    cfg, B, S, T = workload_config(workload, H, W)
    B = args.batch or (B if workload == WORKLOAD else max(1, B // 8))      # per-GPU batch (weak scaling)
    combos = cfg["datasets"]["combinations"]
    build_config_container(cfg, argparse.Namespace(device=str(dev), batch_size=B))
    torch.manual_seed(1234)
    model = nets.get_model((3, H, W), cfg, str(dev))
    criterion = losses.get_loss_function(cfg, str(dev))              # HWSLoss, local+global, learnable sx / sq
    parallel.broadcast_model(model)
    model.train()
    # trainer.py:56-58: two parameter groups, the model's and the criterion's
    opt = FlatAdam([{"params": model.parameters()}, {"params": criterion.parameters()}], lr=LR, weight_decay=WD)
    n_params = sum(p.numel() for p in model.parameters()) + sum(p.numel() for p in criterion.parameters())
    # N > 1: the backward pass is cut at the encoders' feature vector; the gradients of everything downstream (78 % of
    # the bytes: odometry LSTM, fusion, heads, IMU net, fc1, sx / sq) are all-reduced while the encoders' backward runs
    lidar_net = getattr(model, "lidar_feat_net", None)
    split = world > 1 and lidar_net is not None and os.environ.get("DLIO_SPLIT_BWD", "1") != "0"
    reducer = None
    if world > 1:
        late = [criterion] + ([model.imu_feat_net, lidar_net.fc1] if split else [])
        reducer = parallel.OverlappedGradReducer(model, opt, extra_late=[m for m in late if m is not None])
        if split:
            lidar_net.split_backward = True
            model.on_head_grads_ready = None       # fired explicitly after loss.backward()
        # measurement only: the step without any gradient exchange (what the max over N unequal GPUs alone costs)
        reducer.disabled = os.environ.get("DLIO_NO_EXCHANGE", "0") == "1"

    host = {k: v.pin_memory() for k, v in synthetic_host_batch(B, S, T, seed=100 + rank).items()}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}
    state = {"resident": resident}

    def fwd_loss(d):
        gt_f2f, gt_f2g = data.ground_truth(d["gts"], combos)                 # misc.py:83-125 on the device
        xyz = data.PairedFrames(d["frames"], combos, 0, 3)                  # misc.py:65-69 without the gather copy
        normals = data.PairedFrames(d["frames"], combos, 3, 3)
        pos, ori = model([[xyz, normals], d["imus"]])
        p, q = pose.se3_to_SE3(pos, ori, check=False)                        # trainer.py:245
        return criterion(pos, ori, p[:, G0:G1], q[:, G0:G1], gt_f2f[:, :, 0:3], gt_f2f[:, :, 3:],
                         gt_f2g[:, G0:G1, 0:3], gt_f2g[:, G0:G1, 3:7])      # trainer.py:260-263

    def eager_step(d):
        opt.zero_grad()
        loss = fwd_loss(d)
        loss.backward()
        if split:
            reducer.fire()
            lidar_net.backward_encoders()
        scale = reducer.finish() if reducer is not None else 1.0
        opt.step(scale)
        return loss.detach()      # no reference to the autograd graph survives the step (deeplio_b200.graph)

    gstep = [None]

    def graph_step(d):
        # forward + loss + backward: graph launches (two per step when the backward pass is split)
        loss = gstep[0](d, between=reducer.fire if split else None)
        opt.step(reducer.finish() if reducer is not None else 1.0)
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
            train_step(state["resident"])

    # pass 1 times every kernel class in isolation (one stream); the graph pass may put the two encoders on two streams
    enc_streams = E.ENC_STREAMS
    E.ENC_STREAMS = False
    E.FLOPS = {}
    train_step(resident)                          # also counts the algorithmic conv FLOPs of one step, per class
    flops, E.FLOPS = E.FLOPS, None
    for _ in range(max(0, args.warmup - 1)):
        train_step(resident)
    assert pose.raise_for_status(dev) == 0
    torch.cuda.reset_peak_memory_stats(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.profile_enable(1)
    n0 = _lib.launch_count()
    ms_eager = timed(resident_steps, args.steps)
    launches = _lib.launch_count() - n0
    prof = _lib.profile_read()
    _lib.profile_enable(0)
    ms_total, launch_mode = ms_eager, "eager (one C-ABI call per kernel)"
    # pass 2, the same step with forward + loss + backward replayed from CUDA graphs (deeplio_b200.graph); the
    # gradient exchange and the optimizer stay outside.  `value` is this pass; DLIO_GRAPH=0 keeps pass 1.
    if os.environ.get("DLIO_GRAPH", "1") != "0":
        try:
            from deeplio_b200.graph import GraphedTrainStep
            if reducer is not None:
                model.on_head_grads_ready = None       # under the graph nothing fires from a hook
            E.ENC_STREAMS = enc_streams
            gstep[0] = GraphedTrainStep(fwd_loss, resident, opt.zero_grad, model=model,
                                        second_backward=lidar_net.backward_encoders if split else None)
            state["resident"] = gstep[0].input_slots[0]     # the graph's own input buffers: resident steps copy nothing
            for _ in range(args.warmup):
                train_step(state["resident"])
            n0 = _lib.launch_count()
            ms_total = timed(resident_steps, args.steps)
            launches = _lib.launch_count() - n0 + gstep[0].captured_launches * args.steps
            launch_mode = "cuda-graph (forward + loss + backward: %d library kernels per replay%s; all-reduce and Adam eager)" % (
                gstep[0].captured_launches, ", split at the encoder features with the downstream all-reduce under the encoders' backward" if split else "")
        except Exception as e:      # capture is an optimisation: report the eager numbers and say why
            gstep[0] = None
            launch_mode = "eager (CUDA graph capture failed: %s)" % str(e).splitlines()[0][:160]
            print("bench.py: CUDA graph capture failed, eager timings stand: %r" % (e,), file=sys.stderr)

    # end to end through the public API: every step's inputs come from pinned host memory (copied on a copy stream
    # one step ahead, deeplio_b200.pipeline.DevicePrefetcher) and every step's loss is read back on the host (one
    # step late, LaggedScalar); all copies and reads happen inside the timed region
    losses_seen = []

    def primed_prefetcher(steps):
        """The input pipeline in its steady state: the copy of the first batch is in flight when the loop starts, as it
        is at any step of a long epoch.  Inside the timed region every step then waits for its batch, issues the copy
        of the next one (K copies for K steps), trains and reads the previous step's loss back."""
        bufs = gstep[0].input_slots if gstep[0] is not None else None     # copy straight into the graphs' inputs
        return DevicePrefetcher(itertools.repeat(host, steps + 1), dev, buffers=bufs)

    def e2e_steps(steps, pf=None):
        lag = LaggedScalar()
        pf = primed_prefetcher(steps) if pf is None else pf
        for _ in range(steps):
            losses_seen.append(lag.push(train_step(next(pf))))
        losses_seen.append(lag.flush())
    e2e_steps(2)
    pf = primed_prefetcher(args.steps)
    ms_e2e = timed(lambda k: e2e_steps(k, pf), args.steps)
    assert all(v is None or v == v for v in losses_seen), "non-finite loss in the end-to-end run"
    assert pose.raise_for_status(dev) == 0, "pose / ground-truth status flags raised"
    sampler.stop_flag = True
    sampler.join(timeout=2)
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30

    pairs_per_step = B * S * world
    value = pairs_per_step * args.steps / (ms_total / 1e3)
    e2e = pairs_per_step * args.steps / (ms_e2e / 1e3)

    # roofline of the dominant kernel class: the conv class with the largest share of the step.  Algorithmic FLOPs
    # per class are counted by the executor while it runs the step (engine.FLOPS): 2 * Cout * Ho * Wo * Cin * kh * kw
    # per image, real channel counts -- no padding, no split-precision factor, none of the extra MACs of the
    # space-to-depth / pixel-pair views.
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
        if flops.get(name):
            classes[name]["gflop_per_step"] = flops[name] / 1e9
            classes[name]["tflops"] = flops[name] / (per_step_ms * 1e-3) / 1e12 if per_step_ms > 0 else None
    roofline = None
    cands = [k for k in classes if classes[k].get("tflops")]
    if cands:
        dom = max(cands, key=lambda k: classes[k]["ms_per_step"])
        c = classes[dom]
        tc = dom.endswith("_tc")
        # the tcgen05 kernels issue 3 fp16 MMAs per fp32-accurate product (hi*hi, lo*hi, hi*lo), and fp16 runs at the
        # bf16 rate, so the tensor-pipe ceiling for fp32-equivalent algorithmic FLOPs is bf16_peak / 3
        # (frac_of_3xf16_ceiling); `frac` is against the measured bf16 peak itself, as the contract asks.
        # traffic: dram bytes per launch of this class from the committed ncu pass (profiles/*_traffic.json)
        traffic = None
        for f in ("r02_traffic.json", "r01_traffic.json"):
            try:
                t = json.load(open(os.path.join(ROOT, "profiles", f)))
                if workload == WORKLOAD and dom in t:
                    traffic = t[dom].get("dram_bytes_per_launch")
                    break
            except Exception:
                pass
        conv_ms = sum(classes[k]["ms_per_step"] for k in cands)
        conv_fl = sum(flops[k] for k in cands)
        roofline = {"bound": "tensor", "kernel": dom, "achieved": c["tflops"], "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": c["tflops"] / bf16_peak, "traffic": traffic, "peak_source": peak_src,
                    "frac_of_3xf16_ceiling": (c["tflops"] / (bf16_peak / 3.0)) if tc else None,
                    "math": ("3xF16 tcgen05: three kind::f16 MMAs per fp32-accurate product, fp32 accumulation "
                             "(fp32-equivalent algorithmic FLOPs)") if tc else "fp32 FMA (CUDA cores)",
                    "launch_ms": c["ms_per_step"] / c["launches_per_step"],
                    "conv_path": {"gflop_per_step": conv_fl / 1e9, "ms_per_step": conv_ms,
                                  "tflops": conv_fl / (conv_ms * 1e-3) / 1e12,
                                  "frac_of_3xf16_ceiling": conv_fl / (conv_ms * 1e-3) / 1e12 / (bf16_peak / 3.0)},
                    "classes": classes}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and workload == WORKLOAD:
        fps, sec, cores, st, wm, _ = cpu_train_steps(workload, 3, 1, 2)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "batch 2 x S=2 (4 frame pairs of 64x2048) per step, %d timed steps after %d warm-up, %.2f s/step "
                                  "(the --impl reference arm runs the full batch)" % (st, wm, sec)}

    if rank == 0:
        cfgd = workload_dict(workload, B, S, T, world, n_params)
        cfgd.update({"parallelism": "dp%d" % world, "launch": launch_mode, "eager_ms_per_step": ms_eager / args.steps,
                     "encoder_streams": 2 if E.ENC_STREAMS else 1,
                     **({"ab": "bwd_single_pass: dgrad / wgrad at plain fp16 operand accuracy -- NOT the parity-tested "
                               "path, not a bench value"} if single_bwd else {}),
                     **({"ab": "no gradient exchange (DLIO_NO_EXCHANGE=1): replicas drift apart -- measures the "
                               "max-over-ranks cost alone, not a bench value"}
                        if (reducer is not None and reducer.disabled) else {}),
                     "l2": "working set per step (%.1f GB peak, activations) exceeds the 126 MB L2; no explicit flush" % peak_gb})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                        "pipeline": "one batch in flight when the timed region starts; K copies, K steps, K loss reads inside",
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
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's per-GPU batch)")
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
