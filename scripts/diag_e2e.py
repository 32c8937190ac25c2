"""Where does the end-to-end loop lose time against the device-resident loop?  Times 10 graph-replayed steps with
(a) resident inputs, (b) + device-to-device copy into the graph's static buffers, (c) + host->device prefetch,
(d) + lagged loss read (= bench.py's e2e)."""
import argparse, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplio_b200 import functional as Fn, nets, parallel
from deeplio_b200.config import build_config_container
from deeplio_b200.graph import GraphedTrainStep
from deeplio_b200.optim import FlatAdam
from deeplio_b200.pipeline import DevicePrefetcher, LaggedScalar
from deeplio_b200.workloads import workload_config

dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(dev))
H, W = 64, 2048
cfg, B, S, T = workload_config("cfg1_simple1_lstm_b8", H, W)
build_config_container(cfg, argparse.Namespace(device=str(dev), batch_size=B))
torch.manual_seed(1234)
model = nets.get_model((3, H, W), cfg, str(dev)).train()
opt = FlatAdam(model.parameters(), lr=1e-3, weight_decay=1e-4)
g = torch.Generator().manual_seed(100)
host = {"pairs": torch.randn(B, S, 2, 6, H, W, generator=g).pin_memory(), "imus": torch.randn(B, S, T, 6, generator=g).pin_memory(),
        "gt_pos": torch.randn(B, S, 3, generator=g).pin_memory(), "gt_ori": torch.randn(B, S, 3, generator=g).pin_memory()}
resident = {k: v.to(dev) for k, v in host.items()}
other = {k: v.clone() for k, v in resident.items()}

def fwd_loss(d):
    p = d["pairs"]
    pos, ori = model([[p[:, :, :, 0:3], p[:, :, :, 3:].contiguous()], d["imus"]])
    return Fn.hws_loss(pos, ori, d["gt_pos"], d["gt_ori"])

step = GraphedTrainStep(fwd_loss, resident, opt.zero_grad)

def train(d):
    loss = step(d)
    opt.step(1.0)
    return loss

def timed(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def a(n):
    for _ in range(n): train(step.static_in)
def b(n):
    for _ in range(n): train(other)
def c(n):
    for d in DevicePrefetcher(itertools.repeat(host, n), dev): train(d)
def d_(n):
    lag = LaggedScalar()
    for d in DevicePrefetcher(itertools.repeat(host, n), dev): lag.push(train(d))
    lag.flush()
def e(n):
    lag = LaggedScalar()
    for _ in range(n): lag.push(train(other))
    lag.flush()
for name, fn in (("warm", a), ("a resident", a), ("b +d2d", b), ("c +prefetch", c), ("d +lagged read", d_), ("e d2d+lagged", e),
                 ("c20 +prefetch, 20 steps", lambda n: c(20)), ("a again", a)):
    ms = timed(fn) if "20" not in name else timed(fn) / 2
    print("%-28s %.3f ms/step" % (name, ms))
