"""Per-tensor gradient errors of the B200 path at the reference's default 57 x 720 image size, against an fp64 run of
the oracle, next to the fp32 oracle's own distance from fp64 (conditioning)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import deeplio_oracle as O
from oracle.configs import make_cfg
from deeplio_b200 import nets
from deeplio_b200.config import build_config_container

def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-300)

h, w, B, S, T = 57, 720, int(os.environ.get("B", "1")), 2, 15
for kw in [dict(lidar="lidar-feat-simple-1"), dict(lidar="lidar-feat-pointseg"),
           dict(lidar="lidar-feat-resnet", rnn_type="gru", lidar_fusion="cat"), dict(lidar="lidar-feat-flownet")]:
    cfg = make_cfg(height=h, width=w, seq=S, odom_hidden=64, **kw)
    sd = O.synthetic_state(cfg, seed=33)
    inputs = O.synthetic_batch(B, S, h, w, T, seed=33)
    build_config_container(cfg, argparse.Namespace(device="cuda:0", batch_size=B))
    model = nets.get_model((3, h, w), cfg, "cuda:0")
    model.load_state_dict(sd)
    model.train()
    xyz, normals, imus = inputs
    pos, ori = model([[xyz.cuda(), normals.cuda()], imus.cuda()])
    ((pos ** 2).sum() + (ori ** 2).sum()).backward()
    res = {}
    for dt in (torch.float32, torch.float64):
        leaves = {k: ((v.to(dt) if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point() and "running_" not in k))
                  for k, v in sd.items()}
        p, o = O.deeplio_forward(leaves, cfg, *[t.to(dt) for t in inputs], training=True)
        ((p ** 2).sum() + (o ** 2).sum()).backward()
        res[dt] = (p.detach(), o.detach(), {k: v.grad for k, v in leaves.items() if v.requires_grad and v.grad is not None})
    p64, o64, g64 = res[torch.float64]
    p32, o32, g32 = res[torch.float32]
    print("== %s: fwd err b200 %.2e / %.2e   oracle32 %.2e / %.2e" % (kw["lidar"], rel(pos.cpu(), p64), rel(ori.cpu(), o64), rel(p32, p64), rel(o32, o64)))
    rows = []
    for k, p in model.named_parameters():
        if k not in g64 or g64[k].abs().max() == 0:
            continue
        rows.append((rel(p.grad.cpu(), g64[k]), rel(g32[k], g64[k]), k))
    rows.sort(reverse=True)
    bad = [r for r in rows if r[0] > 1e-3]
    print("   %d tensors, %d above 1e-3; worst:" % (len(rows), len(bad)))
    for e, e32, k in rows[:8]:
        print("   %-55s b200 %.2e   oracle32 %.2e" % (k, e, e32))
