import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import case_setup, load_golden
from tests.test_gpu_model import build_b200_model, to_dev
from deeplio_b200 import engine as E

rec = load_golden("pointseg_lstm_rnn")
cfg, sd, inputs = case_setup(rec)
res = {}
for tc in (False, True):
    E.USE_TC = tc
    model = build_b200_model(cfg, rec["H"], rec["W"], rec["B"], sd)
    model.train()
    pos, ori = model(to_dev(inputs))
    ((pos ** 2).sum() + (ori ** 2).sum()).backward()
    res[tc] = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}
    res[tc]["pos"] = pos.detach().cpu()
print("pos diff", (res[True]["pos"] - res[False]["pos"]).abs().max().item())
names = [k for k in res[True] if k.startswith("lidar_feat_net.encoder1.fire_blk5") or k.startswith("lidar_feat_net.encoder1.fire_blk4.1")]
for k in names:
    a, b = res[True][k], res[False][k]
    d = (a - b).abs()
    print("%-58s max|g| %.2e maxdiff %.2e" % (k, b.abs().max().item(), d.max().item()), end="")
    if a.dim() == 4:
        per = d.flatten(1).max(1).values
        top = torch.topk(per, 3)
        print("  worst out-ch", top.indices.tolist(), ["%.1e" % v for v in top.values.tolist()], end="")
    print()
