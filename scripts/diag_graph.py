"""Diagnostic: does whole-step CUDA graph capture work (a) from a fresh model, (b) after eager steps?"""
import sys, os, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_graph import _setup
from deeplio_b200.graph import GraphedTrainStep

def acc_ids(model):
    out = []
    for p in list(model.parameters())[:3] + list(model.parameters())[-3:]:
        out.append(id(p.view_as(p).grad_fn.next_functions[0][0]))
    return out

main = torch.cuda.Stream()
torch.cuda.set_stream(main)
for eager_first in (False, True):
    model, opt, batches, fwd_loss = _setup(no_dropout=True)
    if eager_first:
        a = acc_ids(model)
        for d in batches:
            opt.zero_grad(); loss = fwd_loss(d); loss.backward()
        b = acc_ids(model)
        del loss
        gc.collect()
        c = acc_ids(model)
        print("acc ids stable across eager step:", a == b, "after del/gc:", b == c)
    try:
        step = GraphedTrainStep(fwd_loss, batches[0], opt.zero_grad)
        l = step(batches[1]); torch.cuda.synchronize()
        print("eager_first=%s: capture OK, loss %.6f, launches %d" % (eager_first, float(l), step.captured_launches))
    except Exception as e:
        print("eager_first=%s: capture FAILED: %s" % (eager_first, str(e).splitlines()[0]))
        break
