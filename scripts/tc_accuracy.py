"""Accuracy of the tcgen05 3xTF32 convolution vs fp64 for growing reduction depth K (GPU box)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplio_b200 import _lib as L  # noqa: E402
from tests.test_gpu_ops import DEV, _st, from_nhwc, split_padded  # noqa: E402

for (n, cin, cout, h, w, k) in [(2, 32, 64, 16, 32, 1), (2, 64, 128, 16, 32, 3), (2, 128, 128, 16, 32, 3),
                                (2, 256, 256, 16, 32, 3), (2, 512, 512, 17, 65, 3), (2, 1024, 256, 8, 16, 3)]:
    g = torch.Generator().manual_seed(1)
    for dist in ("randn", "relu"):
        x = torch.randn(n, cin, h, w, generator=g)
        if dist == "relu":
            x = x.relu()
        wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        p = (k - 1) // 2
        ref64 = F.conv2d(x.double(), wt.double(), None, 1, p)
        ref32 = F.conv2d(x, wt, None, 1, p)
        x_hi, x_lo = split_padded(L, x, p, p)
        wd = wt.to(DEV)
        w_hi, w_lo = torch.empty(cout, k, k, cin, device=DEV), torch.empty(cout, k, k, cin, device=DEV)
        L.weight_to_ohwi(wd.data_ptr(), cout, cin, k, k, cin, w_hi.data_ptr(), w_lo.data_ptr(), _st())
        y = torch.empty(n, h, w, cout, device=DEV)
        xt4, yt4 = L.Tensor4(n, h, w, cin, p, p), L.Tensor4(n, h, w, cout, 0, 0)
        cv = L.Conv(k, k, 1, 1, p, p)
        L.conv2d_fwd(xt4, x_hi.data_ptr(), x_lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), None, cv, 0, yt4,
                     y.data_ptr(), None, _st())
        ysimt = torch.empty(n, h, w, cout, device=DEV)
        L.conv2d_fwd(xt4, x_hi.data_ptr(), None, w_hi.data_ptr(), None, None, cv, 0, yt4, ysimt.data_ptr(), None, _st())
        got, gs = from_nhwc(y).double(), from_nhwc(ysimt).double()
        sc = ref64.abs().max().item()
        e = (got - ref64)
        print("K=%5d %-5s  tc: max %.2e rms %.2e mean(signed) %+.2e | simt: max %.2e | torch-cpu32: max %.2e" % (
            cin * k * k, dist, e.abs().max().item() / sc, e.pow(2).mean().sqrt().item() / sc, e.mean().item() / sc,
            (gs - ref64).abs().max().item() / sc, (ref32.double() - ref64).abs().max().item() / sc))
