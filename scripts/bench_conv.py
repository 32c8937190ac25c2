"""Per-layer timing of the tensor-core convolution kernels (forward, dgrad, wgrad) on the Simple-1 layer shapes
of the benchmark workload (16 images per encoder), through the C ABI, with CUDA events on the launching stream.
Prints ms, fp32-equivalent TFLOP/s and the time the tcgen05 pipe alone would need (3 MMAs per product).
usage: python scripts/bench_conv.py [f16|tf32] [N images]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplio_b200 import _lib as L  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
DEV = "cuda:0"
LAYERS = [("conv2", 64, 128, 64, 513, 3, 5), ("conv3", 128, 128, 64, 257, 3, 3), ("conv4", 128, 256, 64, 257, 3, 3),
          ("conv5", 256, 256, 33, 129, 3, 3), ("conv6", 256, 512, 33, 129, 3, 3), ("conv7", 512, 512, 17, 65, 3, 3)]
st = torch.cuda.current_stream().cuda_stream
PEAK_F16 = 2.25e15 * 1965 / 1965   # nominal dense fp16 FLOP/s; TF32 is half


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print("%-6s %-22s %8s %8s %8s   %s" % ("layer", "shape", "fwd ms", "dgrad", "wgrad", "TFLOP/s fp32-equiv (fwd/dgrad/wgrad), MMA-only ms"))
for name, cin, cout, h, w, kh, kw in LAYERS:
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    hp, wp = h + 2 * ph, w + 2 * pw
    x = torch.zeros(N, hp, wp, cin, device=DEV)
    x[:, ph:ph + h, pw:pw + w] = torch.randn(N, h, w, cin, device=DEV)
    dy = torch.zeros(N, hp, wp, cout, device=DEV)
    dy[:, ph:ph + h, pw:pw + w] = torch.randn(N, h, w, cout, device=DEV)
    wt = torch.randn(cout, cin, kh, kw, device=DEV) / (cin * kh * kw) ** 0.5
    y = torch.empty(N, h, w, cout, device=DEV)
    dx = torch.empty(N, h, w, cin, device=DEV)
    dw = torch.empty(cout, kh, kw, cin, device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    xt4, yt4 = L.Tensor4(N, h, w, cin, ph, pw), L.Tensor4(N, h, w, cout, 0, 0)
    dyt4, dxt4 = L.Tensor4(N, h, w, cout, ph, pw), L.Tensor4(N, h, w, cin, 0, 0)
    cv = L.Conv(kh, kw, 1, 1, ph, pw)
    if mode == "f16":
        xh = torch.empty(N, hp, wp, 2, cin, dtype=torch.float16, device=DEV)
        dyh = torch.empty(N, hp, wp, 2, cout, dtype=torch.float16, device=DEV)
        xb, dyb, wb = (torch.empty(1, device=DEV) for _ in range(3))
        L.pack_f16(x.data_ptr(), N * hp * wp, cin, xb.data_ptr(), xh.data_ptr(), st)
        L.pack_f16(dy.data_ptr(), N * hp * wp, cout, dyb.data_ptr(), dyh.data_ptr(), st)
        wh = torch.empty(cout, 2, kh * kw * cin, dtype=torch.float16, device=DEV)
        wth = torch.empty(cin, 2, kh * kw * cout, dtype=torch.float16, device=DEV)
        L.weight_pack_f16(wt.data_ptr(), cout, cin, kh, kw, cin, 0, 1, wb.data_ptr(), wh.data_ptr(), st)
        L.weight_pack_f16(wt.data_ptr(), cout, cin, kh, kw, cin, 1, 0, wb.data_ptr(), wth.data_ptr(), st)
        fwd = lambda: L.conv2d_fwd_f16(xt4, xh.data_ptr(), xb.data_ptr(), wh.data_ptr(), wb.data_ptr(), None, cv, 1, yt4,
                                       y.data_ptr(), stats.data_ptr(), st)
        dgrad = lambda: L.conv2d_bwd_data_f16(dyt4, dyh.data_ptr(), dyb.data_ptr(), wth.data_ptr(), wb.data_ptr(), cv, dxt4,
                                              dx.data_ptr(), 0, st)
        wgrad = lambda: L.conv2d_bwd_weight_f16(xt4, xh.data_ptr(), xb.data_ptr(), dyt4, dyh.data_ptr(), dyb.data_ptr(), cv,
                                                dw.data_ptr(), st)
    else:
        def split(t):
            lo = torch.empty_like(t)
            hi = torch.empty_like(t)
            n_, hp_, wp_, c_ = t.shape
            t4 = L.Tensor4(n_, hp_, wp_, c_, 0, 0)
            L.bn_act_pool_fwd(t4, t.data_ptr(), None, None, t4, None, L.BnPool(0, 0, 1, 1, 1, 0), t4, hi.data_ptr(),
                              lo.data_ptr(), None, None, None, None, st)
            return hi, lo
        xhi, xlo = split(x)
        dyhi, dylo = split(dy)
        whi, wlo = torch.empty(cout, kh, kw, cin, device=DEV), torch.empty(cout, kh, kw, cin, device=DEV)
        L.weight_to_ohwi(wt.data_ptr(), cout, cin, kh, kw, cin, whi.data_ptr(), wlo.data_ptr(), st)
        wthi, wtlo = torch.empty(cin, kh, kw, cout, device=DEV), torch.empty(cin, kh, kw, cout, device=DEV)
        L.weight_flip_transpose(whi.data_ptr(), cout, cin, kh, kw, wthi.data_ptr(), wtlo.data_ptr(), st)
        fwd = lambda: L.conv2d_fwd(xt4, xhi.data_ptr(), xlo.data_ptr(), whi.data_ptr(), wlo.data_ptr(), None, cv, 1, yt4,
                                   y.data_ptr(), stats.data_ptr(), st)
        dgrad = lambda: L.conv2d_bwd_data(dyt4, dyhi.data_ptr(), dylo.data_ptr(), whi.data_ptr(), wlo.data_ptr(),
                                          wthi.data_ptr(), wtlo.data_ptr(), cv, dxt4, dx.data_ptr(), 0, st)
        wgrad = lambda: L.conv2d_bwd_weight(xt4, xhi.data_ptr(), xlo.data_ptr(), dyt4, dyhi.data_ptr(), dylo.data_ptr(), cv,
                                            dw.data_ptr(), st)
    flop = 2.0 * N * h * w * cout * cin * kh * kw
    t = [timed(fwd), timed(dgrad), timed(wgrad)]
    for k, v in zip(tot, t):
        tot[k] += v
    mma_ms = 3 * flop / (PEAK_F16 if mode == "f16" else PEAK_F16 / 2) * 1e3
    print("%-6s %-22s %8.3f %8.3f %8.3f   %6.0f %6.0f %6.0f   %.3f" % (
        name, "%dx%dx%d %d->%d" % (N, h, w, cin, cout), t[0], t[1], t[2], flop / t[0] * 1e-9, flop / t[1] * 1e-9,
        flop / t[2] * 1e-9, mma_ms))
    del x, dy, y, dx
print("total  fwd %.3f  dgrad %.3f  wgrad %.3f ms  (x2 encoders per step)" % (tot["fwd"], tot["dgrad"], tot["wgrad"]))
