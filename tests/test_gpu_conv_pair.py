"""CTA-pair convolution kernel (tcgen05 cta_group::2, csrc/conv_tc.cu conv_tc2_kernel) against the single-CTA kernel
and against F.conv2d in fp64 (dlio_set_option("conv_cg2", 0 / 1) switches between them; the pair kernel is the default)."""
import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_ops import DEV, _lib, _st, from_nhwc, pack_f16, relerr

pytestmark = pytest.mark.gpu

CASES = [
    # n, cin, cout, h, w, (kh, kw), bias, act
    (2, 64, 128, 9, 33, (3, 5), True, 1),
    (1, 128, 128, 8, 17, (3, 3), False, 0),
    (3, 512, 512, 17, 65, (3, 3), True, 1),       # Simple-1 conv7: four N tiles, 15 M tiles of 256 rows
    (1, 128, 256, 6, 11, (3, 3), False, 0),       # fewer rows than one 256-row tile per image
    (2, 64, 128, 64, 513, (3, 5), True, 1),       # Simple-1 conv2 at its real size
    (1, 256, 256, 33, 129, (3, 3), True, 1),      # conv5
    (2, 768, 128, 6, 10, (1, 1), True, 0),        # 1x1, K = 12 chunks
    (2, 128, 64, 9, 33, (3, 3), True, 1),         # 64-channel output tiles: each CTA stages 32 weight rows
    (1, 64, 64, 20, 70, (3, 3), False, 0),        # ResNet layer1
    (2, 128, 64, 64, 513, (3, 5), False, 0),      # the shape of Simple-1 conv2's dgrad
]


@pytest.mark.parametrize("case", CASES)
def test_conv_cta_pair_equals_single_cta(case):
    L = _lib()
    n, cin, cout, h, w, (kh, kw), bias, act = case
    g = torch.Generator().manual_seed(cin * 3 + cout + h)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g) if bias else None
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    ref = F.conv2d(x.double(), wt.double(), b.double() if bias else None, 1, (ph, pw))
    if act:
        ref = F.relu(ref)
    x_h2, x_b, _ = pack_f16(L, x, ph, pw)
    wd = wt.to(DEV)
    w_b = torch.empty(1, device=DEV)
    w_h2 = torch.empty(cout, 2, kh * kw * cin, dtype=torch.float16, device=DEV)
    L.weight_pack_f16(wd.data_ptr(), cout, cin, kh, kw, cin, 0, 1, w_b.data_ptr(), w_h2.data_ptr(), _st())
    xt4, yt4 = L.Tensor4(n, h, w, cin, ph, pw), L.Tensor4(n, h, w, cout, 0, 0)
    cv = L.Conv(kh, kw, 1, 1, ph, pw)
    outs = []
    for pair in (0, 1):
        L.set_option(b"conv_cg2", pair)
        try:
            y = torch.full((n, h, w, cout), float("nan"), device=DEV)
            stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
            L.conv2d_fwd_f16(xt4, x_h2.data_ptr(), x_b.data_ptr(), w_h2.data_ptr(), w_b.data_ptr(),
                             b.to(DEV).data_ptr() if bias else None, cv, act, yt4, y.data_ptr(), stats.data_ptr(), _st())
            torch.cuda.synchronize()
        finally:
            L.set_option(b"conv_cg2", 1)
        outs.append((y, stats))
    (y0, s0), (y1, s1) = outs
    assert relerr(from_nhwc(y1).double(), ref) < 1e-5
    assert torch.equal(y0, y1)                              # same K order, same accumulator segments per output
    assert torch.allclose(s0, s1, rtol=1e-12, atol=1e-9)    # fp64 atomics: order only
    # sums of up to 65 k values of both signs: the bar is relative to the sum of magnitudes
    assert torch.allclose(s1.cpu()[:cout], ref.sum((0, 2, 3)), rtol=5e-5,
                          atol=max(1e-4, 2e-6 * ref.abs().sum((0, 2, 3)).max().item()))


WGRAD_CASES = [
    # n, cin, cout, h, w, (kh, kw): Cin % 128 == 0 and Cout % 256 == 0 take the pair kernel
    (2, 128, 256, 6, 11, (3, 3)),                  # fewer K chunks than pipeline stages
    (1, 128, 256, 64, 257, (3, 3)),                # Simple-1 conv4 (one image)
    (3, 512, 512, 17, 65, (3, 3)),                 # conv7
    (1, 256, 512, 33, 129, (3, 3)),                # conv6
    (2, 256, 256, 9, 20, (1, 1)),                  # one tap
    (1, 128, 256, 1, 3, (3, 3)),                   # a handful of pixels: most K splits are empty
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad_cta_pair_equals_single_cta(case):
    """wgrad_tc2_kernel (256 x 128 tiles of dw per CTA pair) against wgrad_tc_kernel and torch fp64.  The K split
    differs between the two (74 pairs against 148 CTAs per wave), so the fp32 atomics combine other partial sums:
    equal to accumulation round-off, not bitwise."""
    L = _lib()
    n, cin, cout, h, w, (kh, kw) = case
    g = torch.Generator().manual_seed(cin + 5 * cout + h)
    x = torch.randn(n, cin, h, w, generator=g)
    dy = torch.randn(n, cout, h, w, generator=g)
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    wr = torch.zeros(cout, cin, kh, kw, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wr, None, 1, (ph, pw)).backward(dy.double())
    x_h2, x_b, _ = pack_f16(L, x, ph, pw)
    dy_h2, dy_b, _ = pack_f16(L, dy, ph, pw)
    xt4, dyt4 = L.Tensor4(n, h, w, cin, ph, pw), L.Tensor4(n, h, w, cout, ph, pw)
    cv = L.Conv(kh, kw, 1, 1, ph, pw)
    outs = []
    for pair in (0, 1):
        L.set_option(b"conv_cg2", pair)
        try:
            dw = torch.full((cout, kh, kw, cin), float("nan"), device=DEV)
            L.conv2d_bwd_weight_f16(xt4, x_h2.data_ptr(), x_b.data_ptr(), dyt4, dy_h2.data_ptr(), dy_b.data_ptr(), cv,
                                    dw.data_ptr(), _st())
            torch.cuda.synchronize()
        finally:
            L.set_option(b"conv_cg2", 1)
        outs.append(dw.permute(0, 3, 1, 2).cpu().double())
    assert relerr(outs[1], wr.grad) < 1e-5
    assert relerr(outs[1], outs[0]) < 2e-6
