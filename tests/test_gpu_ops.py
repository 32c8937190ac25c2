"""Per-kernel numerics: each C-ABI op against a plain PyTorch fp32 CPU computation of the same op.

Tolerances (stated per test) are fp32 round-off bounds: the kernels accumulate in fp32 (BN statistics in
fp64), in a different order than ATen's CPU kernels.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _lib():
    from deeplio_b200 import _lib as L
    return L


def _st():
    return torch.cuda.current_stream().cuda_stream


def to_padded_nhwc(x, ph, pw, c_pad=None):
    """NCHW cpu -> padded NHWC cuda (zero pads / zero extra channels)."""
    n, c, h, w = x.shape
    c_pad = c_pad or c
    out = torch.zeros(n, h + 2 * ph, w + 2 * pw, c_pad)
    out[:, ph:ph + h, pw:pw + w, :c] = x.permute(0, 2, 3, 1)
    return out.to(DEV).contiguous()


def from_nhwc(t):
    return t.cpu().permute(0, 3, 1, 2).contiguous()


def relerr(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


CONV_CASES = [
    # n, cin, cout, h, w, (kh, kw), (sh, sw), bias, act
    (2, 6, 64, 12, 40, (5, 7), (1, 2), True, 1),
    (2, 64, 128, 9, 33, (3, 5), (1, 1), True, 1),
    (1, 128, 128, 8, 17, (3, 3), (1, 1), False, 0),
    (2, 32, 48, 10, 21, (3, 3), (2, 2), False, 0),
    (2, 64, 16, 7, 19, (1, 1), (1, 1), True, 0),
    (1, 16, 64, 6, 10, (1, 1), (1, 2), False, 0),
    (1, 256, 80, 5, 9, (3, 5), (1, 2), True, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case):
    L = _lib()
    n, cin, cout, h, w, (kh, kw), (sh, sw), bias, act = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g) if bias else None
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    ref = F.conv2d(x, wt, b, (sh, sw), (ph, pw))
    if act:
        ref = F.relu(ref)
    ho, wo = ref.shape[2:]
    cin_pad = (cin + 3) // 4 * 4
    xp = to_padded_nhwc(x, ph, pw, cin_pad)
    wd = wt.to(DEV)
    w_ohwi = torch.empty(cout, kh, kw, cin_pad, device=DEV)
    L.weight_to_ohwi(wd.data_ptr(), cout, cin, kh, kw, cin_pad, w_ohwi.data_ptr(), None, _st())
    assert torch.equal(w_ohwi[..., :cin].cpu(), wt.permute(0, 2, 3, 1))
    y = torch.empty(n, ho, wo, cout, device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    xt4, yt4 = L.Tensor4(n, h, w, cin_pad, ph, pw), L.Tensor4(n, ho, wo, cout, 0, 0)
    cv = L.Conv(kh, kw, sh, sw, ph, pw)
    L.conv2d_fwd(xt4, xp.data_ptr(), None, w_ohwi.data_ptr(), None, b.to(DEV).data_ptr() if bias else None, cv, act,
                 yt4, y.data_ptr(), stats.data_ptr(), _st())
    got = from_nhwc(y)
    assert relerr(got, ref) < 2e-6                                   # fp32 accumulation-order noise
    s = stats.cpu()
    assert torch.allclose(s[:cout], ref.double().sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(s[cout:], (ref.double() ** 2).sum((0, 2, 3)), rtol=1e-5, atol=1e-4)

    # backward (no activation in the backward test: dy is the gradient at the conv output)
    dy = torch.randn(n, cout, ho, wo, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    F.conv2d(xr, wr, None, (sh, sw), (ph, pw)).backward(dy)
    dyd = to_padded_nhwc(dy, 0, 0)
    dx = torch.empty(n, h, w, cin_pad, device=DEV)
    L.conv2d_bwd_data(yt4, dyd.data_ptr(), None, w_ohwi.data_ptr(), None, None, None, cv,
                      L.Tensor4(n, h, w, cin_pad, 0, 0), dx.data_ptr(), 0, _st())
    assert relerr(from_nhwc(dx)[:, :cin], xr.grad) < 3e-6
    # accumulate: dx += the gradient
    L.conv2d_bwd_data(yt4, dyd.data_ptr(), None, w_ohwi.data_ptr(), None, None, None, cv,
                      L.Tensor4(n, h, w, cin_pad, 0, 0), dx.data_ptr(), 1, _st())
    assert relerr(from_nhwc(dx)[:, :cin], 2 * xr.grad) < 3e-6
    dw_ohwi = torch.empty(cout, kh, kw, cin_pad, device=DEV)
    L.conv2d_bwd_weight(xt4, xp.data_ptr(), None, yt4, dyd.data_ptr(), None, cv, dw_ohwi.data_ptr(), _st())
    dw = torch.empty(cout, cin, kh, kw, device=DEV)
    L.weight_grad_to_oihw(dw_ohwi.data_ptr(), cout, cin, kh, kw, cin_pad, dw.data_ptr(), _st())
    assert relerr(dw.cpu(), wr.grad) < 1e-5                          # split-K atomics: order varies


def split_padded(L, x, ph, pw):
    """NCHW cpu -> (full, lo) padded NHWC planes, produced by the library's own copy kernel."""
    n, c, h, w = x.shape
    src = to_padded_nhwc(x, 0, 0)
    hi = torch.empty(n, h + 2 * ph, w + 2 * pw, c, device=DEV)
    lo = torch.empty_like(hi)
    t_src, t_dst = L.Tensor4(n, h, w, c, 0, 0), L.Tensor4(n, h, w, c, ph, pw)
    L.bn_act_pool_fwd(t_src, src.data_ptr(), None, None, t_src, None, L.BnPool(0, 0, 1, 1, 1, 0), t_dst, hi.data_ptr(),
                      lo.data_ptr(), None, None, None, None, _st())
    return hi, lo


def pack_f16(L, x, ph, pw):
    """NCHW cpu -> (packed fp16 hi|lo planes [n, hp, wp, 2, c], bound[1]) through dlio_pack_f16."""
    src = to_padded_nhwc(x, ph, pw)
    n, hp, wp, c = src.shape
    h2 = torch.empty(n, hp, wp, 2, c, dtype=torch.float16, device=DEV)
    bound = torch.full((1,), float("nan"), device=DEV)
    L.pack_f16(src.data_ptr(), n * hp * wp, c, bound.data_ptr(), h2.data_ptr(), _st())
    return h2, bound, src


F16_CASES = [
    # n, cin, cout, h, w, (kh, kw), bias, act, extra tensor pad, magnitude of x, magnitude of dy
    (2, 64, 128, 9, 33, (3, 5), True, 1, 0, 1.0, 1.0),
    (1, 128, 128, 8, 17, (3, 3), False, 0, 0, 1.0, 1.0),
    (1, 256, 192, 5, 9, (3, 3), True, 1, 0, 3e-6, 2e4),
    (2, 128, 256, 6, 11, (3, 3), False, 0, 0, 7e3, 1e-7),
    (3, 512, 512, 17, 65, (3, 3), True, 1, 0, 1.0, 1e-3),
    (2, 768, 64, 6, 10, (1, 1), True, 0, 1, 1.0, 1.0),
    (4, 64, 256, 8, 16, (3, 3), True, 0, 0, 40.0, 1.0),
    (1, 64, 64, 20, 70, (3, 3), False, 0, 0, 1.0, 1.0),
    # real layer shapes of the BASELINE.json workloads (one or two 64x2048-derived images): Simple-1 conv2 (wgrad in
    # swapped-operand mode, Cin = 64), conv4, conv6 -- split-K wgrad over 34 k .. 68 k padded rows, odd widths
    (2, 64, 128, 64, 513, (3, 5), True, 1, 0, 1.0, 1.0),
    (1, 128, 256, 64, 257, (3, 3), True, 1, 0, 1.0, 1.0),
    (1, 256, 512, 33, 129, (3, 3), True, 1, 0, 1.0, 1.0),
]


@pytest.mark.parametrize("case", F16_CASES)
def test_conv_f16_split_fwd_dgrad_wgrad(case):
    """tcgen05 3xF16 convolution (packed fp16 hi|lo operands with device-side power-of-two scales) against F.conv2d
    in fp32, at operand magnitudes from 1e-7 to 2e4.  Tolerance 1e-5 of the largest output, as for 3xTF32: the
    split keeps 22 bits per operand and the four-accumulator scheme bounds the truncating accumulation."""
    L = _lib()
    n, cin, cout, h, w, (kh, kw), bias, act, extra, xmag, dymag = case
    g = torch.Generator().manual_seed(cin * 11 + cout)
    x = torch.randn(n, cin, h, w, generator=g) * xmag
    x[0, 0, 0, 0] = 0.0
    x[0, 1 % cin, 0, 1] = xmag * 1e-9          # far below the fp16 normal range of the scaled tensor
    wt = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = (torch.randn(cout, generator=g) * xmag) if bias else None
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    ref = F.conv2d(x.double(), wt.double(), b.double() if bias else None, 1, (ph, pw))
    if act:
        ref = F.relu(ref)
    tph, tpw = ph + extra, pw + extra
    x_h2, x_b, x_src = pack_f16(L, x, tph, tpw)
    assert abs(x_b.item() - x.abs().max().item()) <= 1e-6 * x.abs().max().item()
    # the packed planes reproduce the fp32 tensor to 2^-22 of each value (or 2^-36 / scale absolutely)
    s = 2.0 ** (14 - torch.tensor(x_b.item()).frexp().exponent.item())
    back = (x_h2[..., 0, :].double() + x_h2[..., 1, :].double() / 2048.0) / s
    assert ((back - x_src.double()).abs() <= x_src.double().abs() * 2.0 ** -21 + 2.0 ** -35 / s).all()
    wd = wt.to(DEV)
    w_b = torch.empty(1, device=DEV)
    w_h2 = torch.empty(cout, 2, kh * kw * cin, dtype=torch.float16, device=DEV)
    L.weight_pack_f16(wd.data_ptr(), cout, cin, kh, kw, cin, 0, 1, w_b.data_ptr(), w_h2.data_ptr(), _st())
    y = torch.full((n, h, w, cout), float("nan"), device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    xt4, yt4 = L.Tensor4(n, h, w, cin, tph, tpw), L.Tensor4(n, h, w, cout, 0, 0)
    cv = L.Conv(kh, kw, 1, 1, ph, pw)
    L.profile_enable(1)
    L.conv2d_fwd_f16(xt4, x_h2.data_ptr(), x_b.data_ptr(), w_h2.data_ptr(), w_b.data_ptr(),
                     b.to(DEV).data_ptr() if bias else None, cv, act, yt4, y.data_ptr(), stats.data_ptr(), _st())
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    assert "conv_fwd_tc" in prof and len(prof) == 1, prof
    assert relerr(from_nhwc(y).double(), ref) < 1e-5
    st_ = stats.cpu()
    assert torch.allclose(st_[:cout], ref.sum((0, 2, 3)), rtol=5e-5, atol=1e-4 * xmag)
    assert torch.allclose(st_[cout:], (ref ** 2).sum((0, 2, 3)), rtol=5e-5, atol=1e-4 * xmag * xmag)

    # dgrad: convolution of the padded dy with the flipped / transposed weights
    dy = torch.randn(n, cout, h, w, generator=g) * dymag
    xr = x.double().requires_grad_(True)
    wr = wt.double().requires_grad_(True)
    F.conv2d(xr, wr, None, 1, (ph, pw)).backward(dy.double())
    dy_h2, dy_b, _ = pack_f16(L, dy, tph, tpw)
    wt_h2 = torch.empty(cin, 2, kh * kw * cout, dtype=torch.float16, device=DEV)
    L.weight_pack_f16(wd.data_ptr(), cout, cin, kh, kw, cin, 1, 0, w_b.data_ptr(), wt_h2.data_ptr(), _st())
    dx = torch.full((n, h, w, cin), float("nan"), device=DEV)
    dyt4 = L.Tensor4(n, h, w, cout, tph, tpw)
    L.conv2d_bwd_data_f16(dyt4, dy_h2.data_ptr(), dy_b.data_ptr(), wt_h2.data_ptr(), w_b.data_ptr(), cv,
                          L.Tensor4(n, h, w, cin, 0, 0), dx.data_ptr(), 0, _st())
    assert relerr(from_nhwc(dx).double(), xr.grad) < 1e-5
    # accumulate: dx += the gradient (what a tensor with two consumers gets from its second one)
    L.conv2d_bwd_data_f16(dyt4, dy_h2.data_ptr(), dy_b.data_ptr(), wt_h2.data_ptr(), w_b.data_ptr(), cv,
                          L.Tensor4(n, h, w, cin, 0, 0), dx.data_ptr(), 1, _st())
    assert relerr(from_nhwc(dx).double(), 2 * xr.grad) < 1e-5

    # wgrad: x and dy on the same padded grid, both read pixel-major (MN-major fp16 operands, 128-byte swizzle)
    if cout % 64 == 0:      # Cout % 128 == 64: the upper half of the last 128-channel tile is TMA zero fill
        dw_ohwi = torch.full((cout, kh, kw, cin), float("nan"), device=DEV)
        L.conv2d_bwd_weight_f16(xt4, x_h2.data_ptr(), x_b.data_ptr(), dyt4, dy_h2.data_ptr(), dy_b.data_ptr(), cv,
                                dw_ohwi.data_ptr(), _st())
        assert relerr(dw_ohwi.permute(0, 3, 1, 2).cpu().double(), wr.grad) < 1e-5
    else:
        with pytest.raises(L.DlioError):
            L.conv2d_bwd_weight_f16(xt4, x_h2.data_ptr(), x_b.data_ptr(), dyt4, dy_h2.data_ptr(), dy_b.data_ptr(), cv,
                                    y.data_ptr(), _st())


TC_CASES = [
    # n, cin, cout, h, w, (kh, kw), bias, act, extra tensor pad
    (2, 64, 128, 9, 33, (3, 5), True, 1, 0),
    (1, 128, 128, 8, 17, (3, 3), False, 0, 0),
    (2, 32, 64, 7, 19, (1, 1), True, 0, 1),
    (1, 256, 192, 5, 9, (3, 3), True, 1, 0),
    (2, 128, 256, 6, 11, (3, 3), False, 0, 0),
    (3, 512, 512, 17, 65, (3, 3), True, 1, 0),
    (1, 64, 48, 4, 6, (3, 3), False, 0, 0),
    # PointSeg Fire shapes: squeeze 1x1 to a narrow tensor, expand 1x1 / 3x3 from it (dgrad on 80 / 48 channels)
    (2, 768, 80, 6, 10, (1, 1), True, 0, 1),
    (2, 80, 384, 6, 10, (1, 1), True, 0, 1),
    (2, 80, 384, 6, 10, (3, 3), True, 0, 0),
    (2, 48, 192, 7, 9, (3, 3), True, 0, 0),
    (4, 64, 256, 8, 16, (3, 3), True, 0, 0),
    # Cin = 32 with Cout % 128 == 0 (the first layer's space-to-depth view): wgrad with swapped operands, four taps
    # per CTA -- 15 taps leave one dummy tap in the last group, 9 taps three
    (2, 32, 128, 9, 21, (5, 3), True, 0, 0),
    (3, 32, 256, 6, 14, (3, 3), False, 0, 1),
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_fwd_and_dgrad(case):
    """tcgen05 3xTF32 implicit-GEMM convolution against F.conv2d in fp32.  Tolerance 1e-5 relative to the largest
    output: the split drops the lo*lo term (~2^-20 per product) and the tensor core's truncating accumulation
    adds ~1e-9 * K (measured 3.6e-6 at K = 4608, the same level as the fp32 CUDA-core kernel)."""
    L = _lib()
    n, cin, cout, h, w, (kh, kw), bias, act, extra = case
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g) if bias else None
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    ref = F.conv2d(x, wt, b, 1, (ph, pw))
    if act:
        ref = F.relu(ref)
    tph, tpw = ph + extra, pw + extra
    x_hi, x_lo = split_padded(L, x, tph, tpw)
    assert torch.equal(x_hi[:, tph:tph + h, tpw:tpw + w].cpu(), x.permute(0, 2, 3, 1))      # hi plane = full fp32
    wd = wt.to(DEV)
    w_hi, w_lo = torch.empty(cout, kh, kw, cin, device=DEV), torch.empty(cout, kh, kw, cin, device=DEV)
    L.weight_to_ohwi(wd.data_ptr(), cout, cin, kh, kw, cin, w_hi.data_ptr(), w_lo.data_ptr(), _st())
    y = torch.full((n, h, w, cout), float("nan"), device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    xt4, yt4 = L.Tensor4(n, h, w, cin, tph, tpw), L.Tensor4(n, h, w, cout, 0, 0)
    cv = L.Conv(kh, kw, 1, 1, ph, pw)
    L.profile_enable(1)
    L.conv2d_fwd(xt4, x_hi.data_ptr(), x_lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(),
                 b.to(DEV).data_ptr() if bias else None, cv, act, yt4, y.data_ptr(), stats.data_ptr(), _st())
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    fwd_tc = cin % 32 == 0 and cout % 16 == 0
    assert ("conv_fwd_tc" if fwd_tc else "conv_fwd_simt") in prof and len(prof) == 1, prof   # which kernel really ran
    got = from_nhwc(y)
    assert relerr(got, ref) < 1e-5
    s = stats.cpu()   # the truncating accumulation biases every output the same way, so sums keep ~1e-5 of it
    assert torch.allclose(s[:cout], ref.double().sum((0, 2, 3)), rtol=5e-5, atol=1e-4)
    assert torch.allclose(s[cout:], (ref.double() ** 2).sum((0, 2, 3)), rtol=5e-5, atol=1e-4)

    # dgrad: convolution of the padded dy with the flipped / transposed weights
    dy = torch.randn(n, cout, h, w, generator=g)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, wt, None, 1, (ph, pw)).backward(dy)
    dy_hi, dy_lo = split_padded(L, dy, kh - 1 - ph, kw - 1 - pw)
    wt_hi, wt_lo = torch.empty(cin, kh, kw, cout, device=DEV), torch.empty(cin, kh, kw, cout, device=DEV)
    L.weight_flip_transpose(w_hi.data_ptr(), cout, cin, kh, kw, wt_hi.data_ptr(), wt_lo.data_ptr(), _st())
    dx = torch.full((n, h, w, cin), float("nan"), device=DEV)
    L.profile_enable(1)
    L.conv2d_bwd_data(L.Tensor4(n, h, w, cout, kh - 1 - ph, kw - 1 - pw), dy_hi.data_ptr(), dy_lo.data_ptr(),
                      w_hi.data_ptr(), w_lo.data_ptr(), wt_hi.data_ptr(), wt_lo.data_ptr(), cv,
                      L.Tensor4(n, h, w, cin, 0, 0), dx.data_ptr(), 0, _st())
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    assert ("conv_dgrad_tc" if (cout % 32 == 0 and cin % 16 == 0) else "conv_dgrad_simt") in prof, prof
    assert relerr(from_nhwc(dx), xr.grad) < 1e-5

    # wgrad: dy on the same padded grid as x, both operands read pixel-major (MN-major tcgen05 operands)
    wr = wt.clone().requires_grad_(True)
    F.conv2d(x, wr, None, 1, (ph, pw)).backward(dy)
    dyp_hi, dyp_lo = split_padded(L, dy, tph, tpw)
    dw_ohwi = torch.full((cout, kh, kw, cin), float("nan"), device=DEV)
    L.profile_enable(1)
    L.conv2d_bwd_weight(xt4, x_hi.data_ptr(), x_lo.data_ptr(), L.Tensor4(n, h, w, cout, tph, tpw), dyp_hi.data_ptr(),
                        dyp_lo.data_ptr(), cv, dw_ohwi.data_ptr(), _st())
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    assert ("conv_wgrad_tc" if (cout % 32 == 0 and cin % 32 == 0) else "conv_wgrad_simt") in prof, prof
    assert relerr(dw_ohwi.permute(0, 3, 1, 2).cpu(), wr.grad) < 1e-5


PAIR_CASES = [
    # n, cin, cout, h, w, (kh, kw), (sh, sw), bias
    (2, 64, 128, 8, 32, (3, 5), (1, 2), False),      # FlowNet conv2 / conv3
    (2, 128, 192, 10, 12, (3, 3), (2, 2), False),    # FlowNet conv4..6, ResNet layer3/4 first blocks
    (3, 64, 64, 6, 20, (3, 3), (1, 2), True),        # ResNet layer1.0.conv1 (Cout = 64: half-filled wgrad tile)
    (2, 64, 128, 7, 16, (1, 1), (2, 2), False),      # ResNet 1x1 downsample, odd height
    (1, 32, 64, 4, 8, (1, 1), (1, 2), False),
    # real layer shapes: FlowNet conv4 (3x3 stride (2,2) on 256 x 64 x 256: row decimation in the epilogue, dy spread
    # over the even rows) and conv2 (3x5 stride (1,2) on 64 x 64 x 1024)
    (1, 256, 512, 64, 256, (3, 3), (2, 2), False),
    (1, 64, 128, 64, 1024, (3, 5), (1, 2), False),
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_pair_view_strided_fwd_dgrad_wgrad(case):
    """W-stride-2 (and (2,2)) convolutions on the tensor cores through the pixel-pair view: input planes in the
    pixel-pair layout (written by dlio_bn_act_pool_fwd with out_group = 2), pair-view weights, H stride by row
    decimation in the epilogue; backward as stride-1 dgrad / wgrad with dy spread over the even rows
    (dlio_bn_bwd_apply).  Against F.conv2d in fp64, 1e-5 of the largest entry as for the stride-1 path."""
    L = _lib()
    n, cin, cout, h, w, (kh, kw), (sh, sw), bias = case
    g = torch.Generator().manual_seed(cin * 13 + cout + kw)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g) if bias else None
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xr, wr = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    ref = F.conv2d(xr, wr, b.double() if bias else None, (sh, sw), (ph, pw))
    ho, wo = ref.shape[2], ref.shape[3]
    assert wo == w // 2 and ho == ((h - 1) // 2 + 1 if sh == 2 else h)
    kw2 = 3 if kw > 1 else 1
    pw2 = (kw2 - 1) // 2
    tph, tpw = ph, 2                                       # even row pads
    hp, wp = h + 2 * tph, w + 2 * tpw
    # x planes: identity pass of the fused BN / pool kernel into a padded tensor, pixel-pair layout
    xs = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    x_b = torch.tensor([x.abs().max().item() * 1.01], device=DEV)
    x_h2 = torch.full((n, hp, wp, 2, cin), float("nan"), dtype=torch.float16, device=DEV)
    x_f32 = torch.empty(n, hp, wp, cin, device=DEV)
    t_in, t_out = L.Tensor4(n, h, w, cin, 0, 0), L.Tensor4(n, h, w, cin, tph, tpw)
    L.bn_act_pool_fwd(t_in, xs.data_ptr(), None, None, t_in, None, L.BnPool(0, 0, 1, 1, 1, 0, 2), t_out,
                      x_f32.data_ptr(), None, x_h2.data_ptr(), x_b.data_ptr(), None, None, _st())
    # the same planes in the plain layout, regrouped on the host: [n,hp,wp/2,pixel,plane,c] -> [.., plane, pixel, c]
    plain = torch.full((n, hp, wp, 2, cin), float("nan"), dtype=torch.float16, device=DEV)
    L.bn_act_pool_fwd(t_in, xs.data_ptr(), None, None, t_in, None, L.BnPool(0, 0, 1, 1, 1, 0, 1), t_out,
                      x_f32.data_ptr(), None, plain.data_ptr(), x_b.data_ptr(), None, None, _st())
    regrouped = plain.view(n, hp, wp // 2, 2, 2, cin).permute(0, 1, 2, 4, 3, 5).contiguous()
    assert torch.equal(regrouped.view(-1).view(torch.int16), x_h2.view(-1).view(torch.int16))

    wd = wt.to(DEV)
    w_b = torch.empty(1, device=DEV)
    w_h2 = torch.empty(cout, 2, kh * kw2 * 2 * cin, dtype=torch.float16, device=DEV)
    L.weight_pack_pair_f16(wd.data_ptr(), cout, cin, kh, kw, 0, 1, w_b.data_ptr(), w_h2.data_ptr(), _st())
    y = torch.full((n, ho, wo, cout), float("nan"), device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    x2t4 = L.Tensor4(n, h, w // 2, 2 * cin, tph, tpw // 2)
    L.profile_enable(1)
    L.conv2d_fwd_f16(x2t4, x_h2.data_ptr(), x_b.data_ptr(), w_h2.data_ptr(), w_b.data_ptr(),
                     b.to(DEV).data_ptr() if bias else None, L.Conv(kh, kw2, sh, 1, ph, pw2), 0,
                     L.Tensor4(n, ho, wo, cout, 0, 0), y.data_ptr(), stats.data_ptr(), _st())
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    assert "conv_fwd_tc" in prof and len(prof) == 1, prof
    refd = ref.detach()
    assert relerr(from_nhwc(y).double(), refd) < 1e-5
    st_ = stats.cpu()
    assert torch.allclose(st_[:cout], refd.sum((0, 2, 3)), rtol=5e-5, atol=1e-4)
    assert torch.allclose(st_[cout:], (refd ** 2).sum((0, 2, 3)), rtol=5e-5, atol=1e-4)

    # backward: dy spread over the even rows of the pair-view grid by dlio_bn_bwd_apply (eval-mode BN with unit scale
    # == copy), then stride-1 dgrad / wgrad
    dy = torch.randn(n, cout, ho, wo, generator=g)
    ref.backward(dy.double())
    dz = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    sums = torch.zeros(2 * cout + 1, dtype=torch.float64, device=DEV)
    sums[2 * cout] = dy.abs().max().item()
    dy_h2 = torch.full((n, hp, w // 2 + 2 * (tpw // 2), 2, cout), float("nan"), dtype=torch.float16, device=DEV)
    dy_b = torch.empty(1, device=DEV)
    dyt4 = L.Tensor4(n, h, wo, cout, tph, tpw // 2)
    L.bn_bwd_apply(L.Tensor4(n, ho, wo, cout, 0, 0), y.data_ptr(), dz.data_ptr(), sums.data_ptr(), n * ho * wo, None,
                   None, None, 0, 0, dyt4, None, None, dy_h2.data_ptr(), dy_b.data_ptr(), None, None, None, None, 0, _st())
    cv1 = L.Conv(kh, kw2, 1, 1, ph, pw2)
    wt_h2 = torch.empty(2 * cin, 2, kh * kw2 * cout, dtype=torch.float16, device=DEV)
    L.weight_pack_pair_f16(wd.data_ptr(), cout, cin, kh, kw, 1, 0, w_b.data_ptr(), wt_h2.data_ptr(), _st())
    dx = torch.full((n, h, w, cin), float("nan"), device=DEV)
    L.conv2d_bwd_data_f16(dyt4, dy_h2.data_ptr(), dy_b.data_ptr(), wt_h2.data_ptr(), w_b.data_ptr(), cv1,
                          L.Tensor4(n, h, w // 2, 2 * cin, 0, 0), dx.data_ptr(), 0, _st())
    assert relerr(from_nhwc(dx).double(), xr.grad) < 1e-5
    dw2 = torch.full((cout, kh, kw2, 2 * cin), float("nan"), device=DEV)
    L.conv2d_bwd_weight_f16(x2t4, x_h2.data_ptr(), x_b.data_ptr(), dyt4, dy_h2.data_ptr(), dy_b.data_ptr(), cv1,
                            dw2.data_ptr(), _st())
    dw = torch.full((cout, cin, kh, kw), float("nan"), device=DEV)
    L.weight_grad_from_pair(dw2.data_ptr(), cout, cin, kh, kw, dw.data_ptr(), _st())
    assert relerr(dw.cpu().double(), wr.grad) < 1e-5


@pytest.mark.parametrize("cfg", [
    # c, h, w, pre_relu, relu, pool, ceil, res_mode
    (64, 9, 20, True, False, (1, 2), True, 0),       # Simple-1 blocks: BN directly followed by the pool -> backward
    (128, 10, 33, True, False, (2, 2), True, 0),     # sums from the pooled side, un-pooling inside the apply pass
    (64, 12, 9, True, False, (2, 2), False, 0),
    (64, 9, 20, True, False, (1, 2), True, 0, "two-pass"),   # the same block through dz (reduce + apply passes)
    (128, 10, 33, False, True, (2, 2), True, 0),
    (48, 8, 16, False, True, (1, 2), False, 0),
    (64, 7, 11, False, True, None, False, 1),
    (32, 6, 9, False, True, None, False, 2),
    (16, 5, 8, False, False, None, False, 0),
])
def test_conv_bn_block_vs_torch(cfg):
    """engine.conv_bn (conv + BN(train) + residual + ReLU + max-pool, forward and backward) against the
    same block written with torch.nn.functional on the CPU.  Tolerance 2e-5 relative (BN divides by a
    batch std computed in a different summation order)."""
    from deeplio_b200 import engine as E
    c, h, w, pre_relu, relu, pool, ceil, res_mode = cfg[:8]
    E.FUSED_POOL_BWD = len(cfg) == 8
    n, cin = 2, 32
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = (torch.randn(c, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(c, generator=g) * 0.1
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    res = torch.randn(n, c, h, w, generator=g) if res_mode else None

    def ref_block(x, wt, b, gamma, beta, res, rm, rv):
        y = F.conv2d(x, wt, b, 1, 1)
        if pre_relu:
            y = F.relu(y)
        y = F.batch_norm(y, rm, rv, gamma, beta, True, 0.1, 1e-5)
        if res_mode == 1:
            y = y + res
        if relu:
            y = F.relu(y)
        if res_mode == 2:
            y = y + res
        if pool:
            y = F.max_pool2d(y, 3, pool, 1, ceil_mode=ceil)
        return y
    leaves = [t.clone().requires_grad_(True) for t in (x, wt, b, gamma, beta)] + ([res.clone().requires_grad_(True)] if res_mode else [None])
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = ref_block(*leaves, rm_ref, rv_ref)
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout)

    params = {"cv.weight": wt.to(DEV), "cv.bias": b.to(DEV), "bn.weight": gamma.to(DEV), "bn.bias": beta.to(DEV)}
    bufs = {"bn.running_mean": rm.to(DEV), "bn.running_var": rv.to(DEV),
            "bn.num_batches_tracked": torch.zeros((), dtype=torch.long, device=DEV)}
    run = E.Run(params, bufs, torch.device(DEV), True, True)
    xa = E.Act(n, h, w, cin, 1, 1, t=to_padded_nhwc(x, 1, 1))
    ra = E.Act(n, h, w, c, 0, 0, t=to_padded_nhwc(res, 0, 0)) if res_mode else None
    try:
        out = E.conv_bn(run, xa, "cv", "bn", (1, 1), pre_relu=pre_relu, relu=relu, pool=pool, ceil=ceil, res=ra,
                        res_mode=res_mode, out_pad=(1, 2))
    finally:
        E.FUSED_POOL_BWD = True
    got = from_nhwc(out.t[:, 1:1 + out.h, 2:2 + out.w])
    assert got.shape == ref.shape
    assert relerr(got, ref.detach()) < 2e-5
    assert out.t[:, 0].abs().max().item() == 0 and out.t[:, :, :2].abs().max().item() == 0   # zero pads
    assert torch.allclose(bufs["bn.running_mean"].cpu(), rm_ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bufs["bn.running_var"].cpu(), rv_ref, rtol=1e-5, atol=1e-6)
    assert int(bufs["bn.num_batches_tracked"]) == 1
    run.agrad[id(out)] = to_padded_nhwc(dout, 0, 0)
    run.backward()
    tol = 5e-5
    assert relerr(from_nhwc(run.agrad[id(xa)]), leaves[0].grad) < tol
    assert relerr(run.pgrad["cv.weight"].cpu(), leaves[1].grad) < tol
    assert (run.pgrad["cv.bias"].cpu() - leaves[2].grad).abs().max() < tol * leaves[1].grad.abs().max() * 10
    assert relerr(run.pgrad["bn.weight"].cpu(), leaves[3].grad) < tol
    assert relerr(run.pgrad["bn.bias"].cpu(), leaves[4].grad) < tol
    if res_mode:
        assert relerr(from_nhwc(run.agrad[id(ra)]), leaves[5].grad) < tol


@pytest.mark.parametrize("k,stride,pre_relu,pool,first_ph,folded", [
    ((5, 7), (1, 2), True, (1, 2), 2, False),   # Simple-1 conv1 (+ ceil-mode pool); FlowNet conv1 has the same convolution
    ((3, 5), (1, 2), False, (1, 2), 1, False),  # PointSeg conv1a
    ((5, 7), (1, 1), False, (1, 2), 2, False),  # ResNet conv1 (stride 1: four output pixels per space-to-depth group)
    ((5, 7), (1, 2), True, (1, 2), 2, True),    # the same layers on packed fp16 input planes ("folded split"
    ((3, 5), (1, 2), False, (1, 2), 1, True),   # operands): what the nets run when they are fed PairedFrames
    ((5, 7), (1, 1), False, (1, 2), 2, True),   # ResNet conv1 folded: four output pixels per group, N = 256
])
def test_first_layer_space_to_depth_at_64x2048(k, stride, pre_relu, pool, first_ph, folded):
    """The first convolution of every encoder at its real size (two 6-channel 64 x 2048 images): dlio_pack_input from
    the strided [N, T, C, H, W] view, the 3xTF32 tcgen05 kernel over the space-to-depth view (csrc/conv_s2d.cu), BN,
    pool; backward through the swapped-operand wgrad.  ``folded``: dlio_pair_gather writes packed fp16 planes and the
    layer runs on the fp16 kernels with the folded-split operands.  Against the same block in torch fp64."""
    from deeplio_b200 import engine as E
    n, h, w, cout = 2, 64, 2048, 64
    kh, kw = k
    g = torch.Generator().manual_seed(kh * 10 + kw + stride[1])
    pairs = torch.randn(n, 2, 6, h, w, generator=g) * torch.tensor([0.13, 0.1, 0.01, 0.34, 0.44, 0.57]).view(1, 1, 6, 1, 1)
    wt = torch.randn(cout, 6, kh, kw, generator=g) / (6 * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    x = pairs[:, :, 0:3].reshape(n, 6, h, w)
    dout_shape = F.max_pool2d(torch.zeros(1, 1, h, (w + 2 * ((kw - 1) // 2) - kw) // stride[1] + 1), 3, pool, 1,
                              ceil_mode=pre_relu).shape[2:]
    dout = torch.randn((n, cout) + tuple(dout_shape), generator=g)

    def reference(mask, argmax):
        """The block in fp64.  ``mask`` / ``argmax``: the ReLU and max-pool arg-max decisions the B200 path took
        (engine.MASK_TRACE).  A weight gradient here is a sum of 1.3e5 terms of random sign per entry, so ONE ReLU
        input within round-off of zero (or one pair of pooling candidates within round-off of a tie) decided the other
        way moves it by ~1 / sqrt(1.3e5) = 3e-3 of its size -- far above the 5e-5 bar; with the decisions imposed, the
        comparison is about arithmetic only (the number of differing decisions is bounded separately)."""
        from oracle import deeplio_oracle as O
        leaves = [t.double().requires_grad_(True) for t in (wt, b, gamma, beta)]
        y = F.conv2d(x.double(), leaves[0], leaves[1], stride, ((kh - 1) // 2, (kw - 1) // 2))
        flips = 0
        if pre_relu:
            flips = int(((y > 0) != mask).sum())
            y = y * mask
        y = F.batch_norm(y, None, None, leaves[2], leaves[3], True, 0.1, 1e-5)
        if not pre_relu:
            flips = int(((y > 0) != mask).sum())
            y = y * mask
        O.TRACE, O.FORCE_MASKS = {}, {"pool": argmax}
        try:
            ref = O._pool(y, pool, pre_relu, "pool")
            flips += int((O.TRACE["pool"] != argmax).sum())
        finally:
            O.TRACE = O.FORCE_MASKS = None
        ref.backward(dout.double())
        return ref, leaves, flips

    params = {"cv.weight": wt.to(DEV), "cv.bias": b.to(DEV), "bn.weight": gamma.to(DEV), "bn.bias": beta.to(DEV)}
    bufs = {"bn.running_mean": torch.zeros(cout, device=DEV), "bn.running_var": torch.ones(cout, device=DEV)}
    run = E.Run(params, bufs, torch.device(DEV), True, True)
    L = _lib()
    if folded:
        from deeplio_b200 import data
        frames = pairs.to(DEV)                            # [B, F = 2, 6, H, W], one pair (0, 1) per sample
        h2, bound = data.pair_gather(torch.device(DEV), data.PairedFrames(frames, [[0, 1]], 0, 3), 8, first_ph, 4,
                                     False, f16=True)
        assert abs(bound.item() - pairs.abs().max().item()) <= 1e-6 * pairs.abs().max().item()
        x0 = E.Act(n, h, w, 8, first_ph, 4, needs_grad=False, f32=False)
        x0.h2, x0.bound = h2, bound
    else:
        view = pairs.to(DEV)[:, :, 0:3]                   # the non-contiguous channel slice the trainer hands over
        x0 = E.pack_input(run, view, 8, first_ph, 4)
    L.profile_enable(1)
    E.MASK_TRACE = {}
    try:
        out = E.conv_bn(run, x0, "cv", "bn", stride, pre_relu=pre_relu, relu=not pre_relu, pool=pool, ceil=pre_relu,
                        out_pad=(1, 2))
        mask, argmax = E.MASK_TRACE["cv"], E.MASK_TRACE["cv#pool"]
    finally:
        E.MASK_TRACE = None
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    assert "conv_fwd_tc" in prof and "conv_fwd_simt" not in prof, prof
    ref, leaves, flips = reference(mask, argmax)
    assert flips <= 1e-5 * mask.numel(), (flips, mask.numel())
    got = from_nhwc(out.t[:, 1:1 + out.h, 2:2 + out.w]).double()
    assert got.shape == ref.shape
    assert relerr(got, ref.detach()) < 2e-5
    run.agrad[id(out)] = to_padded_nhwc(dout, 0, 0)
    L.profile_enable(1)
    run.backward()
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    assert "conv_wgrad_tc" in prof and "conv_wgrad_simt" not in prof, prof
    tol = 5e-5
    assert relerr(run.pgrad["cv.weight"].cpu().double(), leaves[0].grad) < tol
    assert (run.pgrad["cv.bias"].cpu().double() - leaves[1].grad).abs().max() < tol * leaves[0].grad.abs().max() * 10
    assert relerr(run.pgrad["bn.weight"].cpu().double(), leaves[2].grad) < tol
    assert relerr(run.pgrad["bn.bias"].cpu().double(), leaves[3].grad) < tol


@pytest.mark.parametrize("variant", ["simple1", "bn_relu", "eval"])
def test_conv_bn_chain_on_fp16_planes(variant):
    """Two chained engine.conv_bn blocks where the second convolution (forward, dgrad, wgrad) runs on the packed
    fp16 planes the first block's BN pass wrote (no fp32 copy in the 'simple1' variant), against the same chain in
    torch fp64 on the CPU.  Checks the device-side bounds: they must dominate the tensors they scale."""
    from deeplio_b200 import engine as E
    n, cin, c1, c2, h, w = 2, 32, 64, 128, 12, 36
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, cin, h, w, generator=g)
    w1 = torch.randn(c1, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    w2 = torch.randn(c2, c1, 3, 3, generator=g) / (c1 * 9) ** 0.5
    b1, b2 = torch.randn(c1, generator=g) * 0.1, torch.randn(c2, generator=g) * 0.1
    g1, be1 = torch.rand(c1, generator=g) * 3 + 0.5, torch.randn(c1, generator=g)
    g2, be2 = torch.rand(c2, generator=g) + 0.5, torch.randn(c2, generator=g) * 0.1
    pre_relu = variant == "simple1"
    training = variant != "eval"
    rm1, rv1 = torch.randn(c1, generator=g) * 0.1, torch.rand(c1, generator=g) + 0.5
    rm2, rv2 = torch.randn(c2, generator=g) * 0.1, torch.rand(c2, generator=g) + 0.5

    def block(t, wt, b, gamma, beta, rm, rv, pool, res=None):
        y = F.conv2d(t, wt, b, 1, 1)
        if pre_relu:
            y = F.relu(y)
        y = F.batch_norm(y, rm.clone().double(), rv.clone().double(), gamma, beta, training, 0.1, 1e-5)
        if res is not None:
            y = y + res
        if not pre_relu:
            y = F.relu(y)
        if pool:
            y = F.max_pool2d(y, 3, pool, 1, ceil_mode=True)
        return y
    leaves = [t.double().clone().requires_grad_(True) for t in (x, w1, b1, g1, be1, w2, b2, g2, be2)]
    mid = block(leaves[0], *leaves[1:5], rm1, rv1, (1, 2))
    ref = block(mid, *leaves[5:9], rm2, rv2, None)
    dout = torch.randn(ref.shape, generator=g)
    if training:
        ref.backward(dout.double())

    params = {"c1.weight": w1, "c1.bias": b1, "n1.weight": g1, "n1.bias": be1,
              "c2.weight": w2, "c2.bias": b2, "n2.weight": g2, "n2.bias": be2}
    params = {k: v.to(DEV) for k, v in params.items()}
    bufs = {"n1.running_mean": rm1.to(DEV), "n1.running_var": rv1.to(DEV),
            "n2.running_mean": rm2.to(DEV), "n2.running_var": rv2.to(DEV)}
    run = E.Run(params, bufs, torch.device(DEV), training, training)
    xa = E.Act(n, h, w, cin, 1, 1, t=to_padded_nhwc(x, 1, 1))
    L = _lib()
    L.profile_enable(1)
    m = E.conv_bn(run, xa, "c1", "n1", (1, 1), pre_relu=pre_relu, relu=not pre_relu, pool=(1, 2), ceil=True,
                  out_pad=(1, 1), out_f32=variant != "simple1")
    assert m.h2 is not None and (m.t is None) == (variant == "simple1")
    out = E.conv_bn(run, m, "c2", "n2", (1, 1), pre_relu=pre_relu, relu=not pre_relu, out_pad=(0, 0))
    torch.cuda.synchronize()
    prof = L.profile_read()
    assert prof["conv_fwd_tc"][1] == 1 and prof["conv_fwd_simt"][1] == 1, prof
    # the bound dominates the tensor, and the planes reproduce it
    s = 2.0 ** (14 - torch.tensor(m.bound.item()).frexp().exponent.item())
    back = (m.h2[..., 0, :].double() + m.h2[..., 1, :].double() / 2048.0).cpu() / s
    mid_nhwc = F.pad(mid.detach().permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))
    assert mid.detach().abs().max().item() <= m.bound.item() <= 3000 * mid.detach().abs().max().item()
    assert relerr(back, mid_nhwc) < 2e-5
    assert relerr(from_nhwc(out.t).double(), ref.detach()) < 2e-5
    if not training:
        L.profile_enable(0)
        return
    run.agrad[id(out)] = to_padded_nhwc(dout, 0, 0)
    run.backward()
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(0)
    assert prof["conv_wgrad_tc"][1] == 1 and prof["conv_dgrad_tc"][1] == 2, prof   # dgrad of block 1 runs on 3xTF32
    tol = 5e-5
    assert relerr(from_nhwc(run.agrad[id(xa)]).double(), leaves[0].grad) < tol
    for name, leaf in zip(("c1.weight", "c1.bias", "n1.weight", "n1.bias", "c2.weight", "c2.bias", "n2.weight", "n2.bias"),
                          leaves[1:]):
        scale = leaf.grad.abs().max().item() if "bias" not in name[:2] else 1.0
        assert (run.pgrad[name].cpu().double() - leaf.grad).abs().max().item() < tol * max(scale, leaves[5].grad.abs().max().item()), name


def test_bn_eval_mode_uses_running_stats():
    from deeplio_b200 import engine as E
    n, cin, c, h, w = 2, 16, 32, 6, 10
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(c, cin, 3, 3, generator=g) * 0.1
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    ref = F.relu(F.batch_norm(F.conv2d(x, wt, None, 1, 1), rm.clone(), rv.clone(), gamma, beta, False, 0.1, 1e-5))
    params = {"cv.weight": wt.to(DEV), "bn.weight": gamma.to(DEV), "bn.bias": beta.to(DEV)}
    bufs = {"bn.running_mean": rm.to(DEV), "bn.running_var": rv.to(DEV)}
    run = E.Run(params, bufs, torch.device(DEV), False, False)
    out = E.conv_bn(run, E.Act(n, h, w, cin, 1, 1, t=to_padded_nhwc(x, 1, 1)), "cv", "bn")
    assert relerr(from_nhwc(out.t), ref) < 1e-5
    assert torch.equal(bufs["bn.running_mean"].cpu(), rm)


@pytest.mark.parametrize("m,n,k,act", [(16, 128, 512, "leaky_relu"), (240, 128, 6, "leaky_relu"), (7, 3, 1024, None),
                                       (16, 4096, 256, None), (33, 130, 257, "sigmoid"), (5, 64, 64, "relu")])
def test_linear_fwd_bwd(m, n, k, act):
    from deeplio_b200 import functional as Fn
    g = torch.Generator().manual_seed(m * 131 + n)
    x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** 0.5, torch.randn(n, generator=g)
    acts = {None: lambda t: t, "relu": F.relu, "leaky_relu": lambda t: F.leaky_relu(t, 0.01), "sigmoid": torch.sigmoid}
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = acts[act](F.linear(xr, wr, br))
    dy = torch.randn(m, n, generator=g)
    ref.backward(dy)
    xd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    got = Fn.linear(xd, wd, bd, act)
    got.backward(dy.to(DEV))
    tol = 1e-5   # fp32 dot products of length <= 4096
    assert relerr(got.detach().cpu(), ref.detach()) < tol
    assert relerr(xd.grad.cpu(), xr.grad) < tol
    assert relerr(wd.grad.cpu(), wr.grad) < tol
    assert relerr(bd.grad.cpu(), br.grad) < tol


@pytest.mark.parametrize("kind,B,T,I,H,L_,bi", [("lstm", 3, 7, 6, 128, 2, True), ("gru", 2, 5, 6, 64, 2, True),
                                                  ("lstm", 9, 3, 40, 32, 1, False), ("gru", 11, 2, 24, 48, 2, False),
                                                  ("lstm", 2, 2, 256, 1024, 2, True)])
def test_rnn_matches_torch(kind, B, T, I, H, L_, bi):
    """Two chained calls with carried state (the IMU net's usage) against nn.LSTM / nn.GRU on the CPU.
    Tolerance 2e-5 relative: fp32 recurrences, different summation order."""
    from deeplio_b200 import functional as Fn
    torch.manual_seed(B * 100 + T)
    cls = torch.nn.LSTM if kind == "lstm" else torch.nn.GRU
    ref = cls(I, H, L_, bidirectional=bi, batch_first=True)
    x1, x2 = torch.randn(B, T, I), torch.randn(B, T, I)
    x1r, x2r = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    o1, s = ref(x1r)
    o2, s2 = ref(x2r, s)
    wsum = torch.randn(B, T, (2 if bi else 1) * H)
    loss = (o1 * wsum).sum() + (o2 * wsum).sum() + (s2[0] if kind == "lstm" else s2).sum()
    loss.backward()
    ws = [p.detach().clone().to(DEV).requires_grad_(True) for p in ref._flat_weights]
    x1d, x2d = x1.to(DEV).requires_grad_(True), x2.to(DEV).requires_grad_(True)
    g1, st = Fn.rnn(x1d, None, kind, L_, bi, H, ws)
    g2, st2 = Fn.rnn(x2d, st, kind, L_, bi, H, ws)
    wd = wsum.to(DEV)
    lossd = (g1 * wd).sum() + (g2 * wd).sum() + (st2[0] if kind == "lstm" else st2).sum()
    lossd.backward()
    tol = 2e-5
    assert relerr(g1.detach().cpu(), o1.detach()) < tol
    assert relerr(g2.detach().cpu(), o2.detach()) < tol
    assert relerr(x1d.grad.cpu(), x1r.grad) < 10 * tol
    assert relerr(x2d.grad.cpu(), x2r.grad) < 10 * tol
    for w, p in zip(ws, ref._flat_weights):
        assert relerr(w.grad.cpu(), p.grad) < 10 * tol


def test_adam_matches_torch():
    L = _lib()
    n = 10007
    g = torch.Generator().manual_seed(5)
    p0, grads = torch.randn(n, generator=g), [torch.randn(n, generator=g) for _ in range(3)]
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, weight_decay=1e-4)
    # arenas are 16-byte aligned slices of a padded buffer
    pd, m, v = p0.to(DEV).clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step, gr in enumerate(grads, 1):
        pr.grad = gr.clone()
        opt.step()
        gd = gr.to(DEV)
        L.adam_step(pd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-4, step,
                    1.0, _st())
    assert (pd.cpu() - pr.detach()).abs().max().item() < 1e-6


def test_dropout_mask_statistics_and_repeatability():
    from deeplio_b200 import functional as Fn
    torch.manual_seed(123)
    Fn._drop_counter[0] = 0
    a = Fn.dropout_mask((1 << 20,), 0.25, torch.device(DEV))
    Fn._drop_counter[0] = 0
    b = Fn.dropout_mask((1 << 20,), 0.25, torch.device(DEV))
    assert torch.equal(a, b)
    keep = (a > 0).float().mean().item()
    assert abs(keep - 0.75) < 3e-3
    assert abs(a.max().item() - 1 / 0.75) < 1e-6
    c = Fn.dropout_mask((1 << 20,), 0.25, torch.device(DEV))
    assert not torch.equal(a, c)


def test_cpu_tensors_are_rejected():
    """No CPU fallback: the product path fails loudly instead of computing on the host."""
    from deeplio_b200 import functional as Fn
    with pytest.raises(RuntimeError):
        Fn.linear(torch.randn(2, 4), torch.randn(3, 4))


@pytest.mark.parametrize("c,h,w,n,pool,ceil,pre_relu,group", [
    (64, 64, 1024, 2, (1, 2), True, True, 1),       # Simple-1 conv1 block: output width 513 (odd), 33 column segments
    (128, 64, 513, 2, (1, 2), True, True, 1),       # conv2 block: odd input width, overhanging last window
    (256, 64, 257, 1, (2, 2), True, True, 1),       # conv4 block: both strides, 33 x 129 output
    (512, 33, 129, 1, (2, 2), True, True, 1),       # conv6 block: odd height
    (64, 64, 2048, 1, (1, 2), False, False, 2),     # ResNet conv1 block: BN -> ReLU -> pool, pixel-pair planes, even pads
    (64, 5, 7, 3, (1, 1), False, False, 1),         # tiny: fewer rows than ring slots, stride-1 pool
    (192, 9, 11, 2, (2, 1), True, True, 1),         # 240-thread blocks (cg = 48), stride (2, 1)
])
def test_pooled_passes_bulk_copy_rings_equal_per_thread_loads(c, h, w, n, pool, ceil, pre_relu, group):
    """The pooling passes staged through shared memory by cp.async.bulk row rings (forward: BN + ReLU + max-pool +
    arg-max + y-at-arg-max + fp16 split; backward: un-pooling BN apply) against the per-thread-load kernels they
    replace: every output bit-identical (same arithmetic, same tie-breaking), bias-gradient sums to round-off."""
    from deeplio_b200 import engine as E
    L = _lib()
    g = torch.Generator().manual_seed(c + h + w)
    cin = 32
    x = torch.randn(n, cin, h, w, generator=g)
    x[:, :, :, : w // 3] = x[:, :, :1, : w // 3]           # constant columns: ties inside pooling windows
    wt = torch.randn(c, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    params = {"cv.weight": wt.to(DEV), "cv.bias": (torch.randn(c, generator=g) * 0.1).to(DEV),
              "bn.weight": (torch.rand(c, generator=g) + 0.5).to(DEV), "bn.bias": (torch.randn(c, generator=g) * 0.1).to(DEV)}
    xd = to_padded_nhwc(x, 1, 1)
    results = []
    for tma in (0, 1):
        L.set_option(b"pool_tma", tma)
        try:
            bufs = {"bn.running_mean": torch.zeros(c, device=DEV), "bn.running_var": torch.ones(c, device=DEV)}
            run = E.Run(dict(params), bufs, torch.device(DEV), True, True)
            xa = E.Act(n, h, w, cin, 1, 1, t=xd.clone())
            L.profile_enable(1)
            out = E.conv_bn(run, xa, "cv", "bn", (1, 1), pre_relu=pre_relu, relu=not pre_relu, pool=pool, ceil=ceil,
                            out_pad=(1, 2), out_group=group)
            gd = torch.Generator().manual_seed(7)
            run.agrad[id(out)] = to_padded_nhwc(torch.randn(n, c, out.h, out.w, generator=gd), 0, 0)
            run.backward()
            torch.cuda.synchronize()
            L.profile_enable(0)
            results.append(dict(t=out.t.clone() if out.t is not None else None, h2=out.h2.clone(), bound=out.bound.clone(),
                                dx=run.agrad[id(xa)].clone(), dw=run.pgrad["cv.weight"].clone(), db=run.pgrad["cv.bias"].clone(),
                                dg=run.pgrad["bn.weight"].clone(), dbeta=run.pgrad["bn.bias"].clone()))
        finally:
            L.set_option(b"pool_tma", 1)
    a, b = results
    assert torch.equal(a["h2"].view(torch.int16), b["h2"].view(torch.int16)) and torch.equal(a["bound"], b["bound"])
    if a["t"] is not None:
        assert torch.equal(a["t"], b["t"])
    for k in ("dg", "dbeta", "dx", "dw"):            # downstream of fp32 / fp64 atomics whose order differs
        assert relerr(a[k], b[k]) < 2e-5, k
    # the bias of a convolution in front of a train-mode BatchNorm has zero gradient in exact arithmetic: noise only
    assert (a["db"] - b["db"]).abs().max().item() < 1e-5 * a["dw"].abs().max().item()


@pytest.mark.parametrize("n,h,w,c,out_c,c_off,pad,relu,res_mode,planes,zero_tail", [
    (2, 9, 33, 64, 64, 0, (1, 1), 1, 0, "h2", 0),        # a plain conv -> BN -> ReLU layer feeding an fp16 convolution
    (3, 7, 21, 16, 64, 0, (1, 1), 1, 0, "h2+t", 1),      # Fire squeeze: 16 channels into a 64-channel tensor, zero tail
    (2, 5, 18, 64, 128, 64, (1, 1), 1, 2, "t+lo", 0),    # Fire expand3x3: channel offset, bypass added after the ReLU
    (2, 6, 20, 128, 128, 0, (1, 2), 1, 1, "h2+t", 0),    # BasicBlock: residual added before the ReLU
    (1, 64, 513, 128, 128, 0, (1, 1), 0, 0, "h2", 0),    # a real Simple-1 conv3 output (no ReLU after the BN)
    (2, 4, 10, 48, 64, 0, (0, 0), 1, 0, "t", 1),         # no spatial pads, 48 -> 64 channels
])
def test_row_structured_bn_apply_equals_the_generic_kernel(n, h, w, c, out_c, c_off, pad, relu, res_mode, planes,
                                                           zero_tail):
    """dlio_bn_act_pool_fwd without pooling: the row-structured kernel (option "apply_rows", the default) writes the
    same bits as the generic per-pixel kernel -- fp32 plane, TF32 lo plane, packed fp16 planes, spatial pads, zero
    tail -- and both equal scale * y + shift (+ residual, ReLU) in torch."""
    L = _lib()
    g = torch.Generator().manual_seed(n * 100 + c + w)
    y = torch.randn(n, c, h, w, generator=g)
    scale, shift = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    res = torch.randn(n, out_c, h, w, generator=g) if res_mode else None
    ref = y * scale.view(1, c, 1, 1) + shift.view(1, c, 1, 1)
    if res_mode == 1:
        ref = ref + res[:, c_off:c_off + c]
    if relu:
        ref = F.relu(ref)
    if res_mode == 2:
        ref = ref + res[:, c_off:c_off + c]
    yd, sd, fd = to_padded_nhwc(y, 0, 0), scale.to(DEV), shift.to(DEV)
    rd = to_padded_nhwc(res, 0, 0) if res_mode else None
    bound = torch.tensor([ref.abs().max().item() * 1.01], device=DEV)
    ph, pw = pad
    outs = []
    for rows in (0, 1):
        L.set_option(b"apply_rows", rows)
        try:
            shape = (n, h + 2 * ph, w + 2 * pw, out_c)
            t = torch.full(shape, float("nan"), device=DEV) if "t" in planes else None
            lo = torch.full(shape, float("nan"), device=DEV) if "lo" in planes else None
            h2 = torch.full(shape[:3] + (2, out_c), float("nan"), dtype=torch.float16, device=DEV) if "h2" in planes else None
            bp = L.BnPool(relu, res_mode, 1, 1, 1, c_off, 1, zero_tail)
            L.bn_act_pool_fwd(L.Tensor4(n, h, w, c, 0, 0), yd.data_ptr(), sd.data_ptr(), fd.data_ptr(),
                              L.Tensor4(n, h, w, out_c, 0, 0), rd.data_ptr() if rd is not None else None, bp,
                              L.Tensor4(n, h, w, out_c, ph, pw), t.data_ptr() if t is not None else None,
                              lo.data_ptr() if lo is not None else None, h2.data_ptr() if h2 is not None else None,
                              bound.data_ptr() if h2 is not None else None, None, None, _st())
            torch.cuda.synchronize()
        finally:
            L.set_option(b"apply_rows", 1)
        outs.append((t, lo, h2))
    written = slice(c_off, out_c if zero_tail else c_off + c)
    for a, b in zip(*outs):
        if a is not None:
            av, bv = (a[..., written], b[..., written]) if a.dim() == 4 else (a[..., written], b[..., written])
            assert torch.equal(av.view(torch.int32 if av.dtype == torch.float32 else torch.int16),
                               bv.view(torch.int32 if bv.dtype == torch.float32 else torch.int16))
    t, lo, h2 = outs[1]
    if t is not None:
        got = from_nhwc(t[:, ph:ph + h, pw:pw + w, c_off:c_off + c])
        assert relerr(got, ref) < 1e-6
        if zero_tail:
            assert (t[..., c:] == 0).all()
        if ph or pw:
            border = t.clone()
            border[:, ph:ph + h, pw:pw + w] = 0
            assert (border[..., written] == 0).all()
    if h2 is not None:
        s = 2.0 ** (14 - torch.tensor(bound.item()).frexp().exponent.item())
        back = (h2[:, ph:ph + h, pw:pw + w, 0, c_off:c_off + c].double() + h2[:, ph:ph + h, pw:pw + w, 1, c_off:c_off + c].double() / 2048.0) / s
        assert relerr(from_nhwc(back), ref.double()) < 1e-6
        if zero_tail:
            assert (h2[..., c:] == 0).all()


@pytest.mark.parametrize("n,h,w,c,dout_c,c_off,relu,res_mode,want_dz", [
    (2, 9, 33, 64, 64, 0, 1, 0, True),         # conv -> BN -> ReLU, dz materialised
    (1, 64, 257, 128, 128, 0, 0, 0, False),    # Simple-1 conv5 at its real size: sums only (SKIP_DZ)
    (3, 7, 21, 64, 128, 64, 1, 2, True),       # Fire expand3x3: channel offset into the concat gradient, bypass gradient
    (2, 5, 10, 16, 16, 0, 1, 0, True),         # Fire squeeze: 16 channels
])
def test_flat_bn_backward_reduce_equals_the_generic_kernel(n, h, w, c, dout_c, c_off, relu, res_mode, want_dz):
    """Backward pass 1 of a layer without pooling (dz = ReLU-masked gradient, sums of dz and dz * yhat, max |dz|,
    bypass gradient): the flat four-pixels-in-flight kernel against the generic per-pixel kernel and torch."""
    L = _lib()
    g = torch.Generator().manual_seed(n + h + w + c)
    y = torch.randn(n, c, h, w, generator=g)
    dout = torch.randn(n, dout_c, h, w, generator=g)
    scale, shift = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    mean, invstd = y.mean((0, 2, 3)), 1.0 / (y.var((0, 2, 3), unbiased=False) + 1e-5).sqrt()
    yd, dd = to_padded_nhwc(y, 0, 0), to_padded_nhwc(dout, 0, 0)
    dev = lambda t: t.to(DEV).contiguous()
    sd, fd, md, isd = dev(scale), dev(shift), dev(mean), dev(invstd)
    dsl = dout[:, c_off:c_off + c]
    v = y * scale.view(1, c, 1, 1) + shift.view(1, c, 1, 1)
    dz_ref = dsl * (v > 0) if relu else dsl
    yhat = (y - mean.view(1, c, 1, 1)) * invstd.view(1, c, 1, 1)
    outs = []
    for rows in (0, 1):
        L.set_option(b"apply_rows", rows)
        try:
            dz = torch.full((n, h, w, c), float("nan"), device=DEV) if want_dz else None
            dres = torch.ones(n, h, w, dout_c, device=DEV) if res_mode == 2 else None
            sums = torch.zeros(2 * c + 1, dtype=torch.float64, device=DEV)
            L.bn_act_pool_bwd_reduce(L.Tensor4(n, h, w, c, 0, 0), yd.data_ptr(), sd.data_ptr(), fd.data_ptr(),
                                     md.data_ptr(), isd.data_ptr(), L.Tensor4(n, h, w, c, 0, 0), None,
                                     L.BnPool(relu, res_mode, 1, 1, 1, c_off), L.GRAD_DIRECT,
                                     L.Tensor4(n, h, w, dout_c, 0, 0), dd.data_ptr(), 0, None,
                                     dz.data_ptr() if dz is not None else None,
                                     dres.data_ptr() if dres is not None else None, dout_c, 1, sums.data_ptr(),
                                     1 if c % 32 == 0 else 0, _st())
            torch.cuda.synchronize()
        finally:
            L.set_option(b"apply_rows", 1)
        outs.append((dz, dres, sums))
    (dz0, dres0, s0), (dz1, dres1, s1) = outs
    if want_dz:
        assert torch.equal(dz0, dz1)
        assert relerr(from_nhwc(dz1), dz_ref) < 1e-6
    if dres1 is not None:
        assert torch.equal(dres0, dres1)
        assert relerr(from_nhwc(dres1)[:, c_off:c_off + c], 1.0 + dsl) < 1e-6     # accumulated into the existing ones
    assert torch.allclose(s0[:2 * c], s1[:2 * c], rtol=1e-5, atol=1e-4)             # fp32 partial sums: order only
    assert torch.allclose(s1[:c].cpu(), dz_ref.double().sum((0, 2, 3)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(s1[c:2 * c].cpu(), (dz_ref * yhat).double().sum((0, 2, 3)), rtol=1e-4, atol=2e-3)
    if c % 32 == 0:
        assert s0[2 * c].item() == s1[2 * c].item() == dz_ref.abs().max().item()
