// C-ABI entry points of the fp16 ("3xF16") tensor-core convolution path and the kernels that build its packed
// split-plane operands from fp32 data (weights per step; activations come out of the fused BN passes in
// norm_pool.cu already packed).  The MMA kernels themselves are the tcgen05 kernels of conv_tc.cu run with
// kind::f16: same shifted-GEMM formulation, same four-accumulator scheme, K = 64 halves per 128-byte chunk.
#include "common.cuh"
#include "conv.cuh"

namespace dlio {

// max |x| -> *bound (caller zeroes).  One atomic per warp.
__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ x, long long n, float *bound) {
    float m = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_nonneg(bound, m);
}

// OIHW fp32 -> packed split rows.  fwd layout: row co, column (tap * cin_pad + ci).  dgrad layout
// (transpose + flip): row ci (cin_pad rows), column (tap' * cout + co) with tap' the flipped tap.
__global__ void __launch_bounds__(256) weight_pack_f16_kernel(const float *__restrict__ w, int cout, int cin, int kh,
                                                              int kw, int cin_pad, int flip, const float *bound,
                                                              __half *__restrict__ out) {
    // four consecutive output columns per thread (cin_pad and cout are multiples of 4): one index decode and two
    // 8-byte stores per four elements -- with one element per thread the kernel was bound by its divisions
    const int taps = kh * kw;
    const long long total4 = (long long)cout * taps * cin_pad / 4;
    const long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const long long i = i4 * 4;
    const float s = f16_scale_from_bound(*bound);
    int row, col;
    long long K;
    __half h[4], l[4];
    if (!flip) {
        const int ci = (int)(i % cin_pad);
        long long t = i / cin_pad;
        const int tap = (int)(t % taps);
        const int co = (int)(t / taps);
        row = co; col = tap * cin_pad + ci; K = (long long)taps * cin_pad;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            f16_split((ci + j < cin ? w[((size_t)co * cin + ci + j) * taps + tap] : 0.f) * s, h[j], l[j]);
    } else {
        const int co = (int)(i % cout);
        long long t = i / cout;
        const int tapf = (int)(t % taps);
        const int ci = (int)(t / taps);
        const int tap = taps - 1 - tapf;        // (kh-1-dy)*kw + (kw-1-dx)
        row = ci; col = tapf * cout + co; K = (long long)taps * cout;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            f16_split((ci < cin ? w[((size_t)(co + j) * cin + ci) * taps + tap] : 0.f) * s, h[j], l[j]);
    }
    __half *o = out + (size_t)row * 2 * K + col;
    *reinterpret_cast<uint2 *>(o) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
    *reinterpret_cast<uint2 *>(o + K) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
}

__global__ void __launch_bounds__(256) pack_f16_kernel(const float *__restrict__ src, long long rows, int c,
                                                       const float *bound, __half *__restrict__ dst) {
    const float s = f16_scale_from_bound(*bound);
    const long long total = rows * (c / 4);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        long long r = i / (c / 4);
        int cc = (int)(i - r * (c / 4)) * 4;
        st4_h2(dst, (size_t)r, c, cc, ld4(src + r * c + cc), s);
    }
}

static int same_conv_check(const dlio_tensor4 &x, const dlio_tensor4 &y, const dlio_conv &cv, const char *who) {
    DLIO_CHECK_ARG(valid_t4(x) && valid_t4(y) && x.n == y.n, "%s: bad tensor descriptor", who);
    DLIO_CHECK_ARG(cv.sh == 1 && cv.sw == 1 && cv.kh == 2 * cv.ph + 1 && cv.kw == 2 * cv.pw + 1 && x.h == y.h && x.w == y.w,
                   "%s: the fp16 tensor-core path takes stride-1 'same' convolutions only", who);
    return DLIO_OK;
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_weight_pack_f16(const float *w_oihw, int cout, int cin, int kh, int kw, int cin_pad,
                                    int transpose_flip, int compute_bound, float *w_bound, void *w_h2, void *stream) {
    DLIO_CHECK_ARG(w_oihw && w_bound && w_h2 && cout > 0 && cin > 0 && cin_pad >= cin && kh > 0 && kw > 0,
                   "weight_pack_f16: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, st);
    const long long n = (long long)cout * cin * kh * kw;
    if (compute_bound) {
        DLIO_CUDA(cudaMemsetAsync(w_bound, 0, sizeof(float), st));
        int grid = ceil_div(n, 256 * 8);
        absmax_kernel<<<grid > 592 ? 592 : grid, 256, 0, st>>>(w_oihw, n, w_bound);
        DLIO_LAUNCH_CHECK();
    }
    DLIO_CHECK_ARG(cin_pad % 4 == 0 && cout % 4 == 0 && (((uintptr_t)w_h2) & 7) == 0,
                   "weight_pack_f16: cin_pad and cout must be multiples of 4");
    const long long total = (long long)cout * kh * kw * cin_pad / 4;
    weight_pack_f16_kernel<<<ceil_div(total, 256), 256, 0, st>>>(w_oihw, cout, cin, kh, kw, cin_pad, transpose_flip,
                                                                w_bound, (__half *)w_h2);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_pack_f16(const float *src, long long rows, int c, float *bound, void *dst_h2, void *stream) {
    DLIO_CHECK_ARG(src && bound && dst_h2 && rows > 0 && c > 0 && c % 4 == 0, "pack_f16: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, st);
    DLIO_CUDA(cudaMemsetAsync(bound, 0, sizeof(float), st));
    const long long n = rows * c;
    int grid = ceil_div(n, 256 * 8);
    absmax_kernel<<<grid > 1184 ? 1184 : grid, 256, 0, st>>>(src, n, bound);
    DLIO_LAUNCH_CHECK();
    grid = ceil_div(n / 4, 256);
    pack_f16_kernel<<<grid > 2368 ? 2368 : grid, 256, 0, st>>>(src, rows, c, bound, (__half *)dst_h2);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_conv2d_fwd_f16(dlio_tensor4 x, const void *x_h2, const float *x_bound, const void *w_h2,
                                   const float *w_bound, const float *bias, dlio_conv cv, int act, dlio_tensor4 y,
                                   float *y_ptr, double *stats, void *stream) {
    // cv.sh == 2 (cv.sw == 1): computed as a stride-1 convolution, even rows stored (include/deeplio_b200.h)
    const int hdec = cv.sh == 2 ? 2 : 1;
    int rc;
    if (hdec == 2) {
        DLIO_CHECK_ARG(valid_t4(x) && valid_t4(y) && x.n == y.n && cv.sw == 1 && cv.kh == 2 * cv.ph + 1 &&
                           cv.kw == 2 * cv.pw + 1 && x.w == y.w && y.h == (x.h - 1) / 2 + 1,
                       "conv2d_fwd_f16: bad geometry for an H-stride-2 convolution");
    } else if ((rc = same_conv_check(x, y, cv, "conv2d_fwd_f16"))) {
        return rc;
    }
    DLIO_CHECK_ARG(x_h2 && x_bound && w_h2 && w_bound && y_ptr, "conv2d_fwd_f16: null pointer");
    ConvArgs a;
    a.x = Geo(x); a.y = Geo(y); a.o = Geo(y);
    a.kh = cv.kh; a.kw = cv.kw; a.sh = 1; a.sw = 1; a.ph = cv.ph; a.pw = cv.pw;
    a.cin = x.c; a.cout = y.c; a.act = act;
    a.x_hi = a.x_lo = a.w_hi = a.w_lo = nullptr;
    a.x_h2 = (const __half *)x_h2; a.w_h2 = (const __half *)w_h2; a.x_bound = x_bound; a.w_bound = w_bound;
    a.bias = bias; a.out = y_ptr; a.stats = stats; a.p_chunk = 0; a.hdec = hdec;
    rc = conv_tc_fwd(a, DLIO_PROF_CONV_FWD_TC, (cudaStream_t)stream);
    if (rc < 0) return rc;
    DLIO_CHECK_ARG(rc == 1, "conv2d_fwd_f16: shape not supported (cin %d %% 64, cout %d %% 16, input pads %d,%d >= %d,%d)",
                   x.c, y.c, x.ph, x.pw, cv.ph, cv.pw);
    return DLIO_OK;
}

// First layer, "folded split" operands (conv_s2d.cu): x is the [n, h, w/4, 64] view of the packed 8-channel input
// planes (per pixel [8 hi | 8 lo]; pads ph, 1), y the [n, h, w/4, R * cout] view of the fp32 output.
extern "C" int dlio_conv2d_fwd_f16_folded(dlio_tensor4 x, const void *x_h2, const float *x_bound, const void *w4_h2,
                                          const float *w_bound, const float *bias, dlio_conv cv, int act,
                                          dlio_tensor4 y, float *y_ptr, double *stats, void *stream) {
    int rc = same_conv_check(x, y, cv, "conv2d_fwd_f16_folded");
    if (rc) return rc;
    DLIO_CHECK_ARG(x_h2 && x_bound && w4_h2 && w_bound && y_ptr && x.c == 64, "conv2d_fwd_f16_folded: bad argument");
    ConvArgs a;
    a.x = Geo(x); a.y = Geo(y); a.o = Geo(y);
    a.kh = cv.kh; a.kw = cv.kw; a.sh = 1; a.sw = 1; a.ph = cv.ph; a.pw = cv.pw;
    a.cin = x.c; a.cout = y.c; a.act = act;
    a.x_hi = a.x_lo = a.w_hi = a.w_lo = nullptr;
    a.x_h2 = (const __half *)x_h2; a.w_h2 = (const __half *)w4_h2; a.x_bound = x_bound; a.w_bound = w_bound;
    a.bias = bias; a.out = y_ptr; a.stats = stats; a.p_chunk = 0; a.fold = 1;
    rc = conv_tc_fwd(a, DLIO_PROF_CONV_FWD_TC, (cudaStream_t)stream);
    if (rc < 0) return rc;
    DLIO_CHECK_ARG(rc == 1, "conv2d_fwd_f16_folded: shape not supported (cout %d %% 16, input pads %d,%d >= %d,%d)", y.c,
                   x.ph, x.pw, cv.ph, cv.pw);
    return DLIO_OK;
}

// dw64 [dy.c][kh][kw][64] (fp32, K index k = p * 16 + part * 8 + c: dlio_weight_grad_from_s2d_f16 folds it to OIHW).
// dy: the [n, h, w/4, R * 64] view (pads as x) of dy planes stored per OUTPUT pixel as [64 hi | 64 lo].
extern "C" int dlio_conv2d_bwd_weight_f16_folded(dlio_tensor4 x, const void *x_h2, const float *x_bound, dlio_tensor4 dy,
                                                 const void *dy_h2, const float *dy_bound, dlio_conv cv, float *dw64,
                                                 void *stream) {
    int rc = same_conv_check(x, dy, cv, "conv2d_bwd_weight_f16_folded");
    if (rc) return rc;
    DLIO_CHECK_ARG(x_h2 && x_bound && dy_h2 && dy_bound && dw64 && x.c == 64 && dy.c % 128 == 0,
                   "conv2d_bwd_weight_f16_folded: bad argument (x.c == 64, dy.c a multiple of 128)");
    ConvArgs a;
    a.x = Geo(x); a.y = Geo(dy); a.o = Geo(dy);
    a.kh = cv.kh; a.kw = cv.kw; a.sh = 1; a.sw = 1; a.ph = cv.ph; a.pw = cv.pw;
    a.cin = x.c; a.cout = dy.c; a.act = 0;
    a.x_hi = a.x_lo = a.w_hi = a.w_lo = nullptr;
    a.x_h2 = (const __half *)x_h2; a.w_h2 = (const __half *)dy_h2; a.x_bound = x_bound; a.w_bound = dy_bound;
    a.bias = nullptr; a.out = dw64; a.stats = nullptr; a.p_chunk = 0; a.fold = 1;
    rc = conv_tc_wgrad(a, (cudaStream_t)stream);
    if (rc < 0) return rc;
    DLIO_CHECK_ARG(rc == 1, "conv2d_bwd_weight_f16_folded: shape not supported (shared padded grid)");
    return DLIO_OK;
}

// max |x| over n floats -> *bound (one float in device memory; the scale of packed fp16 planes derives from it)
extern "C" int dlio_absmax(const float *x, long long n, float *bound, void *stream) {
    DLIO_CHECK_ARG(x && bound && n > 0, "absmax: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, st);
    DLIO_CUDA(cudaMemsetAsync(bound, 0, sizeof(float), st));
    long long grid = (n + 256 * 16 - 1) / (256 * 16);
    absmax_kernel<<<(unsigned)(grid > 2368 ? 2368 : grid), 256, 0, st>>>(x, n, bound);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_conv2d_bwd_data_f16(dlio_tensor4 dy, const void *dy_h2, const float *dy_bound, const void *wt_h2,
                                        const float *w_bound, dlio_conv cv, dlio_tensor4 dx, float *dx_ptr,
                                        int accumulate, void *stream) {
    int rc = same_conv_check(dx, dy, cv, "conv2d_bwd_data_f16");
    if (rc) return rc;
    DLIO_CHECK_ARG(dy_h2 && dy_bound && wt_h2 && w_bound && dx_ptr, "conv2d_bwd_data_f16: null pointer");
    // dgrad = stride-1 convolution of the padded dy with the flipped / transposed weights
    ConvArgs t;
    t.x = Geo(dy); t.y = Geo(dx); t.o = Geo(dx);
    t.kh = cv.kh; t.kw = cv.kw; t.sh = 1; t.sw = 1; t.ph = cv.kh - 1 - cv.ph; t.pw = cv.kw - 1 - cv.pw;
    t.cin = dy.c; t.cout = dx.c; t.act = 0;
    t.x_hi = t.x_lo = t.w_hi = t.w_lo = nullptr;
    t.x_h2 = (const __half *)dy_h2; t.w_h2 = (const __half *)wt_h2; t.x_bound = dy_bound; t.w_bound = w_bound;
    t.bias = nullptr; t.out = dx_ptr; t.stats = nullptr; t.p_chunk = 0; t.accum = accumulate ? 1 : 0;
    rc = conv_tc_fwd(t, DLIO_PROF_CONV_DGRAD_TC, (cudaStream_t)stream);
    if (rc < 0) return rc;
    DLIO_CHECK_ARG(rc == 1, "conv2d_bwd_data_f16: shape not supported (cout %d %% 64, cin %d %% 16, dy pads %d,%d)", dy.c,
                   dx.c, dy.ph, dy.pw);
    return DLIO_OK;
}

extern "C" int dlio_conv2d_bwd_weight_f16(dlio_tensor4 x, const void *x_h2, const float *x_bound, dlio_tensor4 dy,
                                          const void *dy_h2, const float *dy_bound, dlio_conv cv, float *dw,
                                          void *stream) {
    int rc = same_conv_check(x, dy, cv, "conv2d_bwd_weight_f16");
    if (rc) return rc;
    DLIO_CHECK_ARG(x_h2 && x_bound && dy_h2 && dy_bound && dw, "conv2d_bwd_weight_f16: null pointer");
    ConvArgs a;
    a.x = Geo(x); a.y = Geo(dy); a.o = Geo(dy);
    a.kh = cv.kh; a.kw = cv.kw; a.sh = 1; a.sw = 1; a.ph = cv.ph; a.pw = cv.pw;
    a.cin = x.c; a.cout = dy.c; a.act = 0;
    a.x_hi = a.x_lo = a.w_hi = a.w_lo = nullptr;
    a.x_h2 = (const __half *)x_h2; a.w_h2 = (const __half *)dy_h2; a.x_bound = x_bound; a.w_bound = dy_bound;
    a.bias = nullptr; a.out = dw; a.stats = nullptr; a.p_chunk = 0;
    rc = conv_tc_wgrad(a, (cudaStream_t)stream);
    if (rc < 0) return rc;
    DLIO_CHECK_ARG(rc == 1, "conv2d_bwd_weight_f16: shape not supported (cin %d %% 64, cout %d %% 64, shared padded grid)",
                   x.c, dy.c);
    return DLIO_OK;
}
