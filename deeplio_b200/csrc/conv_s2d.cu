// First-layer convolutions on the tensor cores through a space-to-depth reinterpretation.
//
// The first convolution of every encoder has 6 (padded 8) input channels and a wide kernel (5x7 or 3x5, stride
// (1,2) or (1,1)); as an implicit GEMM it has K = kh*kw*8 with 8-channel rows, which the tcgen05 kernel (32-float
// K chunks) cannot take.  Grouping FOUR consecutive pixels of a row gives a [n, h, w/4, 32] view of the SAME
// padded-NHWC memory (row pads of 4 pixels become 1), and grouping the R = 4/stride outputs computed from them
// gives a [n, h, w/4, R*cout] view of the SAME output memory.  In these views the layer is an ordinary
// stride-1 "same" convolution with kernel kh x 3, Cin = 32, Cout = R*cout -- only the weights are rearranged:
//     W4[(r, co)][dy][t][(p, c)] = W[co][dy][dx][c]   with dx = 4 t' + p - stride*r + pad_w,  t' = t - 1,
// zero where dx falls outside [0, kw).  2.3x more MACs than the direct form, but on the tensor pipe.
#include "common.cuh"

namespace dlio {

__global__ void weight_to_s2d_kernel(const float *__restrict__ w, int cout, int cin, int kh, int kw, int sw,
                                     float *__restrict__ hi, float *__restrict__ lo) {
    const int R = 4 / sw, pw = (kw - 1) / 2;
    const long long total = (long long)R * cout * kh * 3 * 32;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int pc = (int)(i % 32);
    long long t1 = i / 32;
    int t = (int)(t1 % 3);
    t1 /= 3;
    int dy = (int)(t1 % kh);
    int rco = (int)(t1 / kh);
    int p = pc / 8, c = pc % 8;
    int r = rco / cout, co = rco - r * cout;
    int dx = 4 * (t - 1) + p - sw * r + pw;
    float v = 0.f;
    if (c < cin && dx >= 0 && dx < kw) v = w[(((size_t)co * cin + c) * kh + dy) * kw + dx];   // OIHW
    float h, l;
    tf32_split(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
}

// dw[co][c][dy][dx] (OIHW) = sum over the (r, t, p) that map to dx of dw4[(r, co)][dy][t][(p, c)]
__global__ void weight_grad_from_s2d_kernel(const float *__restrict__ dw4, int cout, int cin, int kh, int kw, int sw,
                                            float *__restrict__ dw) {
    const int R = 4 / sw, pw = (kw - 1) / 2;
    const long long total = (long long)cout * cin * kh * kw;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int dx = (int)(i % kw);
    long long t1 = i / kw;
    int dy = (int)(t1 % kh);
    t1 /= kh;
    int c = (int)(t1 % cin);
    int co = (int)(t1 / cin);
    float s = 0.f;
    for (int r = 0; r < R; ++r) {
        int off = dx + sw * r - pw;          // = 4 t' + p
        int tq = (off + 8) / 4 - 2;          // floor(off / 4) for off >= -8
        int p = off - 4 * tq;
        int t = tq + 1;
        if (t >= 0 && t < 3) s += dw4[((((size_t)(r * cout + co)) * kh + dy) * 3 + t) * 32 + p * 8 + c];
    }
    dw[i] = s;
}

// "Folded split" operands of the first layer (fp16 tensor-core path without a 64-byte-swizzle kernel).  The input
// planes hold, per pixel, [8 hi | 8 lo] halves (common.cuh, fp16 operand split), so a group of four pixels is ONE
// 128-byte row of 64 halves, K index k = p * 16 + part * 8 + c (part 0 = hi, 1 = lo).  With the weight planes
//     B_main[k] = part == 0 ? w_hi : 0          B_corr[k] = part == 0 ? w_lo : w_hi
// the kernel's "hi" product x . B_main is x_hi * w_hi and its "hi x lo" product x . B_corr is x_hi * w_lo + x_lo * w_hi:
// the two accumulators of the split scheme from TWO products over a K of 64 instead of three over a K of 32, and no
// second plane of x.  Rows: (r, co), columns: (dy, t, k); layout [R * cout][2][kh * 3 * 64] as every packed weight.
__global__ void __launch_bounds__(256) weight_to_s2d_f16_kernel(const float *__restrict__ w, int cout, int cin, int kh,
                                                                int kw, int sw, const float *bound,
                                                                __half *__restrict__ out) {
    const int R = 4 / sw, pw = (kw - 1) / 2;
    const long long K = (long long)kh * 3 * 64;
    const long long total = (long long)R * cout * K;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float s = f16_scale_from_bound(*bound);
    const int k = (int)(i % 64);
    long long t1 = i / 64;
    const int t = (int)(t1 % 3);
    t1 /= 3;
    const int dy = (int)(t1 % kh);
    const int rco = (int)(t1 / kh);
    const int p = k >> 4, part = (k >> 3) & 1, c = k & 7;
    const int r = rco / cout, co = rco - r * cout;
    const int dx = 4 * (t - 1) + p - sw * r + pw;
    float v = 0.f;
    if (c < cin && dx >= 0 && dx < kw) v = w[(((size_t)co * cin + c) * kh + dy) * kw + dx];   // OIHW
    __half h, l;
    f16_split(v * s, h, l);
    const size_t col = (size_t)(i % K);
    out[(size_t)rco * 2 * K + col] = part == 0 ? h : __float2half(0.f);
    out[(size_t)rco * 2 * K + K + col] = part == 0 ? l : h;
}

// dw[co][c][dy][dx] (OIHW) = sum over (r, t, p) that map to dx, and over both parts, of dw64[(r, co)][dy][t][k]
__global__ void weight_grad_from_s2d_f16_kernel(const float *__restrict__ dw64, int cout, int cin, int kh, int kw,
                                                int sw, float *__restrict__ dw) {
    const int R = 4 / sw, pw = (kw - 1) / 2;
    const long long total = (long long)cout * cin * kh * kw;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int dx = (int)(i % kw);
    long long t1 = i / kw;
    int dy = (int)(t1 % kh);
    t1 /= kh;
    int c = (int)(t1 % cin);
    int co = (int)(t1 / cin);
    float s = 0.f;
    for (int r = 0; r < R; ++r) {
        int off = dx + sw * r - pw;          // = 4 t' + p
        int tq = (off + 8) / 4 - 2;          // floor(off / 4) for off >= -8
        int p = off - 4 * tq;
        int t = tq + 1;
        if (t >= 0 && t < 3) {
            const float *g = dw64 + ((((size_t)(r * cout + co)) * kh + dy) * 3 + t) * 64 + p * 16 + c;
            s += g[0] + g[8];
        }
    }
    dw[i] = s;
}

// out[j*c + k] = sum_r in[j*R*c + r*c + k]   (j = 0: sums, j = 1: sums of squares)
__global__ void fold_stats_kernel(const double *__restrict__ in, int R, int c, double *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * c) return;
    int j = i / c, k = i - j * c;
    double s = 0.0;
    for (int r = 0; r < R; ++r) s += in[(size_t)j * R * c + r * c + k];
    out[i] = s;
}

// ---------------------------------------------------------------- pixel-pair view of W-stride-2 convolutions
// value of the pair-view weight w2[co][dy][t][p * cin + c] (zero where the source column falls outside the kernel)
__device__ __forceinline__ float pair_weight(const float *__restrict__ w, int cin, int kh, int kw, int kw2, int co,
                                             int dy, int t, int pc) {
    const int p = pc >= cin ? 1 : 0, c = pc - p * cin;
    const int dx = 2 * (t - (kw2 - 1) / 2) + p + (kw - 1) / 2;
    return (dx >= 0 && dx < kw) ? w[(((size_t)co * cin + c) * kh + dy) * kw + dx] : 0.f;
}

// packed fp16 split rows of w2 (forward / wgrad layout, or flipped + transposed for dgrad): see weight_pack_f16_kernel
__global__ void __launch_bounds__(256) weight_pack_pair_f16_kernel(const float *__restrict__ w, int cout, int cin,
                                                                   int kh, int kw, int kw2, int flip,
                                                                   const float *bound, __half *__restrict__ out) {
    // four consecutive output columns per thread (cin and cout are multiples of 4): see weight_pack_f16_kernel
    const int taps = kh * kw2, c2 = 2 * cin;
    const long long total4 = (long long)cout * taps * c2 / 4;
    const long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const long long i = i4 * 4;
    const float s = f16_scale_from_bound(*bound);
    int row, col;
    long long K;
    __half h[4], l[4];
    if (!flip) {
        const int pc = (int)(i % c2);
        long long t = i / c2;
        const int tap = (int)(t % taps);
        const int co = (int)(t / taps);
        row = co; col = tap * c2 + pc; K = (long long)taps * c2;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            f16_split(pair_weight(w, cin, kh, kw, kw2, co, tap / kw2, tap % kw2, pc + j) * s, h[j], l[j]);
    } else {
        const int co = (int)(i % cout);
        long long t = i / cout;
        const int tapf = (int)(t % taps);
        const int pc = (int)(t / taps);
        const int tap = taps - 1 - tapf;
        row = pc; col = tapf * cout + co; K = (long long)taps * cout;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            f16_split(pair_weight(w, cin, kh, kw, kw2, co + j, tap / kw2, tap % kw2, pc) * s, h[j], l[j]);
    }
    __half *o = out + (size_t)row * 2 * K + col;
    *reinterpret_cast<uint2 *>(o) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
    *reinterpret_cast<uint2 *>(o + K) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
}

// dw[co][c][dy][dx] (OIHW) = dw2[co][dy][t][p * cin + c] for the one (t, p) that maps to dx
__global__ void weight_grad_from_pair_kernel(const float *__restrict__ dw2, int cout, int cin, int kh, int kw, int kw2,
                                             float *__restrict__ dw) {
    const long long total = (long long)cout * cin * kh * kw;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int dx = (int)(i % kw);
    long long t1 = i / kw;
    const int dy = (int)(t1 % kh);
    t1 /= kh;
    const int c = (int)(t1 % cin);
    const int co = (int)(t1 / cin);
    const int off = dx - (kw - 1) / 2;           // = 2 t' + p
    const int tq = (off + 8) / 2 - 4;            // floor(off / 2)
    const int p = off - 2 * tq, t = tq + (kw2 - 1) / 2;
    dw[i] = dw2[((((size_t)co * kh + dy) * kw2 + t) * 2 + p) * cin + c];
}

__global__ void __launch_bounds__(256) absmax_plain_kernel(const float *__restrict__ x, long long n, float *bound) {
    float m = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_nonneg(bound, m);
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_weight_pack_pair_f16(const float *w_oihw, int cout, int cin, int kh, int kw, int transpose_flip,
                                         int compute_bound, float *w_bound, void *w_h2, void *stream) {
    DLIO_CHECK_ARG(w_oihw && w_bound && w_h2 && cout > 0 && cin > 0 && kh > 0 && (kw == 1 || kw == 3 || kw == 5),
                   "weight_pack_pair_f16: bad argument (kw must be 1, 3 or 5)");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, st);
    const long long n = (long long)cout * cin * kh * kw;
    if (compute_bound) {
        DLIO_CUDA(cudaMemsetAsync(w_bound, 0, sizeof(float), st));
        int grid = ceil_div(n, 256 * 8);
        absmax_plain_kernel<<<grid > 592 ? 592 : grid, 256, 0, st>>>(w_oihw, n, w_bound);
        DLIO_LAUNCH_CHECK();
    }
    const int kw2 = kw > 1 ? 3 : 1;
    DLIO_CHECK_ARG(cin % 4 == 0 && cout % 4 == 0 && (((uintptr_t)w_h2) & 7) == 0,
                   "weight_pack_pair_f16: cin and cout must be multiples of 4");
    const long long total = (long long)cout * kh * kw2 * 2 * cin / 4;
    weight_pack_pair_f16_kernel<<<ceil_div(total, 256), 256, 0, st>>>(w_oihw, cout, cin, kh, kw, kw2, transpose_flip,
                                                                     w_bound, (__half *)w_h2);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_weight_grad_from_pair(const float *dw2, int cout, int cin, int kh, int kw, float *dw_oihw,
                                          void *stream) {
    DLIO_CHECK_ARG(dw2 && dw_oihw && cout > 0 && cin > 0 && kh > 0 && (kw == 1 || kw == 3 || kw == 5),
                   "weight_grad_from_pair: bad argument (kw must be 1, 3 or 5)");
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    const long long total = (long long)cout * cin * kh * kw;
    weight_grad_from_pair_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dw2, cout, cin, kh, kw,
                                                                                        kw > 1 ? 3 : 1, dw_oihw);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_weight_to_s2d(const float *w_oihw, int cout, int cin, int kh, int kw, int sw, float *w4_hi,
                                  float *w4_lo, void *stream) {
    DLIO_CHECK_ARG(w_oihw && w4_hi && cout > 0 && cin > 0 && cin <= 8 && (sw == 1 || sw == 2) && kw >= 1 && kw <= 7 &&
                       (kw & 1),
                   "weight_to_s2d: bad argument");
    long long total = (long long)(4 / sw) * cout * kh * 3 * 32;
    weight_to_s2d_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, cout, cin, kh, kw, sw, w4_hi, w4_lo);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_weight_grad_from_s2d(const float *dw4, int cout, int cin, int kh, int kw, int sw, float *dw_oihw,
                                         void *stream) {
    DLIO_CHECK_ARG(dw4 && dw_oihw && cout > 0 && cin > 0 && cin <= 8 && (sw == 1 || sw == 2) && kw <= 7 && (kw & 1),
                   "weight_grad_from_s2d: bad argument");
    long long total = (long long)cout * cin * kh * kw;
    weight_grad_from_s2d_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dw4, cout, cin, kh, kw, sw, dw_oihw);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_weight_to_s2d_f16(const float *w_oihw, int cout, int cin, int kh, int kw, int sw, float *w_bound,
                                      void *w4_h2, void *stream) {
    DLIO_CHECK_ARG(w_oihw && w_bound && w4_h2 && cout > 0 && cin > 0 && cin <= 8 && (sw == 1 || sw == 2) && kw >= 1 &&
                       kw <= 7 && (kw & 1),
                   "weight_to_s2d_f16: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, st);
    const long long n = (long long)cout * cin * kh * kw;
    DLIO_CUDA(cudaMemsetAsync(w_bound, 0, sizeof(float), st));
    int grid = ceil_div(n, 256 * 8);
    absmax_plain_kernel<<<grid > 592 ? 592 : grid, 256, 0, st>>>(w_oihw, n, w_bound);
    DLIO_LAUNCH_CHECK();
    const long long total = (long long)(4 / sw) * cout * kh * 3 * 64;
    weight_to_s2d_f16_kernel<<<ceil_div(total, 256), 256, 0, st>>>(w_oihw, cout, cin, kh, kw, sw, w_bound, (__half *)w4_h2);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_weight_grad_from_s2d_f16(const float *dw64, int cout, int cin, int kh, int kw, int sw,
                                             float *dw_oihw, void *stream) {
    DLIO_CHECK_ARG(dw64 && dw_oihw && cout > 0 && cin > 0 && cin <= 8 && (sw == 1 || sw == 2) && kw <= 7 && (kw & 1),
                   "weight_grad_from_s2d_f16: bad argument");
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    long long total = (long long)cout * cin * kh * kw;
    weight_grad_from_s2d_f16_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dw64, cout, cin, kh, kw, sw,
                                                                                           dw_oihw);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_fold_stats(const double *in, int r, int c, double *out, void *stream) {
    DLIO_CHECK_ARG(in && out && r >= 1 && c > 0, "fold_stats: bad argument");
    fold_stats_kernel<<<ceil_div(2 * c, 128), 128, 0, (cudaStream_t)stream>>>(in, r, c, out);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
