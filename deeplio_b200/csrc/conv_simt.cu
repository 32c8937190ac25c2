// Generic fp32 implicit-GEMM convolution kernels (CUDA cores): forward, dgrad, wgrad.
//
// These cover every convolution shape on the DeepLIO path (kernels 5x7 / 3x5 / 3x3 / 1x1, strides
// (1,1) (1,2) (2,2), Cin = 6 padded to 8 ...).  The tcgen05 kernels in conv_tc.cu take over the
// stride-1, Cin % 32 == 0 layers, which hold > 95 % of the FLOPs; what stays here is HBM-bound
// (first layers) or small.  Layout: padded NHWC activations, OHWI weights (see deeplio_b200.h).
//
// Tiling: 64 x 64 output tile per CTA, K step 16, 256 threads, 4 x 4 register tile per thread,
// operands staged in shared memory with a register prefetch of the next K step.
#include "common.cuh"
#include "conv.cuh"

namespace dlio {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256, LDS = BM + 4;

__device__ __forceinline__ void mma_tile(const float (*As)[LDS], const float (*Bs)[LDS], float (&acc)[4][4],
                                         int ty, int tx) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
        float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(NT) conv_fwd_kernel(ConvArgs a) {
    __shared__ __align__(16) float As[BK][LDS];
    __shared__ __align__(16) float Bs[BK][LDS];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long M = (long long)a.y.n * a.y.h * a.y.w;
    const int K = a.kh * a.kw * a.cin;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A loader: one pixel row, 4 consecutive k
    const int ar = tid >> 2, kq = (tid & 3) * 4;
    const long long am = m0 + ar;
    const bool am_ok = am < M;
    int an = 0, hi0 = 0, wi0 = 0;
    if (am_ok) {
        int wo = (int)(am % a.y.w);
        long long t = am / a.y.w;
        int ho = (int)(t % a.y.h);
        an = (int)(t / a.y.h);
        hi0 = ho * a.sh - a.ph;
        wi0 = wo * a.sw - a.pw;
    }
    const int bco = n0 + ar;  // B loader: one cout row, same 4 k
    const bool b_ok = bco < a.cout;

    float4 ra, rb;
    auto load = [&](int k0) {
        const int k = k0 + kq;
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = ra;
        if (k < K) {
            if (am_ok) {
                int tap = k / a.cin, ci = k - tap * a.cin;
                int khh = tap / a.kw, kww = tap - khh * a.kw;
                int hi = hi0 + khh, wi = wi0 + kww;
                if (hi >= 0 && hi < a.x.h && wi >= 0 && wi < a.x.w)
                    ra = ld4_sum(a.x_hi, a.x_lo, a.x.off(an, hi, wi) + ci);
            }
            if (b_ok) rb = ld4_sum(a.w_hi, a.w_lo, (size_t)bco * K + k);
        }
    };
    auto stage = [&]() {
        As[kq + 0][ar] = ra.x; As[kq + 1][ar] = ra.y; As[kq + 2][ar] = ra.z; As[kq + 3][ar] = ra.w;
        Bs[kq + 0][ar] = rb.x; Bs[kq + 1][ar] = rb.y; Bs[kq + 2][ar] = rb.z; Bs[kq + 3][ar] = rb.w;
    };

    float acc[4][4] = {};
    const int nt = (K + BK - 1) / BK;
    load(0);
    stage();
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) load((t + 1) * BK);
        mma_tile(As, Bs, acc, ty, tx);
        __syncthreads();
        if (t + 1 < nt) {
            stage();
            __syncthreads();
        }
    }

    // epilogue: bias + activation, store, per-channel statistics
    const int co = n0 + tx * 4;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias && co < a.cout) bias = ld4(a.bias + co);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m < M && co < a.cout) {
            int wo = (int)(m % a.y.w);
            long long t = m / a.y.w;
            int ho = (int)(t % a.y.h);
            int n = (int)(t / a.y.h);
            float4 v = make_float4(act_apply(acc[i][0] + bias.x, a.act), act_apply(acc[i][1] + bias.y, a.act),
                                   act_apply(acc[i][2] + bias.z, a.act), act_apply(acc[i][3] + bias.w, a.act));
            st4(a.out + a.y.off(n, ho, wo) + co, v);
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
        }
    }
    if (a.stats) {
        float (*red)[LDS] = As;      // [16][64] sums, reuse (all threads are past the last mma)
        float (*red2)[LDS] = Bs;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[ty][tx * 4 + j] = s[j];
            red2[ty][tx * 4 + j] = q[j];
        }
        __syncthreads();
        if (tid < BN && n0 + tid < a.cout) {
            float ss = 0.f, qq = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                ss += red[r][tid];
                qq += red2[r][tid];
            }
            atomicAdd(a.stats + n0 + tid, (double)ss);
            atomicAdd(a.stats + a.cout + n0 + tid, (double)qq);
        }
    }
}

// ------------------------------------------------------------------ dgrad: dx[n,h,w,ci] = sum dy[n,ho,wo,co] w[co,kh,kw,ci]
__global__ void __launch_bounds__(NT) conv_dgrad_kernel(ConvArgs a) {
    __shared__ __align__(16) float As[BK][LDS];
    __shared__ __align__(16) float Bs[BK][LDS];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long M = (long long)a.x.n * a.x.h * a.x.w;  // input pixels
    const int K = a.kh * a.kw * a.cout;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    const int ar = tid >> 2, kq = (tid & 3) * 4;
    const long long am = m0 + ar;
    const bool am_ok = am < M;
    int an = 0, h = 0, w = 0;
    if (am_ok) {
        w = (int)(am % a.x.w);
        long long t = am / a.x.w;
        h = (int)(t % a.x.h);
        an = (int)(t / a.x.h);
    }
    const int bk = tid >> 4, bq = (tid & 15) * 4;  // B loader: k row, 4 consecutive cin
    const int bci = n0 + bq;

    float4 ra, rb;
    auto load = [&](int k0) {
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = ra;
        int k = k0 + kq;
        if (k < K && am_ok) {
            int tap = k / a.cout, co = k - tap * a.cout;
            int khh = tap / a.kw, kww = tap - khh * a.kw;
            int th = h + a.ph - khh, tw = w + a.pw - kww;
            if (th >= 0 && tw >= 0) {
                int ho = th / a.sh, wo = tw / a.sw;
                if (ho * a.sh == th && wo * a.sw == tw && ho < a.y.h && wo < a.y.w)
                    ra = ld4_sum(a.x_hi, a.x_lo, a.y.off(an, ho, wo) + co);  // x_hi/x_lo carry dy here
            }
        }
        k = k0 + bk;
        if (k < K && bci < a.cin) {
            int tap = k / a.cout, co = k - tap * a.cout;
            rb = ld4_sum(a.w_hi, a.w_lo, ((size_t)co * a.kh * a.kw + tap) * a.cin + bci);
        }
    };
    auto stage = [&]() {
        As[kq + 0][ar] = ra.x; As[kq + 1][ar] = ra.y; As[kq + 2][ar] = ra.z; As[kq + 3][ar] = ra.w;
        *reinterpret_cast<float4 *>(&Bs[bk][bq]) = rb;
    };

    float acc[4][4] = {};
    const int nt = (K + BK - 1) / BK;
    load(0);
    stage();
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) load((t + 1) * BK);
        mma_tile(As, Bs, acc, ty, tx);
        __syncthreads();
        if (t + 1 < nt) {
            stage();
            __syncthreads();
        }
    }
    const int ci = n0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m < M && ci < a.cin) {
            int ww = (int)(m % a.x.w);
            long long t = m / a.x.w;
            int hh = (int)(t % a.x.h);
            int n = (int)(t / a.x.h);
            float *op = a.out + a.o.off(n, hh, ww) + ci;
            float4 o4 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (a.accum) o4 = add4(o4, ld4(op));
            st4(op, o4);
        }
    }
}

// ------------------------------------------------------------------ wgrad: dw[co,(kh,kw,ci)] = sum_p dy[p,co] x[p shifted, ci]
__global__ void __launch_bounds__(NT) conv_wgrad_kernel(ConvArgs a) {
    __shared__ __align__(16) float As[BK][LDS];
    __shared__ __align__(16) float Bs[BK][LDS];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long P = (long long)a.y.n * a.y.h * a.y.w;
    const int KW = a.kh * a.kw * a.cin;
    const int m0 = blockIdx.x * BM;  // cout tile
    const int n0 = blockIdx.y * BN;  // (tap, ci) tile
    const long long p_begin = (long long)blockIdx.z * a.p_chunk;
    long long p_end = p_begin + a.p_chunk;
    if (p_end > P) p_end = P;

    const int pr = tid >> 4, cq = (tid & 15) * 4;
    const int aco = m0 + cq;
    const int bkidx = n0 + cq;
    int khh = 0, kww = 0, bci = 0;
    const bool b_ok = bkidx < KW;
    if (b_ok) {
        int tap = bkidx / a.cin;
        bci = bkidx - tap * a.cin;
        khh = tap / a.kw;
        kww = tap - khh * a.kw;
    }
    float4 ra, rb;
    auto load = [&](long long p0) {
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = ra;
        long long p = p0 + pr;
        if (p < p_end) {
            int wo = (int)(p % a.y.w);
            long long t = p / a.y.w;
            int ho = (int)(t % a.y.h);
            int n = (int)(t / a.y.h);
            if (aco < a.cout) ra = ld4_sum(a.w_hi, a.w_lo, a.y.off(n, ho, wo) + aco);  // w_hi/w_lo carry dy here
            if (b_ok) {
                int hi = ho * a.sh - a.ph + khh, wi = wo * a.sw - a.pw + kww;
                if (hi >= 0 && hi < a.x.h && wi >= 0 && wi < a.x.w)
                    rb = ld4_sum(a.x_hi, a.x_lo, a.x.off(n, hi, wi) + bci);
            }
        }
    };
    auto stage = [&]() {
        *reinterpret_cast<float4 *>(&As[pr][cq]) = ra;
        *reinterpret_cast<float4 *>(&Bs[pr][cq]) = rb;
    };
    float acc[4][4] = {};
    if (p_begin < p_end) {
        load(p_begin);
        stage();
        __syncthreads();
        for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
            bool more = p0 + BK < p_end;
            if (more) load(p0 + BK);
            mma_tile(As, Bs, acc, ty, tx);
            __syncthreads();
            if (more) {
                stage();
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = m0 + ty * 4 + i;
        if (co >= a.cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int kidx = n0 + tx * 4 + j;
            if (kidx < KW) atomicAdd(a.out + (size_t)co * KW + kidx, acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------ weight layout helpers
__global__ void weight_to_ohwi_kernel(const float *__restrict__ w, int cout, int cin, int khw, int cin_pad,
                                      float *__restrict__ hi, float *__restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)cout * khw * cin_pad;
    if (i >= total) return;
    int ci = (int)(i % cin_pad);
    long long t = i / cin_pad;
    int tap = (int)(t % khw);
    int co = (int)(t / khw);
    float v = ci < cin ? w[((size_t)co * cin + ci) * khw + tap] : 0.f;
    if (lo) {
        float h, l;
        tf32_split(v, h, l);
        hi[i] = h;
        lo[i] = l;
    } else {
        hi[i] = v;
    }
}
__global__ void weight_grad_to_oihw_kernel(const float *__restrict__ dw, int cout, int cin, int khw, int cin_pad,
                                           float *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)cout * cin * khw;
    if (i >= total) return;
    int tap = (int)(i % khw);
    long long t = i / khw;
    int ci = (int)(t % cin);
    int co = (int)(t / cin);
    out[i] = dw[((size_t)co * khw + tap) * cin_pad + ci];
}
__global__ void weight_flip_transpose_kernel(const float *__restrict__ w, int cout, int cin, int kh, int kw,
                                             float *__restrict__ hi, float *__restrict__ lo) {
    // out[ci][kh'][kw'][co] = w[co][kh-1-kh'][kw-1-kw'][ci]
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)cout * cin * kh * kw;
    if (i >= total) return;
    int co = (int)(i % cout);
    long long t = i / cout;
    int kwp = (int)(t % kw);
    t /= kw;
    int khp = (int)(t % kh);
    int ci = (int)(t / kh);
    float v = w[(((size_t)co * kh + (kh - 1 - khp)) * kw + (kw - 1 - kwp)) * cin + ci];
    if (lo) {
        float h, l;
        tf32_split(v, h, l);
        hi[i] = h;
        lo[i] = l;
    } else {
        hi[i] = v;
    }
}

static int check_conv(const dlio_tensor4 &x, const dlio_tensor4 &y, const dlio_conv &cv) {
    DLIO_CHECK_ARG(valid_t4(x) && valid_t4(y), "conv: bad tensor descriptor");
    DLIO_CHECK_ARG(cv.kh > 0 && cv.kw > 0 && cv.sh > 0 && cv.sw > 0 && cv.ph >= 0 && cv.pw >= 0, "conv: bad conv descriptor");
    DLIO_CHECK_ARG(x.c % 4 == 0 && y.c % 4 == 0, "conv: channel counts must be multiples of 4 (got %d, %d)", x.c, y.c);
    DLIO_CHECK_ARG(x.n == y.n, "conv: batch mismatch");
    int ho = (x.h + 2 * cv.ph - cv.kh) / cv.sh + 1, wo = (x.w + 2 * cv.pw - cv.kw) / cv.sw + 1;
    DLIO_CHECK_ARG(ho == y.h && wo == y.w, "conv: output extent %dx%d does not match %dx%d", y.h, y.w, ho, wo);
    return DLIO_OK;
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_conv2d_fwd(dlio_tensor4 x, const float *x_hi, const float *x_lo, const float *w_hi,
                               const float *w_lo, const float *bias, dlio_conv cv, int act, dlio_tensor4 y,
                               float *y_ptr, double *stats, void *stream) {
    int rc = check_conv(x, y, cv);
    if (rc) return rc;
    DLIO_CHECK_ARG(x_hi && w_hi && y_ptr, "conv_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    ConvArgs a;
    a.x = Geo(x); a.y = Geo(y); a.o = Geo(y);
    a.kh = cv.kh; a.kw = cv.kw; a.sh = cv.sh; a.sw = cv.sw; a.ph = cv.ph; a.pw = cv.pw;
    a.cin = x.c; a.cout = y.c; a.act = act;
    a.x_hi = x_hi; a.x_lo = x_lo; a.w_hi = w_hi; a.w_lo = w_lo; a.bias = bias;
    a.out = y_ptr; a.stats = stats; a.p_chunk = 0;
    rc = conv_tc_fwd(a, DLIO_PROF_CONV_FWD_TC, st);
    if (rc != 0) return rc < 0 ? rc : DLIO_OK;
    ProfScope prof(DLIO_PROF_CONV_FWD_SIMT, st);
    if (y.ph > 0 || y.pw > 0) DLIO_CUDA(cudaMemsetAsync(y_ptr, 0, a.y.numel() * sizeof(float), st));
    long long M = (long long)y.n * y.h * y.w;
    dim3 grid(ceil_div(M, BM), ceil_div(y.c, BN));
    conv_fwd_kernel<<<grid, NT, 0, st>>>(a);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_conv2d_bwd_data(dlio_tensor4 dy, const float *dy_hi, const float *dy_lo, const float *w_hi,
                                    const float *w_lo, const float *wt_hi, const float *wt_lo, dlio_conv cv,
                                    dlio_tensor4 dx, float *dx_ptr, int accumulate, void *stream) {
    int rc = check_conv(dx, dy, cv);
    if (rc) return rc;
    DLIO_CHECK_ARG(dy_hi && w_hi && dx_ptr, "conv_bwd_data: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (wt_hi && wt_lo && dy_lo) {
        // dgrad as a stride-1 convolution of the padded dy with the flipped / transposed weights (tcgen05 path)
        ConvArgs t;
        t.x = Geo(dy); t.y = Geo(dx); t.o = Geo(dx);
        t.kh = cv.kh; t.kw = cv.kw; t.sh = cv.sh; t.sw = cv.sw; t.ph = cv.kh - 1 - cv.ph; t.pw = cv.kw - 1 - cv.pw;
        t.cin = dy.c; t.cout = dx.c; t.act = 0;
        t.x_hi = dy_hi; t.x_lo = dy_lo; t.w_hi = wt_hi; t.w_lo = wt_lo; t.bias = nullptr;
        t.out = dx_ptr; t.stats = nullptr; t.p_chunk = 0; t.accum = accumulate ? 1 : 0;
        rc = conv_tc_fwd(t, DLIO_PROF_CONV_DGRAD_TC, st);
        if (rc != 0) return rc < 0 ? rc : DLIO_OK;
    }
    ConvArgs a;
    a.x = Geo(dx); a.x.ph = 0; a.x.pw = 0;  // logical extent only; addressing of dx goes through a.o
    a.y = Geo(dy); a.o = Geo(dx);
    a.kh = cv.kh; a.kw = cv.kw; a.sh = cv.sh; a.sw = cv.sw; a.ph = cv.ph; a.pw = cv.pw;
    a.cin = dx.c; a.cout = dy.c; a.act = 0;
    a.x_hi = dy_hi; a.x_lo = dy_lo; a.w_hi = w_hi; a.w_lo = w_lo; a.bias = nullptr;
    a.out = dx_ptr; a.stats = nullptr; a.p_chunk = 0; a.accum = accumulate ? 1 : 0;
    ProfScope prof(DLIO_PROF_CONV_DGRAD_SIMT, st);
    if ((dx.ph > 0 || dx.pw > 0) && !accumulate) DLIO_CUDA(cudaMemsetAsync(dx_ptr, 0, a.o.numel() * sizeof(float), st));
    long long M = (long long)dx.n * dx.h * dx.w;
    dim3 grid(ceil_div(M, BM), ceil_div(dx.c, BN));
    conv_dgrad_kernel<<<grid, NT, 0, st>>>(a);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_conv2d_bwd_weight(dlio_tensor4 x, const float *x_hi, const float *x_lo, dlio_tensor4 dy,
                                      const float *dy_hi, const float *dy_lo, dlio_conv cv, float *dw,
                                      void *stream) {
    int rc = check_conv(x, dy, cv);
    if (rc) return rc;
    DLIO_CHECK_ARG(x_hi && dy_hi && dw, "conv_bwd_weight: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    ConvArgs a;
    a.x = Geo(x); a.y = Geo(dy); a.o = Geo(dy);
    a.kh = cv.kh; a.kw = cv.kw; a.sh = cv.sh; a.sw = cv.sw; a.ph = cv.ph; a.pw = cv.pw;
    a.cin = x.c; a.cout = dy.c; a.act = 0;
    a.x_hi = x_hi; a.x_lo = x_lo; a.w_hi = dy_hi; a.w_lo = dy_lo; a.bias = nullptr;
    a.out = dw; a.stats = nullptr;
    const int KW = cv.kh * cv.kw * x.c;
    rc = conv_tc_wgrad(a, st);
    if (rc != 0) return rc < 0 ? rc : DLIO_OK;
    ProfScope prof(DLIO_PROF_CONV_WGRAD_SIMT, st);
    DLIO_CUDA(cudaMemsetAsync(dw, 0, (size_t)dy.c * KW * sizeof(float), st));
    long long P = (long long)dy.n * dy.h * dy.w;
    int tiles = ceil_div(dy.c, BM) * ceil_div(KW, BN);
    int split = (4 * 148 + tiles - 1) / tiles;
    long long max_split = (P + 4 * BK - 1) / (4 * BK);
    if (split > max_split) split = (int)max_split;
    if (split < 1) split = 1;
    if (split > 65535) split = 65535;
    long long chunk = (P + split - 1) / split;
    chunk = (chunk + BK - 1) / BK * BK;
    split = (int)((P + chunk - 1) / chunk);
    a.p_chunk = chunk;
    dim3 grid(ceil_div(dy.c, BM), ceil_div(KW, BN), split);
    conv_wgrad_kernel<<<grid, NT, 0, st>>>(a);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_weight_to_ohwi(const float *w_oihw, int cout, int cin, int kh, int kw, int cin_pad,
                                   float *w_hi, float *w_lo, void *stream) {
    DLIO_CHECK_ARG(w_oihw && w_hi && cin_pad >= cin && cout > 0 && cin > 0, "weight_to_ohwi: bad argument");
    long long total = (long long)cout * kh * kw * cin_pad;
    weight_to_ohwi_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, cout, cin, kh * kw, cin_pad, w_hi, w_lo);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
extern "C" int dlio_weight_grad_to_oihw(const float *dw_ohwi, int cout, int cin, int kh, int kw, int cin_pad,
                                        float *dw_oihw, void *stream) {
    DLIO_CHECK_ARG(dw_ohwi && dw_oihw && cin_pad >= cin, "weight_grad_to_oihw: bad argument");
    long long total = (long long)cout * cin * kh * kw;
    weight_grad_to_oihw_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dw_ohwi, cout, cin, kh * kw, cin_pad, dw_oihw);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
extern "C" int dlio_weight_flip_transpose(const float *w_ohwi, int cout, int cin, int kh, int kw, float *wt_hi,
                                          float *wt_lo, void *stream) {
    DLIO_CHECK_ARG(w_ohwi && wt_hi, "weight_flip_transpose: bad argument");
    long long total = (long long)cout * cin * kh * kw;
    weight_flip_transpose_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w_ohwi, cout, cin, kh, kw, wt_hi, wt_lo);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
