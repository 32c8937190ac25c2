// HBM-bound passes around the convolutions: input packing, batch-norm statistics -> scale/shift,
// fused BN-apply (+residual)(+ReLU)(+3x3 max-pool) consumer pass and its two-pass backward, spatial
// reductions (global average pool, SE squeeze), SE channel scaling, small element-wise helpers.
//
// All activations are padded NHWC fp32 (see deeplio_b200.h).  Threads are mapped channel-fastest with one
// float4 (4 channels) per thread, so a warp reads / writes 512 contiguous bytes per pixel group.
// Kernels that reduce over pixels keep a FIXED channel group per thread (block size is a multiple of
// the number of channel groups), accumulate in registers, combine through shared-memory atomics and
// finish with one fp64 global atomic per channel per block.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dlio {

constexpr int MAX_C = 1024;  // channel limit of the reducing kernels (largest on the path: FlowNet conv6)

// threads per block: a multiple of the channel-group count (fixed channel group per thread) and, whenever that
// fits in 256 threads, of the warp size (the reducing kernels use full-warp shuffles)
// DLIO_EW_BLOCK (environment, 64 .. 256, default 256) caps the block size: 128-thread blocks fit beside a resident
// tcgen05 convolution CTA (320 threads x 168 registers, ~205 KB of shared memory), so that with the two encoders on two
// streams (engine.ENC_STREAMS) these HBM-bound passes run under the other encoder's tensor-bound convolutions
static int g_ew_cap = 0;
static int ew_block_cap() {
    int &cap = g_ew_cap;
    if (!cap) {
        const char *e = getenv("DLIO_EW_BLOCK");
        int v = e ? atoi(e) : 256;
        cap = v < 64 ? 64 : (v > 256 ? 256 : v);
    }
    return cap;
}
// DLIO_POOL_TMA=0 falls back to the per-thread-load pooling kernels (A/B switch for bench and tests)
extern int g_conv_cg2;      // conv_tc.cu
extern int g_nvtx;          // lib.cu
extern int g_bwd_single;    // conv_tc.cu
int g_apply_rows = 1;       // option "apply_rows": row-structured non-pooled BN-apply pass (0: the generic kernel)
static int g_pool_tma = -1;
static bool pool_tma_enabled() {
    if (g_pool_tma < 0) {
        const char *e = getenv("DLIO_POOL_TMA");
        g_pool_tma = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pool_tma != 0;
}
static inline int block_for_cg(int cg) {
    const int cap = ew_block_cap();
    int g = cg, b = 32;
    while (b) { int t = g % b; g = b; b = t; }          // g = gcd(cg, 32)
    const int unit = cg * (32 / g);
    if (unit <= cap) return unit * (cap / unit);
    if (unit <= 256) return unit;
    return cg * (256 / cg > 0 ? 256 / cg : 1);
}
// persistent grid of a row-structured kernel: exactly the number of blocks that are resident at once
template <typename K>
static int resident_grid(K kernel, int block) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return 148 * per_sm;
}
// column segments per row such that a persistent grid gets >= 8 work items per block (load balance), while a
// segment keeps at least two pixels per w-lane
static inline int row_segments(int rows, int width, int wlanes, int grid) {
    int segs = (8 * grid + rows - 1) / rows;
    const int max_segs = width / (2 * wlanes) > 0 ? width / (2 * wlanes) : 1;
    if (segs > max_segs) segs = max_segs;
    return segs < 1 ? 1 : segs;
}
// blocks per SM of the reduction passes that end in 2 C fp64 atomics per block: at C = 512 a grid of 8 blocks per SM
// sends 1.2 M atomics to 1024 addresses (~20 us of every launch, whatever the tensor size); two blocks per SM with
// eight 16-byte loads in flight per thread still cover the HBM latency
static inline int reduce_blocks_per_sm(int c) { return c >= 256 ? 2 : (c >= 128 ? 4 : 8); }
static inline int grid_for(long long total, int block, int per_sm = 8) {
    long long g = (total + block - 1) / block;
    long long cap = 148LL * per_sm;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// Pixel loop with a FIXED channel group per thread (blockDim.x is a multiple of cg): 32-bit index arithmetic
// only -- the first version of these kernels decoded a 64-bit flat index with five 64-bit divisions per
// element and was instruction-bound (ncu: 780 instructions per element, 14 % of HBM bandwidth).
#define DLIO_PIX_LOOP(pix, npix, cg)                                                              \
    for (unsigned pix = (blockIdx.x * blockDim.x + threadIdx.x) / (unsigned)(cg),                 \
                  pix##_step = (gridDim.x * blockDim.x) / (unsigned)(cg);                         \
         pix < (unsigned)(npix); pix += pix##_step)

__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float4 fma4(const float4 &a, const float4 &b, const float4 &c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 mul4(const float4 &a, const float4 &b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 relu4(const float4 &a) {
    return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}

// ------------------------------------------------------------------ input packing
__global__ void pack_input_kernel(const float *__restrict__ src, long long sn, long long st, long long sc, int T,
                                  int C, Geo d, float *__restrict__ dst, float *__restrict__ dst_lo) {
    const unsigned total = (unsigned)d.n * d.hp * d.wp;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int xx = (int)(i % (unsigned)d.wp);
        unsigned t = i / (unsigned)d.wp;
        int yy = (int)(t % (unsigned)d.hp);
        int n = (int)(t / (unsigned)d.hp);
        int h = yy - d.ph, w = xx - d.pw;
        bool in = h >= 0 && h < d.h && w >= 0 && w < d.w;
        for (int c0 = 0; c0 < d.c; c0 += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = c0 + j;
                float x = 0.f;
                if (in && c < T * C) {
                    int tt = c / C, cc = c - tt * C;
                    x = src[(size_t)n * sn + (size_t)tt * st + (size_t)cc * sc + (size_t)h * d.w + w];
                }
                v[j] = x;
            }
            st4_split(dst, dst_lo, (size_t)i * d.c + c0, make_float4(v[0], v[1], v[2], v[3]));
        }
    }
}

// ------------------------------------------------------------------ BN statistics -> scale / shift
// One block.  out_bound (optional): an upper bound of |scale*y + shift| (+ |residual|) over the tensor, from the
// BATCH statistics of y whatever statistics define scale/shift:  scale*y + shift = scale*(y - mu_b) + (scale*mu_b +
// shift) and |y - mu_b| <= sqrt(count * var_b) for every one of the `count` samples.  ReLU and max-pooling only
// shrink it.  It defines the power-of-two scale of the fp16 split planes the apply pass writes.
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const double *__restrict__ stats, double count, int c,
                                   const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float *running_mean, float *running_var, float momentum, float eps,
                                   int use_running, float *mean, float *invstd, float *scale, float *shift,
                                   const float *res_bound, float *out_bound, long long *num_batches_tracked) {
    __shared__ float red[32];
    float bmax = 0.f;
    // momentum < 0: nn.BatchNorm2d(momentum=None), the cumulative moving average with factor 1 / num_batches_tracked
    // (after this step's increment).  Every thread reads the counter before thread 0 bumps it.
    if (num_batches_tracked && !use_running) {
        const long long seen = *num_batches_tracked + 1;
        if (momentum < 0.f) momentum = 1.f / (float)seen;
        __syncthreads();
        if (threadIdx.x == 0) *num_batches_tracked = seen;
    } else if (momentum < 0.f) {
        momentum = 1.f;
    }
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        float m, is;
        double mu = 0.0, var = 0.0;
        if (stats) {
            mu = stats[i] / count;
            var = stats[c + i] / count - mu * mu;
            if (var < 0.0) var = 0.0;
        }
        if (use_running) {
            m = running_mean[i];
            is = 1.f / sqrtf(running_var[i] + eps);
        } else {
            m = (float)mu;
            is = (float)(1.0 / sqrt(var + (double)eps));
            if (running_mean) {
                double unb = count > 1.0 ? var * count / (count - 1.0) : var;
                running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * m;
                running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unb;
            }
        }
        float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
        const float sc = g * is, sf = b - m * g * is;
        mean[i] = m;
        invstd[i] = is;
        scale[i] = sc;
        shift[i] = sf;
        if (out_bound) bmax = fmaxf(bmax, (float)(fabs((double)sc) * sqrt(count * var) + fabs((double)sc * mu + (double)sf)));
    }
    if (out_bound) {
        bmax = warp_max(bmax);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bmax;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
            v = warp_max(v);
            if (threadIdx.x == 0) *out_bound = v * 1.001f + (res_bound ? *res_bound : 0.f);
        }
    }
}

// ------------------------------------------------------------------ fused BN-apply / residual / ReLU / max-pool
struct BnPool {
    Geo y, out, res;
    const float *yp, *scale, *shift, *resp;
    int res_mode;  // 0 none, 1 added before the activation, 2 added after it
    int relu, pk, sh, sw, c_off, cg;
    int segs;                   // row-structured kernels: column segments per row
    float *out_hi, *out_lo;     // fp32 plane (may be NULL when only the fp16 planes are wanted), TF32 lo plane
    __half *out_h2;             // packed fp16 hi|lo planes (optional) and the bound that defines their scale
    const float *out_bound;
    int group;                  // layout of out_h2: 1 plain, 2 pixel pairs (common.cuh, st4_h2)
    uint8_t *idx;
    float *ymax;                // optional [n, ho, wo, C]: the conv output y at the arg-max (what the BN backward
                                // sums need from the input side: dlio_pool_bwd_sums)
    int zero_tail;              // generic kernel: also write zeros to the output channels [y.c, out.c)
};
__device__ __forceinline__ void bnpool_store(const BnPool &a, unsigned pix, int c, const float4 &v, float s16) {
    const size_t o = (size_t)pix * a.out.c + a.c_off + c;
    if (a.out_hi) st4_split(a.out_hi, a.out_lo, o, v);
    if (a.out_h2) st4_h2(a.out_h2, pix, a.out.c, a.c_off + c, v, s16, a.group);
}

// channels [y.c, out.c) of one output pixel, shared among the pixel's threads (thread of channel c: y.c + c, 2 y.c + c ..)
__device__ __forceinline__ void bnpool_zero_tail(const BnPool &a, unsigned pix, int c) {
    for (int cz = a.y.c + c; cz < a.out.c; cz += a.y.c) {
        if (a.out_hi) st4_split(a.out_hi, a.out_lo, (size_t)pix * a.out.c + cz, f4(0.f));
        if (a.out_h2) st4_h2_zero(a.out_h2, pix, a.out.c, cz, a.group);
    }
}

__device__ __forceinline__ float4 bnpool_value(const BnPool &a, int n, int h, int w, int c, const float4 &sc,
                                               const float4 &sf) {
    float4 v = ld4(a.yp + a.y.off(n, h, w) + c);
    if (a.scale) v = fma4(sc, v, sf);
    if (a.res_mode == 1) v = add4(v, ld4(a.resp + a.res.off(n, h, w) + a.c_off + c));
    if (a.relu) v = relu4(v);
    if (a.res_mode == 2) v = add4(v, ld4(a.resp + a.res.off(n, h, w) + a.c_off + c));
    return v;
}

__global__ void __launch_bounds__(256) bn_act_pool_fwd_kernel(BnPool a) {
    const int c = (int)(threadIdx.x % a.cg) * 4;
    const float s16 = a.out_h2 ? f16_scale_from_bound(*a.out_bound) : 1.f;
    DLIO_PIX_LOOP(pix, a.out.n * a.out.hp * a.out.wp, a.cg) {
        int xx = (int)(pix % (unsigned)a.out.wp);
        unsigned t = pix / (unsigned)a.out.wp;
        int yy = (int)(t % (unsigned)a.out.hp);
        int n = (int)(t / (unsigned)a.out.hp);
        int ho = yy - a.out.ph, wo = xx - a.out.pw;
        if (a.zero_tail) bnpool_zero_tail(a, pix, c);
        if (ho < 0 || ho >= a.out.h || wo < 0 || wo >= a.out.w) {
            bnpool_store(a, pix, c, f4(0.f), 1.f);
            continue;
        }
        float4 sc = f4(1.f), sf = f4(0.f);
        if (a.scale) {
            sc = ld4(a.scale + c);
            sf = ld4(a.shift + c);
        }
        float4 best;
        if (a.pk == 1) {
            best = bnpool_value(a, n, ho, wo, c, sc, sf);
        } else {
            best = f4(-FLT_MAX);
            uchar4 bi = make_uchar4(0, 0, 0, 0);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                int h = ho * a.sh - 1 + dy;
                if (h < 0 || h >= a.y.h) continue;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    int w = wo * a.sw - 1 + dx;
                    if (w < 0 || w >= a.y.w) continue;
                    float4 v = bnpool_value(a, n, h, w, c, sc, sf);
                    unsigned char r = (unsigned char)(dy * 3 + dx);
                    if (v.x > best.x) { best.x = v.x; bi.x = r; }
                    if (v.y > best.y) { best.y = v.y; bi.y = r; }
                    if (v.z > best.z) { best.z = v.z; bi.z = r; }
                    if (v.w > best.w) { best.w = v.w; bi.w = r; }
                }
            }
            if (a.idx)
                *reinterpret_cast<uchar4 *>(a.idx + (((size_t)n * a.out.h + ho) * a.out.w + wo) * a.y.c + c) = bi;
        }
        bnpool_store(a, pix, c, best, s16);
    }
}

// Row-structured variant of the pass above for the layers that carry almost all of its traffic: 3x3 max-pool with
// compile-time strides and no residual.  One block walks whole output rows (grid-stride over n * padded rows), a
// thread keeps its channel group and strides along w, so the per-element work is loads, FMAs and the arg-max
// tracking -- the generic kernel spent 70 % of its 600 instructions per output on 64-bit index arithmetic (two
// divisions per pixel, nine bounds checks and address computations) and ran at 21 % of the HBM bandwidth.
template <int SH, int SW, bool RELU>
__global__ void __launch_bounds__(256, 5) bn_pool3_fwd_kernel(BnPool a) {
    const int cg = a.cg, C = cg * 4;
    const int c = (int)(threadIdx.x % cg) * 4;
    const int wl = (int)(threadIdx.x / cg), WL = (int)(blockDim.x / cg);
    const float s16 = a.out_h2 ? f16_scale_from_bound(*a.out_bound) : 1.f;
    float4 sc = f4(1.f), sf = f4(0.f);
    if (a.scale) {
        sc = ld4(a.scale + c);
        sf = ld4(a.shift + c);
    }
    // work items: (output row, column segment); a persistent grid strides over them so that every SM stays busy
    // whatever the number of rows
    const int rows = a.out.n * a.out.hp, segs = a.segs, items = rows * segs;
    const int seg_w = (a.out.wp + segs - 1) / segs;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int row = item / segs, seg = item - row * segs;
        const int x_begin = seg * seg_w, x_end = min(a.out.wp, x_begin + seg_w);
        const int n = row / a.out.hp, ho = row - n * a.out.hp - a.out.ph;
        const unsigned pix0 = (unsigned)row * (unsigned)a.out.wp;          // padded output pixel index of column 0
        if (ho < 0 || ho >= a.out.h) {
            for (int x = x_begin + wl; x < x_end; x += WL) bnpool_store(a, pix0 + x, c, f4(0.f), 1.f);
            continue;
        }
        // the three input rows of this output row (block-uniform validity)
        const float *rp[3];
        bool rv[3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int h = ho * SH - 1 + dy;
            rv[dy] = h >= 0 && h < a.y.h;
            rp[dy] = a.yp + a.y.off(n, rv[dy] ? h : 0, 0) + c;
        }
        const bool rows_ok = rv[0] && rv[1] && rv[2];
        uint8_t *irow = a.idx ? a.idx + ((size_t)(n * a.out.h + ho) * a.out.w) * C + c : nullptr;
        for (int x = x_begin + wl; x < x_end; x += WL) {
            const int wo = x - a.out.pw;
            if (wo < 0 || wo >= a.out.w) {
                bnpool_store(a, pix0 + x, c, f4(0.f), 1.f);
                continue;
            }
            const int w0 = wo * SW - 1;
            float4 best = f4(-FLT_MAX), yb = f4(0.f);          // yb: raw conv output at the arg-max
            unsigned bi = 0;                                   // four packed window-relative arg-max bytes
#define DLIO_POOL_TAKE(v, r, raw)                                                                   \
    do {                                                                                            \
        if (v.x > best.x) { best.x = v.x; yb.x = raw.x; bi = (bi & 0xFFFFFF00u) | (r); }            \
        if (v.y > best.y) { best.y = v.y; yb.y = raw.y; bi = (bi & 0xFFFF00FFu) | ((r) << 8); }     \
        if (v.z > best.z) { best.z = v.z; yb.z = raw.z; bi = (bi & 0xFF00FFFFu) | ((r) << 16); }    \
        if (v.w > best.w) { best.w = v.w; yb.w = raw.w; bi = (bi & 0x00FFFFFFu) | ((r) << 24); }    \
    } while (0)
            if (rows_ok && w0 >= 0 && w0 + 2 < a.y.w) {
                // interior window: nine independent loads issued back to back (the kernel is bound by load latency)
                float4 v[9];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) v[dy * 3 + dx] = ld4(rp[dy] + (w0 + dx) * C);
#pragma unroll
                for (int r = 0; r < 9; ++r) {
                    float4 t = fma4(sc, v[r], sf);
                    if (RELU) t = relu4(t);
                    DLIO_POOL_TAKE(t, (unsigned)r, v[r]);
                }
            } else {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    if (!rv[dy]) continue;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int w = w0 + dx;
                        if (w < 0 || w >= a.y.w) continue;
                        const float4 raw = ld4(rp[dy] + w * C);
                        float4 t = fma4(sc, raw, sf);
                        if (RELU) t = relu4(t);
                        DLIO_POOL_TAKE(t, (unsigned)(dy * 3 + dx), raw);
                    }
                }
            }
#undef DLIO_POOL_TAKE
            if (irow) *reinterpret_cast<unsigned *>(irow + (size_t)wo * C) = bi;
            if (a.ymax) st4(a.ymax + ((size_t)(n * a.out.h + ho) * a.out.w + wo) * C + c, yb);
            bnpool_store(a, pix0 + x, c, best, s16);
        }
    }
}

// Row-structured variant of bn_act_pool_fwd_kernel for the layers WITHOUT pooling (every layer of FlowNet / ResNet,
// every Fire convolution of PointSeg, Simple-1 conv3 / 5): a block walks whole padded output rows, a thread keeps its
// channel group and strides along the row with four pixels of loads in flight -- no per-pixel divisions, scale / shift
// loaded once.  The generic kernel ran these at 2.7 - 3.2 TB/s with 60 % of the issue slots busy on index arithmetic.
__global__ void __launch_bounds__(256, 3) bn_apply_rows_kernel(BnPool a) {
    const int cg = a.cg, C = cg * 4;
    const int c = (int)(threadIdx.x % cg) * 4;
    const int wl = (int)(threadIdx.x / cg), WL = (int)(blockDim.x / cg);
    const float s16 = a.out_h2 ? f16_scale_from_bound(*a.out_bound) : 1.f;
    float4 sc = f4(1.f), sf = f4(0.f);
    if (a.scale) {
        sc = ld4(a.scale + c);
        sf = ld4(a.shift + c);
    }
    const int rows = a.out.n * a.out.hp, segs = a.segs, items = rows * segs;
    const int seg_w = (a.out.wp + segs - 1) / segs;
    constexpr int U = 4;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int row = item / segs, seg = item - row * segs;
        const int x_begin = seg * seg_w, x_end = min(a.out.wp, x_begin + seg_w);
        const int n = row / a.out.hp, ho = row - n * a.out.hp - a.out.ph;
        const unsigned pix0 = (unsigned)row * (unsigned)a.out.wp;
        if (ho < 0 || ho >= a.out.h) {
            for (int x = x_begin + wl; x < x_end; x += WL) {
                bnpool_store(a, pix0 + x, c, f4(0.f), 1.f);
                if (a.zero_tail) bnpool_zero_tail(a, pix0 + x, c);
            }
            continue;
        }
        const float *yrow = a.yp + a.y.off(n, ho, 0) + c;
        const float *rrow = a.res_mode ? a.resp + a.res.off(n, ho, 0) + a.c_off + c : nullptr;
        const int rc = a.res.c;
        for (int x0 = x_begin + wl; x0 < x_end; x0 += U * WL) {
            float4 v[U], r[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = x0 + u * WL, wo = x - a.out.pw;
                ok[u] = x < x_end && wo >= 0 && wo < a.out.w;
                v[u] = ok[u] ? ld4(yrow + (size_t)wo * C) : f4(0.f);
                r[u] = (ok[u] && rrow) ? ld4(rrow + (size_t)wo * rc) : f4(0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = x0 + u * WL;
                if (x >= x_end) break;
                float4 t = f4(0.f);
                if (ok[u]) {
                    t = a.scale ? fma4(sc, v[u], sf) : v[u];
                    if (a.res_mode == 1) t = add4(t, r[u]);
                    if (a.relu) t = relu4(t);
                    if (a.res_mode == 2) t = add4(t, r[u]);
                }
                bnpool_store(a, pix0 + x, c, t, ok[u] ? s16 : 1.f);
                if (a.zero_tail) bnpool_zero_tail(a, pix0 + x, c);
            }
        }
    }
}

// backward pass 1: dz (gradient at the BN output position, after un-pooling and the ReLU mask) + sums
struct BnBwd {
    Geo y, res, dout;
    const float *yp, *scale, *shift, *mean, *invstd, *resp, *doutp;
    int res_mode, relu, pk, sh, sw, c_off, cg;
    int grad_src, ld_dout;
    int pooled_h, pooled_w;
    const uint8_t *idx;
    float *dz;
    float *dres;
    int dres_c, dres_acc;
    double *sums;
    int want_absmax;   // sums[2C] receives max |dz| (bit pattern of a non-negative double, atomicMax)
    int segs;          // row-structured kernel: column segments per row
};

// Gradient of a 3x3 max-pool (pad 1, stride SH x SW) gathered at input position (h, w): the sum of dout over
// the windows that contain (h, w) and whose arg-max is (h, w).  Strides are compile-time so the candidate set
// (3 windows along a stride-1 axis, 1 or 2 along a stride-2 axis) and the window-relative position fold into
// constants; all candidate loads are issued unconditionally on clamped addresses and masked afterwards.
template <int SH, int SW>
__device__ __forceinline__ float4 unpool_gather(const BnBwd &a, int n, int h, int w, int c, int C) {
    constexpr int NR = SH == 1 ? 3 : 2, NC = SW == 1 ? 3 : 2;
    const int ho0 = SH == 1 ? h - 1 : h >> 1, wo0 = SW == 1 ? w - 1 : w >> 1;
    float4 g = f4(0.f);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const int ho = ho0 + i;
        const bool okh = ho >= 0 && ho < a.pooled_h && (SH == 1 || i == 0 || (h & 1));
        const int rh = SH == 1 ? 2 - i : h + 1 - 2 * ho;
        const int hoc = okh ? ho : 0;
        const uint8_t *irow = a.idx + ((size_t)(n * a.pooled_h + hoc) * a.pooled_w) * C + c;
        const float *drow = a.doutp + a.dout.off(n, hoc, 0) + a.c_off + c;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int wo = wo0 + j;
            const bool ok = okh && wo >= 0 && wo < a.pooled_w && (SW == 1 || j == 0 || (w & 1));
            const int rw = SW == 1 ? 2 - j : w + 1 - 2 * wo;
            const int woc = ok ? wo : 0;
            const unsigned bi = *reinterpret_cast<const unsigned *>(irow + woc * C);
            const float4 d = ld4(drow + woc * a.dout.c);
            const unsigned r = ok ? (unsigned)(rh * 3 + rw) : 255u;
            const unsigned m = __vcmpeq4(bi, r * 0x01010101u);   // 0xFF in every byte whose arg-max is (h, w)
            if (m & 0x000000FFu) g.x += d.x;
            if (m & 0x0000FF00u) g.y += d.y;
            if (m & 0x00FF0000u) g.z += d.z;
            if (m & 0xFF000000u) g.w += d.w;
        }
    }
    return g;
}
template <>
__device__ __forceinline__ float4 unpool_gather<0, 0>(const BnBwd &, int, int, int, int, int) { return f4(0.f); }

// SH = SW = 0: no pooling (direct or global-average gradient source)
template <int SH, int SW>
__global__ void __launch_bounds__(256) bn_act_pool_bwd_reduce_kernel(BnBwd a) {
    __shared__ float red[2 * MAX_C];
    const int C = a.cg * 4;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int c = (int)(threadIdx.x % a.cg) * 4;  // fixed per thread: blockDim.x % cg == 0
    float4 sc = f4(1.f), sf = f4(0.f), mu = f4(0.f), is = f4(1.f);
    if (a.scale) {
        sc = ld4(a.scale + c);
        sf = ld4(a.shift + c);
    }
    if (a.mean) {
        mu = ld4(a.mean + c);
        is = ld4(a.invstd + c);
    }
    const float inv_hw = 1.f / (float)(a.y.h * a.y.w);
    float4 s1 = f4(0.f), s2 = f4(0.f);
    float amax = 0.f;
    DLIO_PIX_LOOP(pix, a.y.n * a.y.h * a.y.w, a.cg) {
        int w = (int)(pix % (unsigned)a.y.w);
        unsigned t = pix / (unsigned)a.y.w;
        int h = (int)(t % (unsigned)a.y.h);
        int n = (int)(t / (unsigned)a.y.h);
        float4 g = f4(0.f);
        if (a.grad_src == DLIO_GRAD_AVG) {
            g = ld4(a.doutp + (size_t)n * a.ld_dout + a.c_off + c);
            g = make_float4(g.x * inv_hw, g.y * inv_hw, g.z * inv_hw, g.w * inv_hw);
        } else if (a.pk == 1) {
            g = ld4(a.doutp + a.dout.off(n, h, w) + a.c_off + c);
        } else {
            g = unpool_gather<SH, SW>(a, n, h, w, c, C);
        }
        const size_t yo = a.y.off(n, h, w) + c;
        float4 y = ld4(a.yp + yo);
        size_t ro = (size_t)pix * a.dres_c + a.c_off + c;
        if (a.res_mode == 2 && a.dres) st4(a.dres + ro, a.dres_acc ? add4(ld4(a.dres + ro), g) : g);
        if (a.relu) {
            float4 v = a.scale ? fma4(sc, y, sf) : y;
            if (a.res_mode == 1) v = add4(v, ld4(a.resp + a.res.off(n, h, w) + a.c_off + c));
            if (!(v.x > 0.f)) g.x = 0.f;
            if (!(v.y > 0.f)) g.y = 0.f;
            if (!(v.z > 0.f)) g.z = 0.f;
            if (!(v.w > 0.f)) g.w = 0.f;
        }
        if (a.res_mode == 1 && a.dres) st4(a.dres + ro, a.dres_acc ? add4(ld4(a.dres + ro), g) : g);
        if (a.dz) st4(a.dz + (size_t)pix * C + c, g);
        if (a.sums) {
            float4 yh = make_float4((y.x - mu.x) * is.x, (y.y - mu.y) * is.y, (y.z - mu.z) * is.z, (y.w - mu.w) * is.w);
            s1 = add4(s1, g);
            s2 = fma4(g, yh, s2);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w))));
        }
    }
    if (a.sums && a.want_absmax) {
        amax = warp_max(amax);
        if ((threadIdx.x & 31) == 0)
            atomicMax(reinterpret_cast<unsigned long long *>(a.sums + 2 * C), (unsigned long long)__double_as_longlong((double)amax));
    }
    if (a.sums) {
        atomicAdd(&red[c + 0], s1.x); atomicAdd(&red[c + 1], s1.y);
        atomicAdd(&red[c + 2], s1.z); atomicAdd(&red[c + 3], s1.w);
        atomicAdd(&red[C + c + 0], s2.x); atomicAdd(&red[C + c + 1], s2.y);
        atomicAdd(&red[C + c + 2], s2.z); atomicAdd(&red[C + c + 3], s2.w);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(a.sums + i, (double)red[i]);
    }
}

// Backward pass 1 without pooling on flat tensors (y and dout unpadded, no pre-activation residual: every Fire
// convolution of PointSeg, the plain layers of FlowNet, Simple-1 conv3 / 5 / 7): pixel p of every tensor is at p * stride,
// so there is no index arithmetic at all, and four pixels of loads are in flight per thread (the generic kernel issues
// one pixel per iteration behind two divisions: 2.7 - 3.1 TB/s at 35 - 44 % of the issue slots).
__global__ void __launch_bounds__(256) bn_bwd_reduce_flat_kernel(BnBwd a) {
    __shared__ float red[2 * MAX_C];
    const int C = a.cg * 4;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int c = (int)(threadIdx.x % a.cg) * 4;
    float4 sc = f4(1.f), sf = f4(0.f), mu = f4(0.f), is = f4(1.f);
    if (a.scale) {
        sc = ld4(a.scale + c);
        sf = ld4(a.shift + c);
    }
    if (a.mean) {
        mu = ld4(a.mean + c);
        is = ld4(a.invstd + c);
    }
    float4 s1 = f4(0.f), s2 = f4(0.f);
    float amax = 0.f;
    const unsigned npix = (unsigned)(a.y.n * a.y.h * a.y.w);
    constexpr int U = 4;
    const unsigned p0 = (blockIdx.x * blockDim.x + threadIdx.x) / (unsigned)a.cg;
    const unsigned step = (gridDim.x * blockDim.x) / (unsigned)a.cg;
    const float *dbase = a.doutp + a.c_off + c;
    const size_t dstride = (size_t)a.dout.c;
    for (unsigned pix = p0; pix < npix; pix += U * step) {
        float4 g[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned p = pix + u * step;
            const bool ok = p < npix;
            g[u] = ok ? ld4(dbase + (size_t)p * dstride) : f4(0.f);
            y[u] = ok ? ld4(a.yp + (size_t)p * C + c) : mu;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned p = pix + u * step;
            if (p >= npix) break;
            float4 gg = g[u];
            if (a.res_mode == 2 && a.dres) {
                const size_t ro = (size_t)p * a.dres_c + a.c_off + c;
                st4(a.dres + ro, a.dres_acc ? add4(ld4(a.dres + ro), gg) : gg);
            }
            if (a.relu) {
                const float4 v = a.scale ? fma4(sc, y[u], sf) : y[u];
                if (!(v.x > 0.f)) gg.x = 0.f;
                if (!(v.y > 0.f)) gg.y = 0.f;
                if (!(v.z > 0.f)) gg.z = 0.f;
                if (!(v.w > 0.f)) gg.w = 0.f;
            }
            if (a.dz) st4(a.dz + (size_t)p * C + c, gg);
            if (a.sums) {
                const float4 yh = make_float4((y[u].x - mu.x) * is.x, (y[u].y - mu.y) * is.y, (y[u].z - mu.z) * is.z,
                                              (y[u].w - mu.w) * is.w);
                s1 = add4(s1, gg);
                s2 = fma4(gg, yh, s2);
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(gg.x), fabsf(gg.y)), fmaxf(fabsf(gg.z), fabsf(gg.w))));
            }
        }
    }
    if (a.sums && a.want_absmax) {
        amax = warp_max(amax);
        if ((threadIdx.x & 31) == 0)
            atomicMax(reinterpret_cast<unsigned long long *>(a.sums + 2 * C), (unsigned long long)__double_as_longlong((double)amax));
    }
    if (a.sums) {
        atomicAdd(&red[c + 0], s1.x); atomicAdd(&red[c + 1], s1.y);
        atomicAdd(&red[c + 2], s1.z); atomicAdd(&red[c + 3], s1.w);
        atomicAdd(&red[C + c + 0], s2.x); atomicAdd(&red[C + c + 1], s2.y);
        atomicAdd(&red[C + c + 2], s2.z); atomicAdd(&red[C + c + 3], s2.w);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(a.sums + i, (double)red[i]);
    }
}

// Row-structured variant of backward pass 1 for pooled layers without a residual (the bulk of its traffic): a block
// walks whole input rows, so the candidate pooling-window rows, their arg-max / gradient row pointers and the
// window-relative row offsets are computed once per row instead of per element (see bn_pool3_fwd_kernel).
template <int SH, int SW>
__global__ void __launch_bounds__(256, 4) bn_pool3_bwd_reduce_kernel(BnBwd a) {
    __shared__ float red[2 * MAX_C];
    constexpr int NR = SH == 1 ? 3 : 2, NC = SW == 1 ? 3 : 2;
    const int cg = a.cg, C = cg * 4;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int c = (int)(threadIdx.x % cg) * 4;
    const int wl = (int)(threadIdx.x / cg), WL = (int)(blockDim.x / cg);
    float4 sc = f4(1.f), sf = f4(0.f), mu = f4(0.f), is = f4(1.f);
    if (a.scale) {
        sc = ld4(a.scale + c);
        sf = ld4(a.shift + c);
    }
    if (a.mean) {
        mu = ld4(a.mean + c);
        is = ld4(a.invstd + c);
    }
    float4 s1 = f4(0.f), s2 = f4(0.f);
    float amax = 0.f;
    const int rows = a.y.n * a.y.h, segs = a.segs, items = rows * segs;
    const int seg_w = (a.y.w + segs - 1) / segs;
    const int dstride = a.dout.c;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int row = item / segs, seg = item - row * segs;
        const int w_begin = seg * seg_w, w_end = min(a.y.w, w_begin + seg_w);
        const int n = row / a.y.h, h = row - n * a.y.h;
        const int ho0 = SH == 1 ? h - 1 : h >> 1;
        const uint8_t *ip[NR];
        const float *dp[NR];
        unsigned rbase[NR];         // (window-relative row) * 3, or 255 when the window row does not exist
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int ho = ho0 + i;
            const bool okh = ho >= 0 && ho < a.pooled_h && (SH == 1 || i == 0 || (h & 1));
            const int rh = SH == 1 ? 2 - i : h + 1 - 2 * ho;
            const int hoc = okh ? ho : 0;
            ip[i] = a.idx + ((size_t)(n * a.pooled_h + hoc) * a.pooled_w) * C + c;
            dp[i] = a.doutp + a.dout.off(n, hoc, 0) + a.c_off + c;
            rbase[i] = okh ? (unsigned)(rh * 3) : 255u;
        }
        const float *yrow = a.yp + a.y.off(n, h, 0) + c;
        float *dzrow = a.dz + ((size_t)row * a.y.w) * C + c;
        for (int w = w_begin + wl; w < w_end; w += WL) {
            const int wo0 = SW == 1 ? w - 1 : w >> 1;
            float4 g = f4(0.f);
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const int wo = wo0 + j;
                const bool okw = wo >= 0 && wo < a.pooled_w && (SW == 1 || j == 0 || (w & 1));
                const int rw = SW == 1 ? 2 - j : w + 1 - 2 * wo;
                const int woc = okw ? wo : 0;
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const unsigned bi = *reinterpret_cast<const unsigned *>(ip[i] + woc * C);
                    const float4 d = ld4(dp[i] + woc * dstride);
                    const unsigned r = (okw && rbase[i] != 255u) ? rbase[i] + (unsigned)rw : 255u;
                    const unsigned m = __vcmpeq4(bi, r * 0x01010101u);   // 0xFF in every byte whose arg-max is (h, w)
                    if (m & 0x000000FFu) g.x += d.x;
                    if (m & 0x0000FF00u) g.y += d.y;
                    if (m & 0x00FF0000u) g.z += d.z;
                    if (m & 0xFF000000u) g.w += d.w;
                }
            }
            const float4 y = ld4(yrow + w * C);
            if (a.relu) {
                const float4 v = a.scale ? fma4(sc, y, sf) : y;
                if (!(v.x > 0.f)) g.x = 0.f;
                if (!(v.y > 0.f)) g.y = 0.f;
                if (!(v.z > 0.f)) g.z = 0.f;
                if (!(v.w > 0.f)) g.w = 0.f;
            }
            st4(dzrow + (size_t)w * C, g);
            const float4 yh = make_float4((y.x - mu.x) * is.x, (y.y - mu.y) * is.y, (y.z - mu.z) * is.z, (y.w - mu.w) * is.w);
            s1 = add4(s1, g);
            s2 = fma4(g, yh, s2);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w))));
        }
    }
    if (a.sums && a.want_absmax) {
        amax = warp_max(amax);
        if ((threadIdx.x & 31) == 0)
            atomicMax(reinterpret_cast<unsigned long long *>(a.sums + 2 * C), (unsigned long long)__double_as_longlong((double)amax));
    }
    if (a.sums) {
        atomicAdd(&red[c + 0], s1.x); atomicAdd(&red[c + 1], s1.y);
        atomicAdd(&red[c + 2], s1.z); atomicAdd(&red[c + 3], s1.w);
        atomicAdd(&red[C + c + 0], s2.x); atomicAdd(&red[C + c + 1], s2.y);
        atomicAdd(&red[C + c + 2], s2.z); atomicAdd(&red[C + c + 3], s2.w);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(a.sums + i, (double)red[i]);
    }
}

// backward pass 2: dy = scale * (dz - mean(dz) - yhat * mean(dz * yhat)) [* (y > 0)]
struct BnApply {
    Geo y, dy;
    const float *yp, *dz, *scale, *shift, *mean, *invstd;
    const double *sums;
    double count;
    int pre_relu, batch_stats, cg, segs;
    int post_relu;     // dz is the gradient BEFORE the ReLU that follows the BN: masked here by scale*y + shift > 0
    // gather variant (template SH, SW > 0): dz is not read from memory but un-pooled on the fly from the gradient of
    // the pooled output and the arg-max bytes, as backward pass 1 does (no dz round trip through HBM)
    Geo dout;
    const float *doutp;
    const uint8_t *idx;
    int pooled_h, pooled_w, c_off;
    int hs;            // dy row step: y row i lies on dy row hs * i, the other dy rows are zero (H-strided convolution
                       // whose backward runs as a stride-1 convolution over the input grid)
    float *dy_hi, *dy_lo, *dgamma, *dbeta;
    double *dbias;
    __half *dy_h2;     // packed fp16 hi|lo planes of dy (optional); needs sums[2C] = max |dz|
    float *dy_bound;   // written by block 0: the bound that defines their scale
};

// (the gather variant is latency-bound on its candidate loads: registers capped for four resident blocks per SM)
template <int SH, int SW>
__global__ void __launch_bounds__(256, SH > 0 ? 4 : 1) bn_bwd_apply_kernel(BnApply a) {
    __shared__ float red[MAX_C];
    __shared__ float wred[32];
    const int C = a.cg * 4;
    float s16 = 1.f;
    if (a.dy_h2) {
        // |dy| <= |scale| * (max|dz| + |mean dz| + sqrt(count) * |mean(dz * yhat)|)  (|yhat| <= sqrt(count));
        // every block derives the same bound (C <= 1024 values), block 0 publishes it for dgrad / wgrad
        const float mz = (float)a.sums[2 * C];
        float b = 0.f;
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            float t = mz;
            if (a.batch_stats) t += (float)(fabs(a.sums[i]) / a.count + sqrt(a.count) * fabs(a.sums[C + i]) / a.count);
            b = fmaxf(b, fabsf(a.scale ? a.scale[i] : 1.f) * t);
        }
        b = warp_max(b);
        if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = b;
        __syncthreads();
        b = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) b = fmaxf(b, wred[i]);
        b *= 1.001f;
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.dy_bound = b;
        s16 = f16_scale_from_bound(b);
    }
    if (a.dbias) {
        for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
        __syncthreads();
    }
    const int c = (int)(threadIdx.x % a.cg) * 4;
    float4 sc = f4(1.f), sf = f4(0.f), mu = f4(0.f), is = f4(1.f), m1 = f4(0.f), m2 = f4(0.f);
    if (a.scale) sc = ld4(a.scale + c);
    if (a.post_relu && a.shift) sf = ld4(a.shift + c);
    if (a.mean) {
        mu = ld4(a.mean + c);
        is = ld4(a.invstd + c);
    }
    if (a.sums && a.batch_stats) {
        m1 = make_float4((float)(a.sums[c] / a.count), (float)(a.sums[c + 1] / a.count),
                         (float)(a.sums[c + 2] / a.count), (float)(a.sums[c + 3] / a.count));
        m2 = make_float4((float)(a.sums[C + c] / a.count), (float)(a.sums[C + c + 1] / a.count),
                         (float)(a.sums[C + c + 2] / a.count), (float)(a.sums[C + c + 3] / a.count));
    }
    if (blockIdx.x == 0 && a.sums && threadIdx.x < a.cg) {
        // parameter gradients: dgamma = sum dz * yhat, dbeta = sum dz
        int cc = threadIdx.x * 4;
        for (int j = 0; j < 4; ++j) {
            if (a.dbeta) a.dbeta[cc + j] = (float)a.sums[cc + j];
            if (a.dgamma) a.dgamma[cc + j] = (float)a.sums[C + cc + j];
        }
    }
    float4 sb = f4(0.f);
    // Row-structured: a work item is (row of the padded dy grid, column segment), so the pixel decode (two divisions)
    // happens once per row, and the loads of FOUR pixels are issued before any of them is used -- with one pixel per
    // iteration the kernel had two 16-byte loads in flight per thread and ran at 4.7 TB/s.
    constexpr bool GATHER = SH > 0;
    constexpr int U = GATHER ? 1 : 4;
    constexpr int NR = SH == 1 ? 3 : 2, NC = SW == 1 ? 3 : 2;     // candidate windows per axis (gather)
    const int wl = (int)(threadIdx.x / a.cg), WL = (int)(blockDim.x / a.cg);
    const int DC = a.dy.c;
    const int rows = a.dy.n * a.dy.hp, segs = a.segs, items = rows * segs;
    const int seg_w = (a.dy.wp + segs - 1) / segs;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int row = item / segs, seg = item - row * segs;
        const int x_begin = seg * seg_w, x_end = min(a.dy.wp, x_begin + seg_w);
        const int n = row / a.dy.hp;
        int h = row - n * a.dy.hp - a.dy.ph;
        const unsigned pix0 = (unsigned)row * (unsigned)a.dy.wp;
        const bool zero_row = h < 0 || h >= a.dy.h || (a.hs == 2 && (h & 1));
        if (a.hs == 2) h >>= 1;
        const float *yrow = zero_row ? nullptr : a.yp + a.y.off(n, h, 0) + c;
        const float *zrow = (zero_row || GATHER) ? nullptr : a.dz + (((size_t)n * a.y.h + h) * a.y.w) * C + c;
        const uint8_t *ip[NR];
        const float *dp[NR];
        unsigned rbase[NR];         // (window-relative row) * 3, or 255 when the window row does not exist
        if (GATHER && !zero_row) {
            const int ho0 = SH == 1 ? h - 1 : h >> 1;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int ho = ho0 + i;
                const bool okh = ho >= 0 && ho < a.pooled_h && (SH == 1 || i == 0 || (h & 1));
                const int rh = SH == 1 ? 2 - i : h + 1 - 2 * ho;
                const int hoc = okh ? ho : 0;
                ip[i] = a.idx + ((size_t)(n * a.pooled_h + hoc) * a.pooled_w) * C + c;
                dp[i] = a.doutp + a.dout.off(n, hoc, 0) + a.c_off + c;
                rbase[i] = okh ? (unsigned)(rh * 3) : 255u;
            }
        }
        for (int x0 = x_begin + wl; x0 < x_end; x0 += U * WL) {
            float4 y[U], dz[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = x0 + u * WL, w = x - a.dy.pw;
                ok[u] = !zero_row && x < x_end && w >= 0 && w < a.dy.w;
                if (ok[u]) {
                    y[u] = ld4(yrow + (size_t)w * a.y.c);
                    if (!GATHER) {
                        dz[u] = ld4(zrow + (size_t)w * C);
                    } else {
                        const int wo0 = SW == 1 ? w - 1 : w >> 1;
                        float4 g = f4(0.f);
#pragma unroll
                        for (int j = 0; j < NC; ++j) {
                            const int wo = wo0 + j;
                            const bool okw = wo >= 0 && wo < a.pooled_w && (SW == 1 || j == 0 || (w & 1));
                            const int rw = SW == 1 ? 2 - j : w + 1 - 2 * wo;
                            const int woc = okw ? wo : 0;
#pragma unroll
                            for (int i = 0; i < NR; ++i) {
                                const unsigned bi = *reinterpret_cast<const unsigned *>(ip[i] + woc * C);
                                const float4 d = ld4(dp[i] + woc * a.dout.c);
                                const unsigned r = (okw && rbase[i] != 255u) ? rbase[i] + (unsigned)rw : 255u;
                                const unsigned m = __vcmpeq4(bi, r * 0x01010101u);
                                if (m & 0x000000FFu) g.x += d.x;
                                if (m & 0x0000FF00u) g.y += d.y;
                                if (m & 0x00FF0000u) g.z += d.z;
                                if (m & 0xFF000000u) g.w += d.w;
                            }
                        }
                        dz[u] = g;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = x0 + u * WL;
                if (x >= x_end) break;
                const unsigned pix = pix0 + (unsigned)x;
                const size_t o = (size_t)pix * DC + c;
                // dy may carry more channels than y (a narrow layer padded to the tensor cores' granularity): zeros
                for (int cz = C + c; cz < DC; cz += C) {
                    if (a.dy_hi) st4(a.dy_hi + (size_t)pix * DC + cz, f4(0.f));
                    if (a.dy_lo) st4(a.dy_lo + (size_t)pix * DC + cz, f4(0.f));
                    if (a.dy_h2) st4_h2_zero(a.dy_h2, pix, DC, cz);
                }
                if (!ok[u]) {
                    if (a.dy_hi) st4(a.dy_hi + o, f4(0.f));
                    if (a.dy_lo) st4(a.dy_lo + o, f4(0.f));
                    if (a.dy_h2) st4_h2_zero(a.dy_h2, pix, DC, c);
                    continue;
                }
                const float4 yh = make_float4((y[u].x - mu.x) * is.x, (y[u].y - mu.y) * is.y, (y[u].z - mu.z) * is.z,
                                              (y[u].w - mu.w) * is.w);
                if (!GATHER && a.post_relu) {
                    const float4 v = fma4(sc, y[u], sf);     // the forward pass's BN output, same fma
                    if (!(v.x > 0.f)) dz[u].x = 0.f;
                    if (!(v.y > 0.f)) dz[u].y = 0.f;
                    if (!(v.z > 0.f)) dz[u].z = 0.f;
                    if (!(v.w > 0.f)) dz[u].w = 0.f;
                }
                float4 d = make_float4(sc.x * (dz[u].x - m1.x - yh.x * m2.x), sc.y * (dz[u].y - m1.y - yh.y * m2.y),
                                       sc.z * (dz[u].z - m1.z - yh.z * m2.z), sc.w * (dz[u].w - m1.w - yh.w * m2.w));
                if (a.pre_relu) {
                    if (!(y[u].x > 0.f)) d.x = 0.f;
                    if (!(y[u].y > 0.f)) d.y = 0.f;
                    if (!(y[u].z > 0.f)) d.z = 0.f;
                    if (!(y[u].w > 0.f)) d.w = 0.f;
                }
                if (a.dy_hi) st4_split(a.dy_hi, a.dy_lo, o, d);
                if (a.dy_h2) st4_h2(a.dy_h2, pix, DC, c, d, s16);
                sb = add4(sb, d);
            }
        }
    }
    if (a.dbias) {
        atomicAdd(&red[c + 0], sb.x); atomicAdd(&red[c + 1], sb.y);
        atomicAdd(&red[c + 2], sb.z); atomicAdd(&red[c + 3], sb.w);
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(a.dbias + i, (double)red[i]);
    }
}

// BN-backward sums of a pooled layer WITHOUT a ReLU between BN and pool, from the pooled side: every dout is routed
// to exactly one input position, so sum dz = sum dout and sum dz * yhat = sum dout * yhat(arg-max), with y at the
// arg-max saved by the forward pass (ymax).  sums[2C] = windows * max |dout| bounds |dz| (up to `windows` pooled
// outputs can select the same input).  8 bytes per POOLED element instead of a pass over the conv output.
__global__ void __launch_bounds__(256) pool_bwd_sums_kernel(Geo dout, const float *__restrict__ doutp, int c_off,
                                                            const float *__restrict__ ymax,
                                                            const float *__restrict__ mean,
                                                            const float *__restrict__ invstd, int cg, float windows,
                                                            double *sums) {
    __shared__ float red[2 * MAX_C];
    const int C = cg * 4;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int c = (int)(threadIdx.x % cg) * 4;
    const float4 mu = ld4(mean + c), is = ld4(invstd + c);
    float4 s1 = f4(0.f), s2 = f4(0.f);
    float amax = 0.f;
    const unsigned npix = (unsigned)(dout.n * dout.h * dout.w);
    constexpr int U = 4;
    const unsigned p0 = (blockIdx.x * blockDim.x + threadIdx.x) / (unsigned)cg, step = (gridDim.x * blockDim.x) / (unsigned)cg;
    for (unsigned pix = p0; pix < npix; pix += U * step) {
        float4 d[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned p = pix + u * step;
            if (p < npix) {
                d[u] = ld4(doutp + (size_t)p * dout.c + c_off + c);
                y[u] = ld4(ymax + (size_t)p * C + c);
            } else {
                d[u] = f4(0.f);
                y[u] = mu;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float4 yh = make_float4((y[u].x - mu.x) * is.x, (y[u].y - mu.y) * is.y, (y[u].z - mu.z) * is.z,
                                          (y[u].w - mu.w) * is.w);
            s1 = add4(s1, d[u]);
            s2 = fma4(d[u], yh, s2);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(d[u].x), fabsf(d[u].y)), fmaxf(fabsf(d[u].z), fabsf(d[u].w))));
        }
    }
    amax = warp_max(amax) * windows;
    if ((threadIdx.x & 31) == 0)
        atomicMax(reinterpret_cast<unsigned long long *>(sums + 2 * C), (unsigned long long)__double_as_longlong((double)amax));
    atomicAdd(&red[c + 0], s1.x); atomicAdd(&red[c + 1], s1.y);
    atomicAdd(&red[c + 2], s1.z); atomicAdd(&red[c + 3], s1.w);
    atomicAdd(&red[C + c + 0], s2.x); atomicAdd(&red[C + c + 1], s2.y);
    atomicAdd(&red[C + c + 2], s2.z); atomicAdd(&red[C + c + 3], s2.w);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, (double)red[i]);
}

// ------------------------------------------------------------------ bulk-copy row rings (TMA-staged pooling passes)
// The pooled passes above fetch every window / un-pooling candidate with its own global load: 4.5 x (forward) and
// ~4 x (backward) more L1 / L2 requests than DRAM bytes, and they ran at 3.0 - 3.9 TB/s, bound by L2 bandwidth and
// load latency (ncu: L2 -> L1 906 MB for 283 MB of DRAM reads).  The kernels below stage whole row segments in shared
// memory with cp.async.bulk (one elected thread issues; completion on an mbarrier per ring slot), so every byte
// crosses L2 once, many KB per SM are in flight independent of the per-thread window structure, and the window logic
// reads shared memory.  A CTA owns a contiguous range of (image, column segment, row) units -- rows fastest -- and
// streams the rows of a segment through the ring, so rows are not re-read at unit boundaries either.
__device__ __forceinline__ uint32_t ring_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ring_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ring_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void ring_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct PoolRowMax {
    float4 v, raw;      // best value over the row's window columns, and the conv output y it came from
    unsigned dx;        // its window column (0 .. 2), one byte per channel
};

// Forward: BN-apply (+ReLU) + 3x3 max-pool (+ arg-max bytes, + y at the arg-max) + fp16 split / fp32 store.
// blockDim = cg * WS (channel groups x output columns of a segment).  Shared memory: RING_SLOTS input-row segments
// of ((WS - 1) * SW + 3) columns x C floats.  Same arithmetic and tie-breaking as bn_pool3_fwd_kernel.
template <int SH, int SW, bool RELU>
__global__ void __launch_bounds__(256, 3) bn_pool3_fwd_tma_kernel(BnPool a, int WS, int nseg, long long units_per_cta) {
    // an output row keeps 3 input rows alive and consumes SH new ones: 3 + 3 rows of look-ahead (SH = 1), 3 + 5 (SH = 2)
    constexpr int RING_SLOTS = 6;     // two output rows' worth of input rows; three CTAs per SM fit (ncu: these passes
                                      // are issue-bound, 16 resident warps left 40 % of the issue slots empty)
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ __align__(8) unsigned long long bars[RING_SLOTS];
    const int cg = a.cg, C = cg * 4;
    const int c = (int)(threadIdx.x % cg) * 4, xo = (int)(threadIdx.x / cg);
    const int wcols = (WS - 1) * SW + 3;
    const uint32_t slot_bytes = (uint32_t)wcols * C * 4;
    float *ring = reinterpret_cast<float *>(ring_raw);
    const uint32_t ring0 = ring_smem_u32(ring), bar0 = ring_smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_SLOTS; ++i) ring_bar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const float s16 = a.out_h2 ? f16_scale_from_bound(*a.out_bound) : 1.f;
    float4 sc = f4(1.f), sf = f4(0.f);
    if (a.scale) {
        sc = ld4(a.scale + c);
        sf = ld4(a.shift + c);
    }
    const int H = a.y.h, W = a.y.w, OH = a.out.h, OW = a.out.w;
    const long long total = (long long)a.out.n * nseg * OH;
    long long u = (long long)blockIdx.x * units_per_cta;
    const long long u_end = min(total, u + units_per_cta);
    unsigned k_issue = 0, k_base = 0;       // stream indices (input rows issued / first row of the current span)
    while (u < u_end) {
        // ---- span: rows [ho_a, ho_b) of column segment seg of image n
        const int ho_a = (int)(u % OH);
        const long long t = u / OH;
        const int seg = (int)(t % nseg), n = (int)(t / nseg);
        const int ho_b = (int)min((long long)OH, ho_a + (u_end - u));
        const int x0 = seg * WS;                                   // first output column of the segment
        const int w_lo = x0 * SW - 1;                              // input column of ring column 0
        const int wa = max(w_lo, 0), wb = min(w_lo + wcols, W);    // valid input columns [wa, wb)
        const uint32_t dst_off = (uint32_t)(wa - w_lo) * C * 4, bytes = (uint32_t)(wb - wa) * C * 4;
        const int r_first = max(ho_a * SH - 1, 0), r_last = min((ho_b - 1) * SH + 1, H - 1);
        const float *src0 = a.yp + a.y.off(n, 0, wa);
        const size_t row_stride = (size_t)a.y.wp * a.y.c;
        int r_issue = r_first, r_waited = r_first - 1;             // next input row to issue / last row waited for
        const int wo = x0 + xo;
        const bool col_ok = wo < OW;
        const int w0 = wo * SW - 1, ci0 = xo * SW;                 // window column 0: input column / ring column
        const bool okx0 = w0 >= 0, okx1 = w0 + 1 < W, okx2 = w0 + 2 < W;    // (w0 + 1 >= 0 always; ceil-mode windows may overhang)
        PoolRowMax rm[3];
        // horizontal maximum of input row h over the window's three columns: first maximum wins (strict >)
        auto hrow = [&](int h) {
            PoolRowMax r;
            r.v = f4(-FLT_MAX);
            r.raw = f4(0.f);
            r.dx = 0u;
            if (h < 0 || h >= H) return r;
            const unsigned k = k_base + (unsigned)(h - r_first);
            const float *row = ring + (size_t)(k % RING_SLOTS) * (slot_bytes / 4) + c + ci0 * C;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                if ((dx == 0 && !okx0) || (dx == 1 && !okx1) || (dx == 2 && !okx2)) continue;
                const float4 raw = ld4(row + dx * C);
                float4 v = fma4(sc, raw, sf);
                if (RELU) v = relu4(v);
                if (v.x > r.v.x) { r.v.x = v.x; r.raw.x = raw.x; r.dx = (r.dx & 0xFFFFFF00u) | (unsigned)dx; }
                if (v.y > r.v.y) { r.v.y = v.y; r.raw.y = raw.y; r.dx = (r.dx & 0xFFFF00FFu) | ((unsigned)dx << 8); }
                if (v.z > r.v.z) { r.v.z = v.z; r.raw.z = raw.z; r.dx = (r.dx & 0xFF00FFFFu) | ((unsigned)dx << 16); }
                if (v.w > r.v.w) { r.v.w = v.w; r.raw.w = raw.w; r.dx = (r.dx & 0x00FFFFFFu) | ((unsigned)dx << 24); }
            }
            return r;
        };
        for (int ho = ho_a; ho < ho_b; ++ho) {
            const int need_lo = max(ho * SH - 1, 0), need_hi = min(ho * SH + 1, H - 1);
            if (threadIdx.x == 0) {
                // rows below need_lo are dead (the __syncthreads of the previous iteration ordered their last reads)
                while (r_issue <= r_last && r_issue - need_lo < RING_SLOTS) {
                    const unsigned s = k_issue % RING_SLOTS;
                    ring_expect_tx(bar0 + 8 * s, bytes);
                    ring_bulk_load(ring0 + s * slot_bytes + dst_off, src0 + (size_t)r_issue * row_stride, bytes, bar0 + 8 * s);
                    ++r_issue;
                    ++k_issue;
                }
            }
            for (int r = max(need_lo, r_waited + 1); r <= need_hi; ++r) {   // rows waited for earlier stay complete
                const unsigned k = k_base + (unsigned)(r - r_first);
                ring_wait(bar0 + 8 * (k % RING_SLOTS), (k / RING_SLOTS) & 1u);
            }
            r_waited = max(r_waited, need_hi);
            const unsigned pix = ((unsigned)(n * a.out.hp + ho + a.out.ph)) * (unsigned)a.out.wp + (unsigned)(wo + a.out.pw);
            if (col_ok) {
                // separable, rolling: the horizontal (value, raw y, dx) maximum of an input row is computed once and
                // reused by every output row whose window contains it (3 with SH = 1, up to 2 with SH = 2); the kernel
                // is bound by instruction issue, not by memory (ncu: 66 % issue slots busy at 35 % of DRAM bandwidth)
                if (SH == 1) {
                    if (ho == ho_a) {
                        rm[0] = hrow(ho - 1);
                        rm[1] = hrow(ho);
                    }
                    rm[2] = hrow(ho + 1);
                } else {
                    if (ho == ho_a) rm[0] = hrow(2 * ho - 1);
                    rm[1] = hrow(2 * ho);
                    rm[2] = hrow(2 * ho + 1);
                }
                float4 best = rm[0].v, yb = rm[0].raw;
                unsigned bi = rm[0].dx;
#pragma unroll
                for (int dy = 1; dy < 3; ++dy) {
                    const unsigned code = rm[dy].dx + 0x01010101u * (unsigned)(3 * dy);
                    if (rm[dy].v.x > best.x) { best.x = rm[dy].v.x; yb.x = rm[dy].raw.x; bi = (bi & 0xFFFFFF00u) | (code & 0x000000FFu); }
                    if (rm[dy].v.y > best.y) { best.y = rm[dy].v.y; yb.y = rm[dy].raw.y; bi = (bi & 0xFFFF00FFu) | (code & 0x0000FF00u); }
                    if (rm[dy].v.z > best.z) { best.z = rm[dy].v.z; yb.z = rm[dy].raw.z; bi = (bi & 0xFF00FFFFu) | (code & 0x00FF0000u); }
                    if (rm[dy].v.w > best.w) { best.w = rm[dy].v.w; yb.w = rm[dy].raw.w; bi = (bi & 0x00FFFFFFu) | (code & 0xFF000000u); }
                }
                if (SH == 1) {
                    rm[0] = rm[1];
                    rm[1] = rm[2];
                } else {
                    rm[0] = rm[2];
                }
                const size_t po = ((size_t)(n * OH + ho) * OW + wo) * C + c;
                if (a.idx) *reinterpret_cast<unsigned *>(a.idx + po) = bi;
                if (a.ymax) st4(a.ymax + po, yb);
                bnpool_store(a, pix, c, best, s16);
            }
            // zero pads of the output tensor around this row segment
            if (a.out.pw > 0 && (seg == 0 || seg == nseg - 1)) {
                const unsigned rowpix = ((unsigned)(n * a.out.hp + ho + a.out.ph)) * (unsigned)a.out.wp;
                for (int i = xo; i < a.out.pw; i += WS) {
                    if (seg == 0) bnpool_store(a, rowpix + i, c, f4(0.f), 1.f);
                    if (seg == nseg - 1) bnpool_store(a, rowpix + a.out.pw + OW + i, c, f4(0.f), 1.f);
                }
            }
            if (a.out.ph > 0 && (ho == 0 || ho == OH - 1)) {
                // pad rows above / below: the columns of this segment (+ the column pads at the outer segments)
                const int xa = seg == 0 ? 0 : a.out.pw + x0;
                const int xb = seg == nseg - 1 ? a.out.wp : a.out.pw + min(x0 + WS, OW);
                for (int pr = 0; pr < a.out.ph; ++pr) {
                    if (ho == 0) {
                        const unsigned rowpix = ((unsigned)(n * a.out.hp + pr)) * (unsigned)a.out.wp;
                        for (int x = xa + xo; x < xb; x += WS) bnpool_store(a, rowpix + x, c, f4(0.f), 1.f);
                    }
                    if (ho == OH - 1) {
                        const unsigned rowpix = ((unsigned)(n * a.out.hp + a.out.ph + OH + pr)) * (unsigned)a.out.wp;
                        for (int x = xa + xo; x < xb; x += WS) bnpool_store(a, rowpix + x, c, f4(0.f), 1.f);
                    }
                }
            }
            __syncthreads();
        }
        k_base += (unsigned)(r_last - r_first + 1);
        u += ho_b - ho_a;
    }
}

// Backward apply pass of a  conv -> (ReLU) -> BN -> 3x3 max-pool  layer with the un-pooling inside (the gather variant
// of bn_bwd_apply_kernel), rows staged by bulk copies: the conv output y (each element read once), and the pooled-side
// rows of dout and of the arg-max bytes that the 2 - 3 input rows and 1.5 - 3 input columns of a window share.
// blockDim = cg * WT; a thread owns a channel group and U input columns WT apart.  dout.c == C, c_off == 0.
constexpr int APPLY_YSLOTS_MAX = 4, APPLY_PSLOTS_MAX = 6, APPLY_U = 2;
// ring depths: y rows / pooled rows in flight per CTA.  H stride 2 touches at most two pooled rows per input row, so
// shallower rings keep three CTAs per SM resident at C = 512 too
__host__ __device__ constexpr int apply_yslots(int sh) { return sh == 1 ? 4 : 3; }
__host__ __device__ constexpr int apply_pslots(int sh) { return sh == 1 ? 6 : 4; }
template <int SH, int SW>
__global__ void __launch_bounds__(256, 3) bn_pool_bwd_apply_tma_kernel(BnApply a, int WT, int nseg, long long units_per_cta) {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ __align__(8) unsigned long long bars[APPLY_YSLOTS_MAX + APPLY_PSLOTS_MAX];
    __shared__ float red[MAX_C];
    __shared__ float wred[32];
    constexpr int APPLY_YSLOTS = apply_yslots(SH), APPLY_PSLOTS = apply_pslots(SH);
    const int cg = a.cg, C = cg * 4;
    constexpr int NR = SH == 1 ? 3 : 2, NC = SW == 1 ? 3 : 2;
    const int WSI = APPLY_U * WT;                                   // input columns per segment
    const int ocols = WSI / SW + 2;                                 // pooled columns a segment can touch
    const uint32_t yslot = (uint32_t)WSI * C * 4, dslot = (uint32_t)ocols * C * 4, islot = (uint32_t)ocols * C;
    const uint32_t ring0 = ring_smem_u32(ring_raw), bar0 = ring_smem_u32(bars);
    const uint32_t dring0 = ring0 + APPLY_YSLOTS * yslot, iring0 = dring0 + APPLY_PSLOTS * dslot;
    const float *yring = reinterpret_cast<const float *>(ring_raw);
    const float *dring = reinterpret_cast<const float *>(ring_raw + (size_t)APPLY_YSLOTS * yslot);
    const uint8_t *iring = ring_raw + (size_t)APPLY_YSLOTS * yslot + (size_t)APPLY_PSLOTS * dslot;
    if (threadIdx.x == 0) {
        for (int i = 0; i < APPLY_YSLOTS + APPLY_PSLOTS; ++i) ring_bar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- prologue shared with bn_bwd_apply_kernel: bound of dy, parameter gradients, per-channel constants
    float s16 = 1.f;
    if (a.dy_h2) {
        const float mz = (float)a.sums[2 * C];
        float b = 0.f;
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            float t = mz;
            if (a.batch_stats) t += (float)(fabs(a.sums[i]) / a.count + sqrt(a.count) * fabs(a.sums[C + i]) / a.count);
            b = fmaxf(b, fabsf(a.scale ? a.scale[i] : 1.f) * t);
        }
        b = warp_max(b);
        if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = b;
        __syncthreads();
        b = 0.f;
        for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) b = fmaxf(b, wred[i]);
        b *= 1.001f;
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.dy_bound = b;
        s16 = f16_scale_from_bound(b);
    }
    if (a.dbias)
        for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int c = (int)(threadIdx.x % cg) * 4, xi = (int)(threadIdx.x / cg);
    float4 sc = f4(1.f), mu = f4(0.f), is = f4(1.f), m1 = f4(0.f), m2 = f4(0.f);
    if (a.scale) sc = ld4(a.scale + c);
    if (a.mean) {
        mu = ld4(a.mean + c);
        is = ld4(a.invstd + c);
    }
    if (a.sums && a.batch_stats) {
        m1 = make_float4((float)(a.sums[c] / a.count), (float)(a.sums[c + 1] / a.count),
                         (float)(a.sums[c + 2] / a.count), (float)(a.sums[c + 3] / a.count));
        m2 = make_float4((float)(a.sums[C + c] / a.count), (float)(a.sums[C + c + 1] / a.count),
                         (float)(a.sums[C + c + 2] / a.count), (float)(a.sums[C + c + 3] / a.count));
    }
    if (blockIdx.x == 0 && a.sums && threadIdx.x < cg) {
        const int cc = threadIdx.x * 4;
        for (int j = 0; j < 4; ++j) {
            if (a.dbeta) a.dbeta[cc + j] = (float)a.sums[cc + j];
            if (a.dgamma) a.dgamma[cc + j] = (float)a.sums[C + cc + j];
        }
    }
    // dy = sc * (dz - m1 - (y - mu) * is * m2) = sc * dz + cA + cB * y
    const float4 cB = make_float4(-sc.x * is.x * m2.x, -sc.y * is.y * m2.y, -sc.z * is.z * m2.z, -sc.w * is.w * m2.w);
    const float4 cA = make_float4(-sc.x * m1.x - cB.x * mu.x, -sc.y * m1.y - cB.y * mu.y, -sc.z * m1.z - cB.z * mu.z,
                                  -sc.w * m1.w - cB.w * mu.w);
    float4 sb = f4(0.f);
    const int H = a.y.h, W = a.y.w, OH = a.pooled_h, OW = a.pooled_w;
    const long long total = (long long)a.y.n * nseg * H;
    long long u = (long long)blockIdx.x * units_per_cta;
    const long long u_end = min(total, u + units_per_cta);
    unsigned ky_issue = 0, ky_base = 0, kp_issue = 0, kp_base = 0;
    while (u < u_end) {
        const int h_a = (int)(u % H);
        const long long t = u / H;
        const int seg = (int)(t % nseg), n = (int)(t / nseg);
        const int h_b = (int)min((long long)H, h_a + (u_end - u));
        const int w_seg = seg * WSI, w_end = min(W, w_seg + WSI);
        const uint32_t ybytes = (uint32_t)(w_end - w_seg) * C * 4;
        // pooled columns [o_lo, o_hi) that the segment's input columns can select
        const int o_lo = SW == 1 ? max(w_seg - 1, 0) : (w_seg >> 1);
        const int o_hi = min(OW, SW == 1 ? w_end + 1 : ((w_end - 1) >> 1) + 2);
        const uint32_t dbytes = (uint32_t)(o_hi - o_lo) * C * 4, ibytes = (uint32_t)(o_hi - o_lo) * C;
        // pooled rows the rows [h_a, h_b) can select
        const int p_first = SH == 1 ? max(h_a - 1, 0) : (h_a >> 1);
        const int p_last = min(OH - 1, SH == 1 ? h_b : (((h_b - 1) >> 1) + ((h_b - 1) & 1)));
        const float *ysrc = a.yp + a.y.off(n, 0, w_seg);
        const size_t yrow = (size_t)a.y.wp * a.y.c;
        const float *dsrc = a.doutp + a.dout.off(n, 0, o_lo);
        const size_t drow = (size_t)a.dout.wp * a.dout.c;
        const uint8_t *isrc = a.idx + ((size_t)n * OH * OW + o_lo) * C;
        const size_t irow = (size_t)OW * C;
        int y_issue = h_a, p_issue = p_first, p_waited = p_first - 1;
        for (int h = h_a; h < h_b; ++h) {
            const int ho0 = SH == 1 ? h - 1 : h >> 1;
            const int p_lo = max(ho0, 0);
            const int p_hi = min(OH - 1, SH == 1 ? h + 1 : ho0 + (h & 1));
            if (threadIdx.x == 0) {
                while (y_issue < h_b && y_issue - h < APPLY_YSLOTS) {
                    const unsigned s = ky_issue % APPLY_YSLOTS;
                    ring_expect_tx(bar0 + 8 * s, ybytes);
                    ring_bulk_load(ring0 + s * yslot, ysrc + (size_t)y_issue * yrow, ybytes, bar0 + 8 * s);
                    ++y_issue;
                    ++ky_issue;
                }
                while (p_issue <= p_last && p_issue - p_lo < APPLY_PSLOTS) {
                    const unsigned s = kp_issue % APPLY_PSLOTS;
                    const uint32_t bar = bar0 + 8 * (APPLY_YSLOTS + s);
                    ring_expect_tx(bar, dbytes + ibytes);
                    ring_bulk_load(dring0 + s * dslot, dsrc + (size_t)p_issue * drow, dbytes, bar);
                    ring_bulk_load(iring0 + s * islot, isrc + (size_t)p_issue * irow, ibytes, bar);
                    ++p_issue;
                    ++kp_issue;
                }
            }
            {
                const unsigned k = ky_base + (unsigned)(h - h_a);
                ring_wait(bar0 + 8 * (k % APPLY_YSLOTS), (k / APPLY_YSLOTS) & 1u);
            }
            for (int p = max(p_lo, p_waited + 1); p <= p_hi; ++p) {      // rows waited for by an earlier h stay complete
                const unsigned k = kp_base + (unsigned)(p - p_first);
                ring_wait(bar0 + 8 * (APPLY_YSLOTS + k % APPLY_PSLOTS), (k / APPLY_PSLOTS) & 1u);
            }
            p_waited = max(p_waited, p_hi);
            const unsigned ks = ky_base + (unsigned)(h - h_a);
            const float *yrow_s = yring + (ks % APPLY_YSLOTS) * (yslot / 4) + c;
            // candidate pooled rows of this input row (block-uniform): ring rows and (window-relative row) * 3
            const int nr = SH == 1 ? 3 : 1 + (h & 1);
            const float *dp[NR];
            const uint8_t *ip[NR];
            unsigned rbase[NR];
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int ho = ho0 + i;
                const bool okh = i < nr && ho >= 0 && ho < OH;
                const int rh = SH == 1 ? 2 - i : h + 1 - 2 * ho;
                const unsigned k = kp_base + (unsigned)((okh ? ho : p_lo) - p_first);
                dp[i] = dring + (k % APPLY_PSLOTS) * (dslot / 4) + c;
                ip[i] = iring + (k % APPLY_PSLOTS) * islot + c;
                rbase[i] = okh ? (unsigned)(rh * 3) : 64u;      // 64 + rw never equals an arg-max byte (0 .. 8)
            }
            const unsigned pixrow = ((unsigned)(n * a.dy.hp + h + a.dy.ph)) * (unsigned)a.dy.wp + (unsigned)a.dy.pw;
            // dy = scale * (dz - mean(dz) - yhat * mean(dz * yhat)) = scale * dz + (cA + cB * y), constants per channel
            auto finish = [&](int w, const float4 &y, const float4 &g) {
                float4 d = make_float4(fmaf(sc.x, g.x, fmaf(cB.x, y.x, cA.x)), fmaf(sc.y, g.y, fmaf(cB.y, y.y, cA.y)),
                                       fmaf(sc.z, g.z, fmaf(cB.z, y.z, cA.z)), fmaf(sc.w, g.w, fmaf(cB.w, y.w, cA.w)));
                if (a.pre_relu) {
                    if (!(y.x > 0.f)) d.x = 0.f;
                    if (!(y.y > 0.f)) d.y = 0.f;
                    if (!(y.z > 0.f)) d.z = 0.f;
                    if (!(y.w > 0.f)) d.w = 0.f;
                }
                const unsigned pix = pixrow + (unsigned)w;
                if (a.dy_hi) st4_split(a.dy_hi, a.dy_lo, (size_t)pix * C + c, d);
                if (a.dy_h2) st4_h2(a.dy_h2, pix, C, c, d, s16);
                sb = add4(sb, d);
            };
            if (SW == 2) {
                // a thread owns the pixel pair (2 wo, 2 wo + 1): both can be the arg-max of window column wo (window
                // columns 1 and 2), the odd one also of window column wo + 1 (window column 0) -- the candidate loads
                // of column wo are shared, and the arg-max bytes are unpacked once per load
                const int oc = xi;                                  // ring column of pooled column wo = o_lo + xi
                const int w_e = w_seg + 2 * xi;
                if (w_e < w_end) {
                    const bool odd_ok = w_e + 1 < w_end;
                    const bool okA = o_lo + oc < OW, okB = o_lo + oc + 1 < OW;
                    float4 ge = f4(0.f), go = f4(0.f);
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        if (i >= nr) break;
                        {
                            const unsigned bi = okA ? *reinterpret_cast<const unsigned *>(ip[i] + oc * C) : 0xFFFFFFFFu;
                            const float4 d = ld4(dp[i] + oc * C);
                            const unsigned r1 = rbase[i] + 1u, r2 = rbase[i] + 2u;
                            const unsigned b0 = bi & 0xFFu, b1 = (bi >> 8) & 0xFFu, b2 = (bi >> 16) & 0xFFu, b3 = bi >> 24;
                            if (b0 == r1) ge.x += d.x;
                            if (b1 == r1) ge.y += d.y;
                            if (b2 == r1) ge.z += d.z;
                            if (b3 == r1) ge.w += d.w;
                            if (b0 == r2) go.x += d.x;
                            if (b1 == r2) go.y += d.y;
                            if (b2 == r2) go.z += d.z;
                            if (b3 == r2) go.w += d.w;
                        }
                        {
                            const unsigned bi = okB ? *reinterpret_cast<const unsigned *>(ip[i] + (oc + 1) * C) : 0xFFFFFFFFu;
                            const float4 d = ld4(dp[i] + (oc + 1) * C);
                            const unsigned r0 = rbase[i];
                            if ((bi & 0xFFu) == r0) go.x += d.x;
                            if (((bi >> 8) & 0xFFu) == r0) go.y += d.y;
                            if (((bi >> 16) & 0xFFu) == r0) go.z += d.z;
                            if ((bi >> 24) == r0) go.w += d.w;
                        }
                    }
                    const float *yp2 = yrow_s + (w_e - w_seg) * C;
                    finish(w_e, ld4(yp2), ge);
                    if (odd_ok) finish(w_e + 1, ld4(yp2 + C), go);
                }
            } else {
#pragma unroll
                for (int uu = 0; uu < APPLY_U; ++uu) {
                    const int w = w_seg + xi + uu * WT;
                    if (w >= w_end) break;
                    const int wo0 = w - 1;
                    float4 g = f4(0.f);
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        const int wo = wo0 + j;
                        const bool okw = wo >= 0 && wo < OW;
                        const unsigned rw = (unsigned)(2 - j);
                        const int oc = (okw ? wo : o_lo) - o_lo;
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            if (i >= nr) break;
                            const unsigned bi = okw ? *reinterpret_cast<const unsigned *>(ip[i] + oc * C) : 0xFFFFFFFFu;
                            const float4 d = ld4(dp[i] + oc * C);
                            const unsigned r = rbase[i] + rw;
                            if ((bi & 0xFFu) == r) g.x += d.x;
                            if (((bi >> 8) & 0xFFu) == r) g.y += d.y;
                            if (((bi >> 16) & 0xFFu) == r) g.z += d.z;
                            if ((bi >> 24) == r) g.w += d.w;
                        }
                    }
                    finish(w, ld4(yrow_s + (w - w_seg) * C), g);
                }
            }
            // zero pads of the dy grid around this row segment
            if (a.dy.pw > 0 && (seg == 0 || seg == nseg - 1)) {
                const unsigned rowpix = ((unsigned)(n * a.dy.hp + h + a.dy.ph)) * (unsigned)a.dy.wp;
                for (int i = xi; i < a.dy.pw; i += WT) {
                    for (int e = 0; e < 2; ++e) {
                        if (e == 0 ? seg != 0 : seg != nseg - 1) continue;
                        const unsigned pix = rowpix + (e == 0 ? i : a.dy.pw + W + i);
                        if (a.dy_hi) st4(a.dy_hi + (size_t)pix * C + c, f4(0.f));
                        if (a.dy_lo) st4(a.dy_lo + (size_t)pix * C + c, f4(0.f));
                        if (a.dy_h2) st4_h2_zero(a.dy_h2, pix, C, c);
                    }
                }
            }
            if (a.dy.ph > 0 && (h == 0 || h == H - 1)) {
                const int xa = seg == 0 ? 0 : a.dy.pw + w_seg;
                const int xb = seg == nseg - 1 ? a.dy.wp : a.dy.pw + w_end;
                for (int pr = 0; pr < a.dy.ph; ++pr) {
                    for (int e = 0; e < 2; ++e) {
                        if (e == 0 ? h != 0 : h != H - 1) continue;
                        const unsigned rowpix = ((unsigned)(n * a.dy.hp + (e == 0 ? pr : a.dy.ph + H + pr))) * (unsigned)a.dy.wp;
                        for (int x = xa + xi; x < xb; x += WT) {
                            const unsigned pix = rowpix + x;
                            if (a.dy_hi) st4(a.dy_hi + (size_t)pix * C + c, f4(0.f));
                            if (a.dy_lo) st4(a.dy_lo + (size_t)pix * C + c, f4(0.f));
                            if (a.dy_h2) st4_h2_zero(a.dy_h2, pix, C, c);
                        }
                    }
                }
            }
            __syncthreads();
        }
        ky_base += (unsigned)(h_b - h_a);
        kp_base += (unsigned)(p_last - p_first + 1);
        u += h_b - h_a;
    }
    if (a.dbias) {
        atomicAdd(&red[c + 0], sb.x); atomicAdd(&red[c + 1], sb.y);
        atomicAdd(&red[c + 2], sb.z); atomicAdd(&red[c + 3], sb.w);
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(a.dbias + i, (double)red[i]);
    }
}

__global__ void f64_to_f32_kernel(const double *__restrict__ src, float *__restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

// ------------------------------------------------------------------ spatial reductions
// mode 0: out[n, c_off + c] (+)= inv * sum_hw act(scale * x + shift)
// mode 1: out[n, c_off + c] (+)= sum_hw x * other          (SE gate gradient)
struct SpatialRed {
    Geo x, other;
    const float *xp, *otherp, *scale, *shift;
    int relu, mode, cg, ld_out, c_off, use_atomic;
    float inv;
    float *out;
};

__global__ void __launch_bounds__(256) spatial_reduce_kernel(SpatialRed a) {
    __shared__ float red[MAX_C];
    const int C = a.cg * 4;
    for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int n = blockIdx.y;
    const int c = (int)(threadIdx.x % a.cg) * 4;
    const int row = threadIdx.x / a.cg, rows = blockDim.x / a.cg;
    float4 sc = f4(1.f), sf = f4(0.f);
    if (a.scale) {
        sc = ld4(a.scale + c);
        sf = ld4(a.shift + c);
    }
    float4 s = f4(0.f);
    const int hw = a.x.h * a.x.w;
    for (int p = blockIdx.x * rows + row; p < hw; p += gridDim.x * rows) {
        int h = p / a.x.w, w = p - h * a.x.w;
        float4 v = ld4(a.xp + a.x.off(n, h, w) + c);
        if (a.mode == 0) {
            if (a.scale) v = fma4(sc, v, sf);
            if (a.relu) v = relu4(v);
            s = add4(s, v);
        } else {
            s = fma4(v, ld4(a.otherp + a.other.off(n, h, w) + c), s);
        }
    }
    atomicAdd(&red[c + 0], s.x); atomicAdd(&red[c + 1], s.y);
    atomicAdd(&red[c + 2], s.z); atomicAdd(&red[c + 3], s.w);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float *o = a.out + (size_t)n * a.ld_out + a.c_off + i;
        if (a.use_atomic) atomicAdd(o, red[i] * a.inv);
        else *o = red[i] * a.inv;
    }
}

__global__ void zero_strided_kernel(float *p, int rows, int cols, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) p[(size_t)(i / cols) * ld + (i % cols)] = 0.f;
}

// ------------------------------------------------------------------ SE channel scaling
__global__ void __launch_bounds__(256) channel_scale_fwd_kernel(Geo x, const float *__restrict__ xp,
                                                                const float *__restrict__ gate, Geo o,
                                                                float *out_hi, float *out_lo, int cg) {
    const int c = (int)(threadIdx.x % cg) * 4;
    DLIO_PIX_LOOP(pix, o.n * o.hp * o.wp, cg) {
        int xx = (int)(pix % (unsigned)o.wp);
        unsigned t = pix / (unsigned)o.wp;
        int yy = (int)(t % (unsigned)o.hp);
        int n = (int)(t / (unsigned)o.hp);
        int h = yy - o.ph, w = xx - o.pw;
        size_t oo = (size_t)pix * o.c + c;
        if (h < 0 || h >= o.h || w < 0 || w >= o.w) {
            st4(out_hi + oo, f4(0.f));
            if (out_lo) st4(out_lo + oo, f4(0.f));
            continue;
        }
        float4 v = mul4(ld4(xp + x.off(n, h, w) + c), ld4(gate + (size_t)n * x.c + c));
        st4_split(out_hi, out_lo, oo, v);
    }
}

// dx[n,h,w,c] = dout * gate[n,c] + dmean[n,c] * inv   (dx, dout unpadded)
__global__ void __launch_bounds__(256) channel_scale_bwd_kernel(const float *__restrict__ dout,
                                                                const float *__restrict__ gate,
                                                                const float *__restrict__ dmean, float inv,
                                                                float *__restrict__ dx, int n, int hw, int cg) {
    const int C = cg * 4;
    const int c = (int)(threadIdx.x % cg) * 4;
    DLIO_PIX_LOOP(pix, n * hw, cg) {
        int in = (int)(pix / hw);
        float4 g = ld4(gate + (size_t)in * C + c);
        float4 v = mul4(ld4(dout + (size_t)pix * C + c), g);
        if (dmean) {
            float4 m = ld4(dmean + (size_t)in * C + c);
            v = make_float4(fmaf(m.x, inv, v.x), fmaf(m.y, inv, v.y), fmaf(m.z, inv, v.z), fmaf(m.w, inv, v.w));
        }
        st4(dx + (size_t)pix * C + c, v);
    }
}

// ------------------------------------------------------------------ element-wise helpers
__global__ void axpby_kernel(const float *__restrict__ a, float alpha, const float *__restrict__ b, float beta,
                             float *out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n4 = n >> 2;
    for (long long q = i; q < n4; q += stride) {
        float4 x = ld4(a + q * 4), y = ld4(b + q * 4);
        st4(out + q * 4, make_float4(alpha * x.x + beta * y.x, alpha * x.y + beta * y.y, alpha * x.z + beta * y.z,
                                     alpha * x.w + beta * y.w));
    }
    for (long long q = n4 * 4 + i; q < n; q += stride) out[q] = alpha * a[q] + beta * b[q];
}
// out[a, c] = sum_t x[a, t, c]
__global__ void sum_mid_kernel(const float *__restrict__ x, float *__restrict__ out, long long A, int T, int C) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= A * C) return;
    long long a = i / C;
    int c = (int)(i - a * C);
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += x[((size_t)a * T + t) * C + c];
    out[i] = s;
}
// dst[r * ldd + c] = src[r * lds + c]: a [rows, cols] block between two row-strided buffers (concatenation along
// the last axis and its backward, stacking per-window features)
__global__ void copy2d_kernel(const float *__restrict__ src, long long lds, float *__restrict__ dst, long long ldd,
                              long long rows, int cols) {
    const long long total = rows * cols;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        dst[r * ldd + c] = src[r * lds + c];
    }
}
__global__ void mul_kernel(const float *__restrict__ a, const float *__restrict__ b, float *out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = a[i] * b[i];
}
// counter-based keep mask: splitmix64(seed + i) -> uniform [0,1); mask = keep ? 1/(1-p) : 0
// epoch (optional, device memory): added to the seed, so that a launch replayed from a CUDA graph draws a new mask
// whenever the owner of the graph has advanced the counter
__global__ void dropout_mask_kernel(float *mask, long long n, float p, unsigned long long seed,
                                    const unsigned long long *__restrict__ epoch) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float keep_scale = 1.f / (1.f - p);
    if (epoch) seed += *epoch * 0xD1B54A32D192ED03ULL;
    for (; i < n; i += stride) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        float u = (float)(z >> 40) * (1.f / 16777216.f);
        mask[i] = u >= p ? keep_scale : 0.f;
    }
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_set_option(const char *name, int value) {
    DLIO_CHECK_ARG(name, "set_option: null name");
    if (!strcmp(name, "pool_tma")) dlio::g_pool_tma = value ? 1 : 0;
    else if (!strcmp(name, "ew_block")) dlio::g_ew_cap = value < 64 ? 64 : (value > 256 ? 256 : value);
    else if (!strcmp(name, "conv_cg2")) dlio::g_conv_cg2 = value ? 1 : 0;
    else if (!strcmp(name, "nvtx")) dlio::g_nvtx = value ? 1 : 0;
    else if (!strcmp(name, "bwd_single_pass")) dlio::g_bwd_single = value ? 1 : 0;
    else if (!strcmp(name, "apply_rows")) dlio::g_apply_rows = value ? 1 : 0;
    else {
        set_error("set_option: unknown option %s", name);
        return DLIO_ERR_INVALID;
    }
    return DLIO_OK;
}

extern "C" size_t dlio_workspace_bytes(const char *op, const long long *dims, int ndims) {
    const size_t bad = (size_t)-1;
    if (!op || (ndims > 0 && !dims)) {
        set_error("workspace_bytes: null argument");
        return bad;
    }
    auto need = [&](int n) {
        if (ndims == n) return true;
        set_error("workspace_bytes: %s takes %d dims, got %d", op, n, ndims);
        return false;
    };
    if (!strcmp(op, "conv2d_fwd")) return need(1) ? (size_t)(2 * dims[0]) * sizeof(double) : bad;
    if (!strcmp(op, "bn_bwd")) return need(1) ? (size_t)(2 * dims[0] + 1) * sizeof(double) : bad;
    if (!strcmp(op, "rnn_fwd") || !strcmp(op, "rnn_bwd")) {
        if (!need(7)) return bad;
        const int d[7] = {(int)dims[0], (int)dims[1], (int)dims[2], (int)dims[3], (int)dims[4], (int)dims[5], (int)dims[6]};
        const size_t f = op[4] == 'f' ? dlio_rnn_reserve_floats(d[0], d[1], d[2], d[3], d[4], d[5], d[6])
                                      : dlio_rnn_bwd_scratch_floats(d[0], d[1], d[2], d[3], d[4], d[5], d[6]);
        return f * sizeof(float);
    }
    if (!strcmp(op, "scan_project")) return need(2) ? dlio_scan_scratch_bytes((int)dims[0], (int)dims[1]) : bad;
    set_error("workspace_bytes: unknown op %s", op);
    return bad;
}

extern "C" int dlio_pack_input(const float *src, long long sn, long long st, long long sc, int T, int C,
                               dlio_tensor4 dst, float *dst_ptr, float *dst_lo, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(src && dst_ptr && valid_t4(dst) && dst.c % 4 == 0 && dst.c >= T * C, "pack_input: bad argument");
    Geo d(dst);
    long long total = (long long)d.n * d.hp * d.wp;
    pack_input_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(src, sn, st, sc, T, C, d, dst_ptr, dst_lo);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_bn_finalize(const double *stats, long long count, int c, const float *gamma,
                                const float *beta, float *running_mean, float *running_var, float momentum,
                                float eps, int use_running, float *mean, float *invstd, float *scale,
                                float *shift, const float *res_bound, float *out_bound,
                                long long *num_batches_tracked, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(c > 0 && mean && invstd && scale && shift, "bn_finalize: bad argument");
    DLIO_CHECK_ARG(use_running ? (running_mean && running_var) : (stats && count > 0), "bn_finalize: missing statistics");
    DLIO_CHECK_ARG(!out_bound || (stats && count > 0), "bn_finalize: the output bound needs the batch statistics");
    bn_finalize_kernel<<<1, c >= 1024 ? 1024 : (c + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(
        stats, (double)count, c, gamma, beta, running_mean, running_var, momentum, eps, use_running, mean, invstd,
        scale, shift, res_bound, out_bound, num_batches_tracked);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

static int check_cg(int c, const char *who) {
    DLIO_CHECK_ARG(c % 4 == 0 && c > 0 && c <= MAX_C, "%s: channel count %d must be a multiple of 4 and <= %d", who, c, MAX_C);
    return DLIO_OK;
}

extern "C" int dlio_bn_act_pool_fwd(dlio_tensor4 y, const float *y_ptr, const float *scale, const float *shift,
                                    dlio_tensor4 res, const float *res_ptr, dlio_bnpool p, dlio_tensor4 out,
                                    float *out_hi, float *out_lo, void *out_h2, const float *out_bound,
                                    uint8_t *pool_idx, float *pool_ymax, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(valid_t4(y) && valid_t4(out) && y_ptr && (out_hi || out_h2), "bn_act_pool_fwd: bad argument");
    DLIO_CHECK_ARG(!out_h2 || (out_bound && (((uintptr_t)out_h2) & 15) == 0), "bn_act_pool_fwd: fp16 planes need their bound");
    DLIO_CHECK_ARG(out_hi || !out_lo, "bn_act_pool_fwd: out_lo without out_hi");
    int rc = check_cg(y.c, "bn_act_pool_fwd");
    if (rc) return rc;
    DLIO_CHECK_ARG(p.c_off % 4 == 0 && p.c_off + y.c <= out.c && out.c % 4 == 0, "bn_act_pool_fwd: bad channel offset");
    DLIO_CHECK_ARG(p.pool_k == 1 || p.pool_k == 3, "bn_act_pool_fwd: pool_k must be 1 or 3");
    DLIO_CHECK_ARG(p.pool_k == 3 || (y.h == out.h && y.w == out.w), "bn_act_pool_fwd: extent mismatch");
    DLIO_CHECK_ARG(p.res_mode == 0 || (res_ptr && res.h == y.h && res.w == y.w && res.c >= p.c_off + y.c),
                   "bn_act_pool_fwd: bad residual");
    DLIO_CHECK_ARG((scale == nullptr) == (shift == nullptr), "bn_act_pool_fwd: scale/shift");
    BnPool a;
    a.y = Geo(y); a.out = Geo(out); a.res = p.res_mode ? Geo(res) : Geo(y);
    a.yp = y_ptr; a.scale = scale; a.shift = shift; a.resp = res_ptr;
    a.res_mode = p.res_mode; a.relu = p.relu; a.pk = p.pool_k; a.sh = p.pool_sh; a.sw = p.pool_sw;
    a.c_off = p.c_off; a.cg = y.c / 4;
    a.out_hi = out_hi; a.out_lo = out_lo; a.idx = pool_idx; a.ymax = pool_ymax;
    a.out_h2 = (__half *)out_h2; a.out_bound = out_bound;
    a.group = p.out_group == 2 ? 2 : 1;
    a.zero_tail = p.zero_tail ? 1 : 0;
    DLIO_CHECK_ARG(!a.zero_tail || (p.pool_k == 1 && p.c_off == 0), "bn_act_pool_fwd: zero_tail needs pool_k == 1 and c_off == 0");
    DLIO_CHECK_ARG(a.group == 1 || (out_h2 && a.out.wp % 2 == 0), "bn_act_pool_fwd: the pixel-pair layout needs fp16 planes and an even padded width");
    long long total = (long long)a.out.n * a.out.hp * a.out.wp * a.cg;
    const int block = block_for_cg(a.cg);
    cudaStream_t st = (cudaStream_t)stream;
    if (a.pk == 3 && a.res_mode == 0 && (a.sh == 1 || a.sh == 2) && (a.sw == 1 || a.sw == 2) &&
        (long long)a.out.n * a.out.hp * a.out.wp < (1LL << 31) && pool_tma_enabled() && a.cg <= 256 &&
        (((uintptr_t)y_ptr) & 15) == 0) {
        // bulk-copy row ring (see bn_pool3_fwd_tma_kernel)
        const int WS = a.cg >= 256 ? 1 : 256 / a.cg;
        const int nseg = (a.out.w + WS - 1) / WS;
        const int wcols = (WS - 1) * a.sw + 3;
        const size_t smem = (size_t)6 * wcols * y.c * sizeof(float);
        const long long total = (long long)a.out.n * nseg * a.out.h;
#define DLIO_POOL3_TMA(SH_, SW_)                                                                              \
    do {                                                                                                      \
        auto kern = a.relu ? bn_pool3_fwd_tma_kernel<SH_, SW_, true> : bn_pool3_fwd_tma_kernel<SH_, SW_, false>; \
        DLIO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        int per_sm = 0;                                                                                       \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, a.cg * WS, smem) != cudaSuccess || per_sm < 1) per_sm = 1; \
        long long grid = 148LL * per_sm;                                                                      \
        if (grid > total) grid = total;                                                                       \
        const long long upc = (total + grid - 1) / grid;                                                      \
        grid = (total + upc - 1) / upc;                                                                       \
        kern<<<(unsigned)grid, a.cg * WS, smem, st>>>(a, WS, nseg, upc);                                      \
    } while (0)
        if (smem <= 200 * 1024) {
            if (a.sh == 1 && a.sw == 2) DLIO_POOL3_TMA(1, 2);
            else if (a.sh == 2 && a.sw == 2) DLIO_POOL3_TMA(2, 2);
            else if (a.sh == 1 && a.sw == 1) DLIO_POOL3_TMA(1, 1);
            else DLIO_POOL3_TMA(2, 1);
            DLIO_LAUNCH_CHECK();
            return DLIO_OK;
        }
#undef DLIO_POOL3_TMA
    }
    if (a.pk == 3 && a.res_mode == 0 && (a.sh == 1 || a.sh == 2) && (a.sw == 1 || a.sw == 2) &&
        (long long)a.out.n * a.out.hp * a.out.wp < (1LL << 31)) {
        const int rows = a.out.n * a.out.hp;
#define DLIO_POOL3(SH_, SW_)                                                              \
    do {                                                                                  \
        auto kern = a.relu ? bn_pool3_fwd_kernel<SH_, SW_, true> : bn_pool3_fwd_kernel<SH_, SW_, false>; \
        const int grid = resident_grid(kern, block);                                      \
        a.segs = row_segments(rows, a.out.wp, block / a.cg, grid);                        \
        kern<<<grid, block, 0, st>>>(a);                                                  \
    } while (0)
        if (a.sh == 1 && a.sw == 2) DLIO_POOL3(1, 2);
        else if (a.sh == 2 && a.sw == 2) DLIO_POOL3(2, 2);
        else if (a.sh == 1 && a.sw == 1) DLIO_POOL3(1, 1);
        else DLIO_POOL3(2, 1);
#undef DLIO_POOL3
    } else if (a.pk == 1 && (long long)a.out.n * a.out.hp * a.out.wp < (1LL << 31) && g_apply_rows) {
        const int grid = resident_grid(bn_apply_rows_kernel, block);
        a.segs = row_segments(a.out.n * a.out.hp, a.out.wp, block / a.cg, grid);
        bn_apply_rows_kernel<<<grid, block, 0, st>>>(a);
    } else {
        DLIO_CHECK_ARG(!pool_ymax, "bn_act_pool_fwd: pool_ymax needs a 3x3 pool without a residual");
        bn_act_pool_fwd_kernel<<<grid_for(total, block, 16), block, 0, st>>>(a);
    }
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_bn_act_pool_bwd_reduce(dlio_tensor4 y, const float *y_ptr, const float *scale,
                                           const float *shift, const float *mean, const float *invstd,
                                           dlio_tensor4 res, const float *res_ptr, dlio_bnpool p, int grad_src,
                                           dlio_tensor4 dout, const float *dout_ptr, int ld_dout,
                                           const uint8_t *pool_idx, float *dz, float *dres, int dres_c,
                                           int dres_accumulate, double *sums, int sums_absmax, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(valid_t4(y) && y_ptr && dout_ptr && (dz || sums), "bn_act_pool_bwd_reduce: bad argument");
    int rc = check_cg(y.c, "bn_act_pool_bwd_reduce");
    if (rc) return rc;
    DLIO_CHECK_ARG(grad_src == DLIO_GRAD_AVG || valid_t4(dout), "bn_act_pool_bwd_reduce: bad dout");
    DLIO_CHECK_ARG(p.pool_k == 1 || (p.pool_k == 3 && pool_idx), "bn_act_pool_bwd_reduce: pooling needs pool_idx");
    DLIO_CHECK_ARG(p.res_mode != 1 || res_ptr, "bn_act_pool_bwd_reduce: residual pointer missing");
    BnBwd a;
    a.y = Geo(y); a.res = p.res_mode == 1 ? Geo(res) : Geo(y); a.dout = grad_src == DLIO_GRAD_AVG ? Geo(y) : Geo(dout);
    a.yp = y_ptr; a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.resp = res_ptr;
    a.doutp = dout_ptr;
    a.res_mode = p.res_mode; a.relu = p.relu; a.pk = grad_src == DLIO_GRAD_AVG ? 1 : p.pool_k;
    a.sh = p.pool_sh; a.sw = p.pool_sw; a.c_off = p.c_off; a.cg = y.c / 4;
    a.grad_src = grad_src; a.ld_dout = ld_dout;
    a.pooled_h = dout.h; a.pooled_w = dout.w;
    a.idx = pool_idx; a.dz = dz; a.dres = dres; a.dres_c = dres_c; a.dres_acc = dres_accumulate; a.sums = sums;
    a.want_absmax = sums_absmax;
    DLIO_CHECK_ARG(!sums_absmax || (sums && y.c % 32 == 0), "bn_act_pool_bwd_reduce: sums_absmax needs sums and c %% 32 == 0");
    int block = block_for_cg(a.cg);
    long long total = (long long)y.n * y.h * y.w * a.cg;
    const int grid = grid_for(total, block, 8);
    cudaStream_t st = (cudaStream_t)stream;
    const bool fast = a.pk == 3 && a.res_mode == 0 && !a.dres && a.sums && a.dz && y.c % 32 == 0 &&
                      (long long)y.n * y.h * y.w < (1LL << 31);
    if (fast) {
        const int rows = y.n * y.h;
#define DLIO_RED3(SH_, SW_)                                                               \
    do {                                                                                  \
        const int g2 = resident_grid(bn_pool3_bwd_reduce_kernel<SH_, SW_>, block);        \
        a.segs = row_segments(rows, y.w, block / a.cg, g2);                               \
        bn_pool3_bwd_reduce_kernel<SH_, SW_><<<g2, block, 0, st>>>(a);                    \
    } while (0)
        if (a.sh == 1 && a.sw == 2) DLIO_RED3(1, 2);
        else if (a.sh == 2 && a.sw == 2) DLIO_RED3(2, 2);
        else if (a.sh == 1 && a.sw == 1) DLIO_RED3(1, 1);
        else if (a.sh == 2 && a.sw == 1) DLIO_RED3(2, 1);
        else {
            set_error("bn_act_pool_bwd_reduce: unsupported pool stride %dx%d", a.sh, a.sw);
            return DLIO_ERR_INVALID;
        }
        DLIO_LAUNCH_CHECK();
        return DLIO_OK;
    }
    const bool flat = a.pk == 1 && grad_src == DLIO_GRAD_DIRECT && a.res_mode != 1 && y.ph == 0 && y.pw == 0 &&
                      dout.ph == 0 && dout.pw == 0 && (long long)y.n * y.h * y.w < (1LL << 30) && g_apply_rows;
    if (flat) bn_bwd_reduce_flat_kernel<<<grid_for((total + 3) / 4, block, reduce_blocks_per_sm(y.c)), block, 0, st>>>(a);
    else if (a.pk == 1) bn_act_pool_bwd_reduce_kernel<0, 0><<<grid, block, 0, st>>>(a);
    else if (a.sh == 1 && a.sw == 2) bn_act_pool_bwd_reduce_kernel<1, 2><<<grid, block, 0, st>>>(a);
    else if (a.sh == 2 && a.sw == 2) bn_act_pool_bwd_reduce_kernel<2, 2><<<grid, block, 0, st>>>(a);
    else if (a.sh == 1 && a.sw == 1) bn_act_pool_bwd_reduce_kernel<1, 1><<<grid, block, 0, st>>>(a);
    else if (a.sh == 2 && a.sw == 1) bn_act_pool_bwd_reduce_kernel<2, 1><<<grid, block, 0, st>>>(a);
    else {
        set_error("bn_act_pool_bwd_reduce: unsupported pool stride %dx%d", a.sh, a.sw);
        return DLIO_ERR_INVALID;
    }
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

// dz != nullptr: dz read from memory.  dz == nullptr: un-pooled on the fly from (dout, pool_idx) through the 3x3 pool p.
static int bn_bwd_apply_impl(dlio_tensor4 y, const float *y_ptr, const float *dz, const float *shift, int post_relu,
                             const dlio_bnpool *p,
                             dlio_tensor4 dout_t, const float *dout, const uint8_t *pool_idx, const double *sums,
                             long long count, const float *scale, const float *mean, const float *invstd,
                             int pre_relu, int batch_stats, dlio_tensor4 dy_t, float *dy_hi, float *dy_lo,
                             void *dy_h2, float *dy_bound, float *dgamma, float *dbeta, double *dbias_sums,
                             void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(valid_t4(y) && valid_t4(dy_t) && y_ptr && (dz || (p && dout && pool_idx)) && (dy_hi || dy_h2),
                   "bn_bwd_apply: bad argument");
    DLIO_CHECK_ARG(!dy_h2 || (dy_bound && sums && (((uintptr_t)dy_h2) & 15) == 0), "bn_bwd_apply: fp16 planes need sums[2C] and dy_bound");
    DLIO_CHECK_ARG(!dy_h2 || y.c % 32 == 0, "bn_bwd_apply: fp16 planes need c %% 32 == 0 (whole warps in the bound reduction)");
    DLIO_CHECK_ARG(dy_hi || !dy_lo, "bn_bwd_apply: dy_lo without dy_hi");
    // dy_t.h == y.h: dy on y's own grid.  dy_t.h == input height of an H-stride-2 convolution: y row i goes to dy row
    // 2 i and the odd rows are zero (the backward of that convolution runs as a stride-1 one over the input grid)
    const int hs = dy_t.h == y.h ? 1 : 2;
    // dy_t.c > y.c (dz variant only): the extra channels are written as zeros
    DLIO_CHECK_ARG(dy_t.n == y.n && (hs == 1 || (dy_t.h - 1) / 2 + 1 == y.h) && dy_t.w == y.w &&
                       (dy_t.c == y.c || (dz && dy_t.c > y.c && dy_t.c % 4 == 0)),
                   "bn_bwd_apply: dy geometry");
    int rc = check_cg(y.c, "bn_bwd_apply");
    if (rc) return rc;
    BnApply a;
    a.y = Geo(y); a.dy = Geo(dy_t);
    a.yp = y_ptr; a.dz = dz; a.scale = scale; a.mean = mean; a.invstd = invstd; a.sums = sums;
    a.shift = shift; a.post_relu = post_relu;
    DLIO_CHECK_ARG(!post_relu || (dz && (shift || !scale)), "bn_bwd_apply: post_relu needs dz and the BN shift");
    a.count = (double)count; a.pre_relu = pre_relu; a.batch_stats = batch_stats; a.cg = y.c / 4;
    a.dy_hi = dy_hi; a.dy_lo = dy_lo; a.dgamma = dgamma; a.dbeta = dbeta; a.dbias = dbias_sums;
    a.dy_h2 = (__half *)dy_h2; a.dy_bound = dy_bound; a.hs = hs;
    a.doutp = nullptr; a.idx = nullptr; a.pooled_h = a.pooled_w = a.c_off = 0;
    int block = block_for_cg(a.cg);
    DLIO_CHECK_ARG((long long)a.dy.n * a.dy.hp * a.dy.wp < (1LL << 31), "bn_bwd_apply: tensor too large");
    cudaStream_t st = (cudaStream_t)stream;
#define DLIO_APPLY(SH_, SW_)                                                        \
    do {                                                                            \
        const int grid = resident_grid(bn_bwd_apply_kernel<SH_, SW_>, block);       \
        a.segs = row_segments(a.dy.n * a.dy.hp, a.dy.wp, block / a.cg, grid);       \
        bn_bwd_apply_kernel<SH_, SW_><<<grid, block, 0, st>>>(a);                   \
    } while (0)
    if (dz) {
        DLIO_APPLY(0, 0);
    } else {
        DLIO_CHECK_ARG(p->pool_k == 3 && p->relu == 0 && p->res_mode == 0 && hs == 1 && (p->pool_sh == 1 || p->pool_sh == 2) &&
                           (p->pool_sw == 1 || p->pool_sw == 2),
                       "bn_pool_bwd_apply: needs a 3x3 pool (strides 1 or 2) directly after the BN, no ReLU / residual between");
        DLIO_CHECK_ARG(valid_t4(dout_t) && dout_t.n == y.n && p->c_off % 4 == 0 && p->c_off + y.c <= dout_t.c && dout_t.c % 4 == 0,
                       "bn_pool_bwd_apply: bad dout descriptor");
        a.dout = Geo(dout_t); a.doutp = dout; a.idx = pool_idx; a.pooled_h = dout_t.h; a.pooled_w = dout_t.w;
        a.c_off = p->c_off;
        {
            // bulk-copy row rings (bn_pool_bwd_apply_tma_kernel): dout carries exactly this layer's channels
            const int WT = a.cg >= 256 ? 1 : 256 / a.cg, WSI = APPLY_U * WT;
            const int ocols = WSI / p->pool_sw + 2;
            const size_t smem = (size_t)apply_yslots(p->pool_sh) * WSI * y.c * 4 +
                                (size_t)apply_pslots(p->pool_sh) * ocols * y.c * 5;
            if (pool_tma_enabled() && dout_t.c == y.c && p->c_off == 0 && y.c % 16 == 0 && (a.cg * WT) % 32 == 0 &&
                a.cg * WT <= 256 && smem <= 200 * 1024 && a.y.w % 1 == 0 &&
                ((((uintptr_t)y_ptr) | ((uintptr_t)dout) | ((uintptr_t)pool_idx)) & 15) == 0) {
                const int nseg = (a.y.w + WSI - 1) / WSI;
                const long long total = (long long)a.y.n * nseg * a.y.h;
#define DLIO_APPLY_TMA(SH_, SW_)                                                                                  \
    do {                                                                                                          \
        auto kern = bn_pool_bwd_apply_tma_kernel<SH_, SW_>;                                                       \
        DLIO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        int per_sm = 0;                                                                                           \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, a.cg * WT, smem) != cudaSuccess || per_sm < 1) per_sm = 1; \
        long long grid = 148LL * per_sm;                                                                          \
        if (grid > total) grid = total;                                                                           \
        const long long upc = (total + grid - 1) / grid;                                                          \
        grid = (total + upc - 1) / upc;                                                                           \
        kern<<<(unsigned)grid, a.cg * WT, smem, st>>>(a, WT, nseg, upc);                                          \
    } while (0)
                if (p->pool_sh == 1 && p->pool_sw == 2) DLIO_APPLY_TMA(1, 2);
                else if (p->pool_sh == 2 && p->pool_sw == 2) DLIO_APPLY_TMA(2, 2);
                else if (p->pool_sh == 1 && p->pool_sw == 1) DLIO_APPLY_TMA(1, 1);
                else DLIO_APPLY_TMA(2, 1);
#undef DLIO_APPLY_TMA
                DLIO_LAUNCH_CHECK();
                return DLIO_OK;
            }
        }
        if (p->pool_sh == 1 && p->pool_sw == 2) DLIO_APPLY(1, 2);
        else if (p->pool_sh == 2 && p->pool_sw == 2) DLIO_APPLY(2, 2);
        else if (p->pool_sh == 1 && p->pool_sw == 1) DLIO_APPLY(1, 1);
        else DLIO_APPLY(2, 1);
    }
#undef DLIO_APPLY
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_bn_bwd_apply(dlio_tensor4 y, const float *y_ptr, const float *dz, const double *sums,
                                 long long count, const float *scale, const float *mean, const float *invstd,
                                 int pre_relu, int batch_stats, dlio_tensor4 dy_t, float *dy_hi, float *dy_lo,
                                 void *dy_h2, float *dy_bound, float *dgamma, float *dbeta, double *dbias_sums,
                                 const float *shift, int post_relu, void *stream) {
    DLIO_CHECK_ARG(dz, "bn_bwd_apply: dz is NULL");
    return bn_bwd_apply_impl(y, y_ptr, dz, shift, post_relu, nullptr, y, nullptr, nullptr, sums, count, scale, mean, invstd, pre_relu,
                             batch_stats, dy_t, dy_hi, dy_lo, dy_h2, dy_bound, dgamma, dbeta, dbias_sums, stream);
}

extern "C" int dlio_bn_pool_bwd_apply(dlio_tensor4 y, const float *y_ptr, dlio_bnpool p, dlio_tensor4 dout,
                                      const float *dout_ptr, const uint8_t *pool_idx, const double *sums,
                                      long long count, const float *scale, const float *mean, const float *invstd,
                                      int pre_relu, int batch_stats, dlio_tensor4 dy_t, float *dy_hi, float *dy_lo,
                                      void *dy_h2, float *dy_bound, float *dgamma, float *dbeta,
                                      double *dbias_sums, void *stream) {
    return bn_bwd_apply_impl(y, y_ptr, nullptr, nullptr, 0, &p, dout, dout_ptr, pool_idx, sums, count, scale, mean, invstd,
                             pre_relu, batch_stats, dy_t, dy_hi, dy_lo, dy_h2, dy_bound, dgamma, dbeta, dbias_sums,
                             stream);
}

extern "C" int dlio_pool_bwd_sums(dlio_tensor4 dout, const float *dout_ptr, int c_off, int c, const float *ymax,
                                  const float *mean, const float *invstd, int windows, double *sums, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(valid_t4(dout) && dout.ph == 0 && dout.pw == 0 && dout_ptr && ymax && mean && invstd && sums &&
                       windows >= 1 && c_off % 4 == 0 && c_off + c <= dout.c && dout.c % 4 == 0,
                   "pool_bwd_sums: bad argument");
    int rc = check_cg(c, "pool_bwd_sums");
    if (rc) return rc;
    const int cg = c / 4, block = block_for_cg(cg);
    const long long total = (long long)dout.n * dout.h * dout.w * cg;
    DLIO_CHECK_ARG((long long)dout.n * dout.h * dout.w < (1LL << 30), "pool_bwd_sums: tensor too large");
    pool_bwd_sums_kernel<<<grid_for((total + 3) / 4, block, reduce_blocks_per_sm(c)), block, 0, (cudaStream_t)stream>>>(
        Geo(dout), dout_ptr, c_off, ymax, mean, invstd, cg, (float)windows, sums);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_f64_to_f32(const double *src, float *dst, int n, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(src && dst && n > 0, "f64_to_f32: bad argument");
    f64_to_f32_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

static int launch_spatial(SpatialRed &a, cudaStream_t st) {
    int rc = check_cg(a.x.c, "spatial_reduce");
    if (rc) return rc;
    a.cg = a.x.c / 4;
    int block = block_for_cg(a.cg);
    int rows = block / a.cg;
    int hw = a.x.h * a.x.w;
    int split = ceil_div(hw, rows * 64);  // ~64 pixels per thread
    int cap = ceil_div(148 * 8, a.x.n);
    if (split > cap) split = cap;
    if (split < 1) split = 1;
    a.use_atomic = split > 1;
    if (a.use_atomic) {
        zero_strided_kernel<<<ceil_div((long long)a.x.n * a.x.c, 256), 256, 0, st>>>(a.out + a.c_off, a.x.n, a.x.c, a.ld_out);
        DLIO_LAUNCH_CHECK();
    }
    spatial_reduce_kernel<<<dim3(split, a.x.n), block, 0, st>>>(a);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_spatial_mean_fwd(dlio_tensor4 x, const float *x_ptr, const float *scale, const float *shift,
                                     int relu, float *out, int ld_out, int c_off, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(valid_t4(x) && x_ptr && out && ld_out >= c_off + x.c, "spatial_mean_fwd: bad argument");
    SpatialRed a;
    a.x = Geo(x); a.other = a.x; a.xp = x_ptr; a.otherp = nullptr; a.scale = scale; a.shift = shift;
    a.relu = relu; a.mode = 0; a.ld_out = ld_out; a.c_off = c_off; a.inv = 1.f / (float)(x.h * x.w); a.out = out;
    return launch_spatial(a, (cudaStream_t)stream);
}

extern "C" int dlio_spatial_dot(dlio_tensor4 a_t, const float *a_ptr, dlio_tensor4 b_t, const float *b_ptr,
                                float *out, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(valid_t4(a_t) && valid_t4(b_t) && a_ptr && b_ptr && out, "spatial_dot: bad argument");
    DLIO_CHECK_ARG(a_t.n == b_t.n && a_t.h == b_t.h && a_t.w == b_t.w && a_t.c == b_t.c, "spatial_dot: geometry mismatch");
    SpatialRed a;
    a.x = Geo(a_t); a.other = Geo(b_t); a.xp = a_ptr; a.otherp = b_ptr; a.scale = nullptr; a.shift = nullptr;
    a.relu = 0; a.mode = 1; a.ld_out = a_t.c; a.c_off = 0; a.inv = 1.f; a.out = out;
    return launch_spatial(a, (cudaStream_t)stream);
}

extern "C" int dlio_channel_scale_fwd(dlio_tensor4 x, const float *x_ptr, const float *gate, dlio_tensor4 out,
                                      float *out_hi, float *out_lo, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(valid_t4(x) && valid_t4(out) && x_ptr && gate && out_hi, "channel_scale_fwd: bad argument");
    DLIO_CHECK_ARG(x.n == out.n && x.h == out.h && x.w == out.w && x.c == out.c && x.c % 4 == 0, "channel_scale_fwd: geometry");
    Geo xg(x), og(out);
    long long total = (long long)og.n * og.hp * og.wp * (x.c / 4);
    channel_scale_fwd_kernel<<<grid_for(total, block_for_cg(x.c / 4), 16), block_for_cg(x.c / 4), 0, (cudaStream_t)stream>>>(xg, x_ptr, gate, og, out_hi, out_lo, x.c / 4);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_channel_scale_bwd(const float *dout, const float *gate, const float *dmean, int n, int hw,
                                      int c, float *dx, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(dout && gate && dx && c % 4 == 0 && n > 0 && hw > 0, "channel_scale_bwd: bad argument");
    long long total = (long long)n * hw * (c / 4);
    channel_scale_bwd_kernel<<<grid_for(total, block_for_cg(c / 4), 16), block_for_cg(c / 4), 0, (cudaStream_t)stream>>>(dout, gate, dmean, 1.f / (float)hw, dx, n, hw, c / 4);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_axpby(const float *a, float alpha, const float *b, float beta, float *out, long long n,
                          void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(a && b && out && n > 0, "axpby: bad argument");
    DLIO_CHECK_ARG((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0, "axpby: pointers must be 16-byte aligned");
    axpby_kernel<<<grid_for((n + 3) / 4, 256, 16), 256, 0, (cudaStream_t)stream>>>(a, alpha, b, beta, out, n);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
extern "C" int dlio_sum_mid(const float *x, float *out, long long a, int t, int c, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(x && out && a > 0 && t > 0 && c > 0, "sum_mid: bad argument");
    sum_mid_kernel<<<ceil_div(a * c, 256), 256, 0, (cudaStream_t)stream>>>(x, out, a, t, c);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
extern "C" int dlio_copy2d(const float *src, long long lds, float *dst, long long ldd, long long rows, int cols,
                           void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= cols, "copy2d: bad argument");
    copy2d_kernel<<<grid_for(rows * cols, 256, 16), 256, 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, rows, cols);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
extern "C" int dlio_mul(const float *a, const float *b, float *out, long long n, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(a && b && out && n > 0, "mul: bad argument");
    mul_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
extern "C" int dlio_dropout_mask(float *mask, long long n, float p, unsigned long long seed,
                                 const unsigned long long *seed_epoch, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(mask && n > 0 && p >= 0.f && p < 1.f, "dropout_mask: bad argument");
    dropout_mask_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(mask, n, p, seed, seed_epoch);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
