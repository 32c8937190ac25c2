// Shared helpers for the deeplio_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deeplio_b200.h"

namespace dlio {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// kernel classes timed by dlio_profile_* (see deeplio_b200.h)
int prof_begin(int kind, cudaStream_t st);
void prof_end(int idx, cudaStream_t st);
struct ProfScope {
    int idx;
    cudaStream_t st;
    ProfScope(int kind, cudaStream_t s) : idx(prof_begin(kind, s)), st(s) {}
    ~ProfScope() { prof_end(idx, st); }
};

#define DLIO_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            dlio::set_error(__VA_ARGS__);    \
            return DLIO_ERR_INVALID;         \
        }                                    \
    } while (0)

#define DLIO_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            dlio::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return DLIO_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define DLIO_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        dlio::count_launch();                                                             \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            dlio::set_error("%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return DLIO_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

// geometry of a padded NHWC tensor
struct Geo {
    int n, h, w, c, ph, pw;
    int hp, wp;  // h + 2ph, w + 2pw
    __host__ __device__ Geo() {}
    __host__ __device__ Geo(const dlio_tensor4 &t)
        : n(t.n), h(t.h), w(t.w), c(t.c), ph(t.ph), pw(t.pw), hp(t.h + 2 * t.ph), wp(t.w + 2 * t.pw) {}
    // element offset of logical (n, y, x), channel 0
    __host__ __device__ __forceinline__ size_t off(int in, int y, int x) const {
        return (((size_t)in * hp + (y + ph)) * wp + (x + pw)) * (size_t)c;
    }
    __host__ __device__ size_t numel() const { return (size_t)n * hp * wp * c; }
    __host__ __device__ size_t pixels() const { return (size_t)n * h * w; }
    __host__ __device__ size_t padded_pixels() const { return (size_t)n * hp * wp; }
};

inline bool valid_t4(const dlio_tensor4 &t) {
    return t.n > 0 && t.h > 0 && t.w > 0 && t.c > 0 && t.ph >= 0 && t.pw >= 0;
}

__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
// TF32 operand split used by the tcgen05 kernels: the "hi" plane keeps the FULL fp32 value (kind::tf32 reads
// the top 19 bits of each 32-bit container, i.e. trunc_tf32(v)); the "lo" plane holds the residual
// v - trunc_tf32(v), itself rounded to TF32.  hi*hi + lo*hi + hi*lo then reproduces the fp32 product to ~2^-20.
__device__ __forceinline__ void tf32_split(float v, float &hi, float &lo) {
    hi = v;
    lo = tf32_rna(v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u));
}
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
// fp32 value of a (possibly split) tensor: the hi plane already holds it
__device__ __forceinline__ float4 ld4_sum(const float *hi, const float *, size_t off) { return ld4(hi + off); }
__device__ __forceinline__ void st4_split(float *hi, float *lo, size_t off, const float4 &v) {
    if (lo) {
        float4 h, l;
        tf32_split(v.x, h.x, l.x);
        tf32_split(v.y, h.y, l.y);
        tf32_split(v.z, h.z, l.z);
        tf32_split(v.w, h.w, l.w);
        st4(hi + off, h);
        st4(lo + off, l);
    } else {
        st4(hi + off, v);
    }
}

// ---------------------------------------------------------------- fp16 operand split ("3xF16", conv_tc.cu)
// A tensor that feeds the fp16 tensor-core kernels is stored as ONE packed plane of halves, row-major over
// pixels (or output channels for weights): row r = [C hi halves | C lo halves], with
//     hi = fp16(s * v),   lo = fp16((s * v - hi) * 2^11),   s = f16_scale_from_bound(bound)
// where `bound` >= max |v| over the tensor lives in device memory next to it (one float, written by the
// producer's statistics kernel -- no host round trip).  s is a power of two, so scaling is exact; s * bound lies
// in (2^13, 2^14], far below the fp16 maximum, and values down to 2^-14 / s keep full 2 x 11-bit precision
// (smaller ones degrade gracefully to an ABSOLUTE error of 2^-36 / s, ~2^-50 of the tensor's bound).
// hi*hi + (lo*hi + hi*lo) * 2^-11 reproduces the fp32 product to ~2^-22.
__host__ __device__ __forceinline__ float f16_scale_from_bound(float bound) {
    if (!(bound > 0.f) || bound > 3.0e38f) return 1.f;
    int e;
    frexpf(bound, &e);          // bound = m * 2^e, m in [0.5, 1)
    int k = 14 - e;
    k = k > 80 ? 80 : (k < -80 ? -80 : k);
    return ldexpf(1.f, k);
}
__device__ __forceinline__ void f16_split(float v, __half &hi, __half &lo) {   // v already scaled
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
}
// stores 4 consecutive channels (c % 4 == 0) of packed row `row` (C channels per half-row).
// group == 2 ("pixel-pair layout"): a packed row holds TWO consecutive pixels, [p0 C hi | p1 C hi | p0 C lo | p1 C lo]
// -- the layout of a tensor whose consumer is a stride-2 convolution run on the tensor cores through the
// pixel-pair view [n, h, w/2, 2C] (conv_s2d.cu); `row` is still the pixel index (the padded width must be even).
__device__ __forceinline__ __half *h2_addr(__half *base, size_t row, int C, int c, int group) {
    return group == 2 ? base + (row >> 1) * (size_t)(4 * C) + (row & 1) * (size_t)C + c
                      : base + row * (size_t)(2 * C) + c;
}
// (packed conversions: cvt.rn.f16x2.f32 handles two channels per instruction; same round-to-nearest results as the
// scalar f16_split)
__device__ __forceinline__ unsigned pack_h2(__half a, __half b) {
    return (unsigned)__half_as_ushort(a) | ((unsigned)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void st4_h2(__half *base, size_t row, int C, int c, const float4 &v, float s, int group = 1) {
    const float x0 = v.x * s, x1 = v.y * s, x2 = v.z * s, x3 = v.w * s;
    const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
    const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((x0 - b01.x) * 2048.f, (x1 - b01.y) * 2048.f);
    const __half2 l23 = __floats2half2_rn((x2 - b23.x) * 2048.f, (x3 - b23.y) * 2048.f);
    __half *p = h2_addr(base, row, C, c, group);
    uint2 hi, lo;
    hi.x = *reinterpret_cast<const unsigned *>(&h01);
    hi.y = *reinterpret_cast<const unsigned *>(&h23);
    lo.x = *reinterpret_cast<const unsigned *>(&l01);
    lo.y = *reinterpret_cast<const unsigned *>(&l23);
    *reinterpret_cast<uint2 *>(p) = hi;
    *reinterpret_cast<uint2 *>(p + (group == 2 ? 2 * C : C)) = lo;
}
__device__ __forceinline__ void st4_h2_zero(__half *base, size_t row, int C, int c, int group = 1) {
    __half *p = h2_addr(base, row, C, c, group);
    *reinterpret_cast<uint2 *>(p) = make_uint2(0u, 0u);
    *reinterpret_cast<uint2 *>(p + (group == 2 ? 2 * C : C)) = make_uint2(0u, 0u);
}
// atomic max of non-negative floats (their bit patterns order like unsigned integers)
__device__ __forceinline__ void atomic_max_nonneg(float *addr, float v) {
    atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float act_apply(float v, int act) {
    switch (act) {
        case DLIO_ACT_RELU: return v > 0.f ? v : 0.f;
        case DLIO_ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
        case DLIO_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case DLIO_ACT_TANH: return tanhf(v);
        default: return v;
    }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    switch (act) {
        case DLIO_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case DLIO_ACT_LEAKY: return y > 0.f ? 1.f : 0.01f;
        case DLIO_ACT_SIGMOID: return y * (1.f - y);
        case DLIO_ACT_TANH: return 1.f - y * y;
        default: return 1.f;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace dlio
