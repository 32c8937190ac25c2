// Per-point / per-pixel math of the LiDAR preprocessing kernels (scan.cu), host + device so that the CPU test tier
// runs the very same arithmetic (tests/native/scan_host.cu).  Follows deeplio/common/laserscan.py:122-248 and
// deeplio/datasets/kitti.py:83-97 in float32, as numpy evaluates those expressions on float32 arrays.
#pragma once
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace dlio {

struct ScanGeom {
    int H, W;
    float fov_down_abs, fov;     // radians, rounded to float32 like numpy's weak python scalars
    float min_depth, max_depth;
};

#ifdef __CUDA_ARCH__
#define DLIO_FMUL(a, b) __fmul_rn(a, b)
#define DLIO_FADD(a, b) __fadd_rn(a, b)
#else
#define DLIO_FMUL(a, b) ((a) * (b))
#define DLIO_FADD(a, b) ((a) + (b))
#endif

// depth as np.linalg.norm(points, 2, axis=1) gives it for float32 rows: sqrt((x*x + y*y) + z*z), no fused multiply-add
__host__ __device__ __forceinline__ float scan_depth(float x, float y, float z) {
    return sqrtf(DLIO_FADD(DLIO_FADD(DLIO_FMUL(x, x), DLIO_FMUL(y, y)), DLIO_FMUL(z, z)));
}
// open_scan's filter (laserscan.py:86-91): points with depth > max_depth or < min_depth are dropped (NaNs too here)
__host__ __device__ __forceinline__ bool scan_keep(float depth, const ScanGeom &g) {
    return depth >= g.min_depth && depth <= g.max_depth;
}
// pixel of a point (laserscan.py:141-162)
__host__ __device__ __forceinline__ int scan_pixel(float x, float y, float z, float depth, const ScanGeom &g) {
    const float pi = 3.14159265358979323846f;
    const float yaw = -atan2f(y, x);
    const float pitch = asinf(z / depth);
    float px = DLIO_FMUL(0.5f, DLIO_FADD(yaw / pi, 1.0f));
    float py = 1.0f - DLIO_FADD(pitch, g.fov_down_abs) / g.fov;
    px = floorf(DLIO_FMUL(px, (float)g.W));
    py = floorf(DLIO_FMUL(py, (float)g.H));
    px = fmaxf(0.f, fminf((float)(g.W - 1), px));
    py = fmaxf(0.f, fminf((float)(g.H - 1), py));
    return (int)py * g.W + (int)px;
}
// z-buffer key: nearest point wins, lower index on exact depth ties
__host__ __device__ __forceinline__ unsigned long long scan_key(float depth, unsigned idx) {
#ifdef __CUDA_ARCH__
    return ((unsigned long long)__float_as_uint(depth) << 32) | idx;
#else
    union { float f; unsigned u; } c;
    c.f = depth;
    return ((unsigned long long)c.u << 32) | idx;
#endif
}
constexpr unsigned long long SCAN_EMPTY = ~0ULL;

struct ScanPix {
    float x, y, z, r, rem;     // zeros when the pixel is empty (the reference initialises its images with zeros)
};
__host__ __device__ __forceinline__ ScanPix scan_fetch(const float *points4, const unsigned long long *zbuf, int pix) {
    ScanPix p = {0.f, 0.f, 0.f, 0.f, 0.f};
    const unsigned long long k = zbuf[pix];
    if (k == SCAN_EMPTY) return p;
    const float *q = points4 + (size_t)(unsigned)(k & 0xFFFFFFFFu) * 4;
    p.x = q[0]; p.y = q[1]; p.z = q[2]; p.rem = q[3];
    p.r = scan_depth(p.x, p.y, p.z);
    return p;
}
// normal at an interior pixel from its four neighbours (laserscan.py:215-248)
__host__ __device__ inline void scan_normal(const ScanPix &c, const ScanPix &t, const ScanPix &l, const ScanPix &b,
                                            const ScanPix &r, float *n) {
    const ScanPix *nb[4] = {&t, &l, &b, &r};
    float d[4][3];
    for (int i = 0; i < 4; ++i) {
        const float w = expf(DLIO_FMUL(-0.8f, fabsf(nb[i]->r - c.r)));
        d[i][0] = DLIO_FMUL(w, nb[i]->x - c.x);
        d[i][1] = DLIO_FMUL(w, nb[i]->y - c.y);
        d[i][2] = DLIO_FMUL(w, nb[i]->z - c.z);
    }
    float s[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < 4; ++i) {        // cross(t,l) + cross(l,b) + cross(b,r) + cross(r,t)
        const float *u = d[i], *v = d[(i + 1) & 3];
        const float c0 = DLIO_FADD(DLIO_FMUL(u[1], v[2]), -DLIO_FMUL(u[2], v[1]));
        const float c1 = DLIO_FADD(DLIO_FMUL(u[2], v[0]), -DLIO_FMUL(u[0], v[2]));
        const float c2 = DLIO_FADD(DLIO_FMUL(u[0], v[1]), -DLIO_FMUL(u[1], v[0]));
        s[0] = DLIO_FADD(s[0], c0);
        s[1] = DLIO_FADD(s[1], c1);
        s[2] = DLIO_FADD(s[2], c2);
    }
    const float nn = sqrtf(DLIO_FADD(DLIO_FADD(DLIO_FMUL(s[0], s[0]), DLIO_FMUL(s[1], s[1])), DLIO_FMUL(s[2], s[2]))) + 1e-8f;
    n[0] = s[0] / nn; n[1] = s[1] / nn; n[2] = s[2] / nn;
}
// the 8 channels of one pixel (kitti.py:83-97): xyz / max_depth, remission, normal, range
__host__ __device__ inline void scan_channels(const float *points4, const unsigned long long *zbuf, int y, int x,
                                              const ScanGeom &g, float *ch) {
    const int pix = y * g.W + x;
    const ScanPix c = scan_fetch(points4, zbuf, pix);
    float n[3] = {0.f, 0.f, 0.f};
    if (y > 0 && y < g.H - 1 && x > 0 && x < g.W - 1)
        scan_normal(c, scan_fetch(points4, zbuf, pix - g.W), scan_fetch(points4, zbuf, pix - 1),
                    scan_fetch(points4, zbuf, pix + g.W), scan_fetch(points4, zbuf, pix + 1), n);
    ch[0] = c.x / g.max_depth; ch[1] = c.y / g.max_depth; ch[2] = c.z / g.max_depth; ch[3] = c.rem;
    ch[4] = n[0]; ch[5] = n[1]; ch[6] = n[2]; ch[7] = c.r;
}

// IMU windowing (kitti.py:317-343,366-368): first OXTS sample with ts >= t0
__host__ __device__ __forceinline__ int imu_lower_bound(const double *ts, int m, double t0) {
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ts[mid] < t0) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

}  // namespace dlio
