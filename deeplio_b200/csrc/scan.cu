// LiDAR / IMU preprocessing on the device: what the reference's dataset workers do per
// frame in numpy on the CPU -- spherical range projection with a depth-sorted scatter (deeplio/common/laserscan.py
// :122-191), normal estimation (:215-248), image assembly / mean subtraction / channel selection (deeplio/datasets/
// kitti.py:83-97,345-364) and IMU windowing (:317-343,366-368).  ~120 k points per frame cannot feed thousands of
// frame pairs per second from Python; here a frame is two launches:
//   1. scan_zbuffer_kernel: every point computes its pixel and does ONE 64-bit atomicMin of (depth bits, index) --
//      the nearest point wins a pixel (what the reference's decreasing-depth argsort + scatter computes), no sort;
//   2. scan_image_kernel: every pixel fetches its winner and its four neighbours' winners (16-byte loads), computes
//      the normal and writes the selected channels, raw and mean-subtracted, planar [C, H, W] as the trainer wants.
// HBM traffic: 16 B per point read twice + 8 B of z-buffer per pixel + the image written once.
#include "scan_math.cuh"

namespace dlio {

__global__ void __launch_bounds__(256) scan_zbuffer_kernel(const float4 *__restrict__ pts, int n, ScanGeom g,
                                                           unsigned long long *zbuf) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        const float depth = scan_depth(p.x, p.y, p.z);
        if (!scan_keep(depth, g)) continue;
        atomicMin(zbuf + scan_pixel(p.x, p.y, p.z, depth, g), scan_key(depth, (unsigned)i));
    }
}

struct ScanOut {
    int nsel;
    int channel[8];
    float mean[8];
    float *org, *normed;       // [nsel, H, W] each, either may be NULL
    int *idx;                  // [H, W] original point index of the winner, -1 where empty (optional)
};
__global__ void __launch_bounds__(256) scan_image_kernel(const float *__restrict__ pts4,
                                                         const unsigned long long *__restrict__ zbuf, ScanGeom g, ScanOut o) {
    const int hw = g.H * g.W;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += gridDim.x * blockDim.x) {
        float ch[8];
        scan_channels(pts4, zbuf, pix / g.W, pix % g.W, g, ch);
        for (int k = 0; k < o.nsel; ++k) {
            const int c = o.channel[k];
            if (o.org) o.org[(size_t)k * hw + pix] = ch[c];
            if (o.normed) o.normed[(size_t)k * hw + pix] = ch[c] - o.mean[c];
        }
        if (o.idx) {
            const unsigned long long key = zbuf[pix];
            o.idx[pix] = key == SCAN_EMPTY ? -1 : (int)(unsigned)(key & 0xFFFFFFFFu);
        }
    }
}

struct ImuNorm {
    int on;
    double mean[6], std[6];
};
__global__ void imu_windows_kernel(const double *__restrict__ ts, const float *__restrict__ imu, int m,
                                   const double *__restrict__ velo_ts, int nwin, int T, ImuNorm nrm, float *out, int *valid) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nwin * T) return;
    const int w = id / T, k = id - w * T;
    const double t0 = velo_ts[w], t1 = velo_ts[w + 1];
    const int lo = imu_lower_bound(ts, m, t0);
    const int i = lo + k;
    const bool in = i < m && ts[i] < t1;
    for (int c = 0; c < 6; ++c) {
        double v = in ? (double)imu[(size_t)i * 6 + c] : 0.0;      // rows past the window are zero padding ...
        if (nrm.on) v = (v - nrm.mean[c]) / nrm.std[c];             // ... normalised like real samples (kitti.py:366-368)
        out[(size_t)id * 6 + c] = (float)v;
    }
    if (k == 0 && valid) valid[w] = (lo < m && ts[lo] < t1) ? 1 : 0;
}

}  // namespace dlio

using namespace dlio;

extern "C" size_t dlio_scan_scratch_bytes(int H, int W) { return (size_t)H * W * sizeof(unsigned long long); }

extern "C" int dlio_scan_project(const float *points4, int n_points, int H, int W, float fov_up_deg, float fov_down_deg,
                                 float min_depth, float max_depth, const int *channels, int n_channels,
                                 const float *mean8, void *scratch, float *out_org, float *out_normed, int *out_idx,
                                 void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG((points4 || n_points == 0) && n_points >= 0 && H > 0 && W > 0 && scratch &&
                       (out_org || out_normed || out_idx) && (((uintptr_t)points4) & 15) == 0 && (((uintptr_t)scratch) & 7) == 0,
                   "scan_project: bad argument");
    DLIO_CHECK_ARG(n_channels >= 0 && n_channels <= 8 && (n_channels == 0 || channels), "scan_project: 0 .. 8 channels");
    DLIO_CHECK_ARG(max_depth > 0.f && min_depth >= 0.f, "scan_project: bad depth range");
    ScanGeom g;
    g.H = H; g.W = W;
    const double up = (double)fov_up_deg / 180.0 * 3.14159265358979323846, down = (double)fov_down_deg / 180.0 * 3.14159265358979323846;
    g.fov_down_abs = (float)fabs(down);
    g.fov = (float)(fabs(down) + fabs(up));
    g.min_depth = min_depth; g.max_depth = max_depth;
    ScanOut o;
    o.nsel = n_channels;
    for (int k = 0; k < 8; ++k) {
        o.channel[k] = k < n_channels ? channels[k] : 0;
        DLIO_CHECK_ARG(o.channel[k] >= 0 && o.channel[k] < 8, "scan_project: channel index out of range");
        o.mean[k] = mean8 ? mean8[k] : 0.f;
    }
    o.org = out_org; o.normed = out_normed; o.idx = out_idx;
    cudaStream_t st = (cudaStream_t)stream;
    DLIO_CUDA(cudaMemsetAsync(scratch, 0xFF, dlio_scan_scratch_bytes(H, W), st));
    if (n_points > 0) {
        int blocks = ceil_div(n_points, 256);
        scan_zbuffer_kernel<<<blocks > 148 * 8 ? 148 * 8 : blocks, 256, 0, st>>>(
            reinterpret_cast<const float4 *>(points4), n_points, g, (unsigned long long *)scratch);
        DLIO_LAUNCH_CHECK();
    }
    int blocks = ceil_div((long long)H * W, 256);
    scan_image_kernel<<<blocks > 148 * 8 ? 148 * 8 : blocks, 256, 0, st>>>(points4, (const unsigned long long *)scratch, g, o);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_imu_windows(const double *ts, const float *imu, int m, const double *velo_ts, int n_frames, int T,
                                const float *mean6, const float *std6, float *out, int *valid, void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(ts && imu && velo_ts && out && m >= 0 && n_frames >= 2 && T > 0, "imu_windows: bad argument");
    DLIO_CHECK_ARG((mean6 == nullptr) == (std6 == nullptr), "imu_windows: mean and std go together");
    ImuNorm nrm;
    nrm.on = mean6 ? 1 : 0;
    for (int c = 0; c < 6; ++c) {
        nrm.mean[c] = mean6 ? (double)mean6[c] : 0.0;
        nrm.std[c] = std6 ? (double)std6[c] : 1.0;
    }
    const int total = (n_frames - 1) * T;
    imu_windows_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(ts, imu, m, velo_ts, n_frames - 1, T, nrm, out, valid);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
