// Host-side self-check of the LiDAR preprocessing kernels' arithmetic: compiles the per-point / per-pixel functions of
// deeplio_b200/csrc/scan_math.cuh (the source the CUDA kernels are built from) for the CPU; the 64-bit atomicMin of
// the z-buffer kernel becomes a serial minimum.  tests/test_host_logic_cpu.py compares the result with
// oracle/scan_oracle.py and the reference-generated fixture.
//
// stdin:  int32 n, H, W; float fov_up_deg, fov_down_deg, min_depth, max_depth; float points4[n*4]
// stdout: float image[H*W*8] (pixel-major, 8 channels), int32 idx[H*W]
#include <cstdio>
#include <vector>

#include "../../deeplio_b200/csrc/scan_math.cuh"

int main() {
    int hdr[3];
    float f[4];
    if (fread(hdr, 4, 3, stdin) != 3 || fread(f, 4, 4, stdin) != 4) return 2;
    const int n = hdr[0], H = hdr[1], W = hdr[2];
    std::vector<float> pts((size_t)n * 4);
    if (fread(pts.data(), 4, pts.size(), stdin) != pts.size()) return 2;
    dlio::ScanGeom g;
    g.H = H; g.W = W;
    const double up = (double)f[0] / 180.0 * 3.14159265358979323846, down = (double)f[1] / 180.0 * 3.14159265358979323846;
    g.fov_down_abs = (float)fabs(down);
    g.fov = (float)(fabs(down) + fabs(up));
    g.min_depth = f[2]; g.max_depth = f[3];
    std::vector<unsigned long long> zbuf((size_t)H * W, dlio::SCAN_EMPTY);
    for (int i = 0; i < n; ++i) {
        const float *p = &pts[(size_t)i * 4];
        const float depth = dlio::scan_depth(p[0], p[1], p[2]);
        if (!dlio::scan_keep(depth, g)) continue;
        const int pix = dlio::scan_pixel(p[0], p[1], p[2], depth, g);
        const unsigned long long key = dlio::scan_key(depth, (unsigned)i);
        if (key < zbuf[pix]) zbuf[pix] = key;
    }
    std::vector<float> img((size_t)H * W * 8);
    std::vector<int> idx((size_t)H * W);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            dlio::scan_channels(pts.data(), zbuf.data(), y, x, g, &img[((size_t)y * W + x) * 8]);
            const unsigned long long k = zbuf[(size_t)y * W + x];
            idx[(size_t)y * W + x] = k == dlio::SCAN_EMPTY ? -1 : (int)(unsigned)(k & 0xFFFFFFFFu);
        }
    fwrite(img.data(), 4, img.size(), stdout);
    fwrite(idx.data(), 4, idx.size(), stdout);
    return 0;
}
