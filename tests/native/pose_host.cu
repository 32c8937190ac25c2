// Host-side self-check of the pose kernels' arithmetic: compiles the per-sample functions of
// deeplio_b200/csrc/pose_math.cuh (the same source the CUDA kernels are built from) for the CPU and runs them on
// binary input from stdin.  tests/test_host_logic_cpu.py compares the result with oracle/pose_oracle.py and with the
// reference-generated fixture, so the math is verified in the CPU tier too (no GPU, no CUDA API call here).
//
// stdin:  int32 mode (0 chain, 1 ground truth)
//   chain: int32 B, S;  float x[B*S*3], w[B*S*3], gx[B*S*3], gq[B*S*4]
//          -> stdout float ox[B*S*3], oq[B*S*4], dx[B*S*3], dw[B*S*3], then int32 status
//   gt:    int32 B, F, S; int32 comb[2*S]; float gts[B*F*15]  -> float f2f[B*S*6], f2g[B*S*7], int32 status
#include <cstdio>
#include <vector>

#include "../../deeplio_b200/csrc/pose_math.cuh"

template <typename T>
static bool rd(std::vector<T> &v) { return fread(v.data(), sizeof(T), v.size(), stdin) == v.size(); }
template <typename T>
static void wr(const std::vector<T> &v) { fwrite(v.data(), sizeof(T), v.size(), stdout); }

int main() {
    std::vector<int> hdr(1);
    if (!rd(hdr)) return 2;
    if (hdr[0] == 0) {
        std::vector<int> bs(2);
        if (!rd(bs)) return 2;
        const int B = bs[0], S = bs[1];
        if (S > dlio::CHAIN_MAX_S) return 3;
        std::vector<float> x(B * S * 3), w(B * S * 3), gx(B * S * 3), gq(B * S * 4);
        if (!rd(x) || !rd(w) || !rd(gx) || !rd(gq)) return 2;
        std::vector<float> ox(B * S * 3), oq(B * S * 4), dx(B * S * 3), dw(B * S * 3);
        std::vector<int> status(1, 0);
        for (int b = 0; b < B; ++b) {
            const size_t o = (size_t)b * S;
            status[0] |= dlio::chain_fwd_sample(&x[o * 3], &w[o * 3], S, &ox[o * 3], &oq[o * 4]);
            dlio::chain_bwd_sample(&x[o * 3], &w[o * 3], S, &gx[o * 3], &gq[o * 4], &dx[o * 3], &dw[o * 3]);
        }
        wr(ox); wr(oq); wr(dx); wr(dw); wr(status);
        return 0;
    }
    std::vector<int> bfs(3);
    if (!rd(bfs)) return 2;
    const int B = bfs[0], F = bfs[1], S = bfs[2];
    std::vector<int> comb(2 * S);
    std::vector<float> gts((size_t)B * F * 15);
    if (!rd(comb) || !rd(gts)) return 2;
    std::vector<float> f2f((size_t)B * S * 6), f2g((size_t)B * S * 7);
    std::vector<int> status(1, 0);
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < S; ++s) {
            const float *g0 = &gts[(size_t)b * F * 15];
            status[0] |= dlio::gt_relative_sample(g0 + comb[2 * s] * 15, g0 + comb[2 * s + 1] * 15, g0,
                                                  &f2f[((size_t)b * S + s) * 6], &f2g[((size_t)b * S + s) * 7]);
        }
    wr(f2f); wr(f2g); wr(status);
    return 0;
}
