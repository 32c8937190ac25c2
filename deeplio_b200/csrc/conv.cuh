// Argument block shared by the generic (conv_simt.cu) and tcgen05 (conv_tc.cu) convolution kernels.
#pragma once
#include "common.cuh"

namespace dlio {

struct ConvArgs {
    Geo x, y;            // x: conv input geometry, y: conv output geometry (dy for backward)
    int kh, kw, sh, sw, ph, pw;
    int cin, cout;
    int act;
    const float *x_hi, *x_lo;
    const float *w_hi, *w_lo;
    // fp16 split mode (x_h2 != nullptr): packed hi|lo planes of halves and the device-resident bounds that
    // define their power-of-two scales (common.cuh, "fp16 operand split"); x_hi .. w_lo are unused then
    const __half *x_h2 = nullptr, *w_h2 = nullptr;
    const float *x_bound = nullptr, *w_bound = nullptr;
    const float *bias;
    float *out;          // y (fwd), dx (dgrad), dw (wgrad)
    double *stats;
    Geo o;               // geometry of `out` for dgrad (dx)
    long long p_chunk;   // wgrad: pixels per z-slice
    int hdec = 1;        // conv_tc_fwd: 2 = store / count only the even rows of the (stride-1) result
    int accum = 0;       // dgrad: add to `out` instead of overwriting it
    int fold = 0;        // "folded split" first layer (conv_s2d.cu): x_h2 is ONE plane of 64-half rows [hi|lo per pixel],
                         // cin == 64; forward: w_h2 planes are [w_hi | 0] and [w_lo | w_hi]; wgrad: w_h2 (= dy) rows are
                         // [64 hi | 64 lo] per output pixel, cout / 64 pixels per row
};

// conv_tc.cu: return 1 if the tcgen05 kernel took the problem, 0 if it does not apply, < 0 on error
int conv_tc_fwd(const ConvArgs &a, int prof_kind, cudaStream_t st);
int conv_tc_wgrad(const ConvArgs &a, cudaStream_t st);

}  // namespace dlio
