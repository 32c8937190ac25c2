// Multi-layer (bi)directional LSTM / GRU, forward and backward through time.
//
// Structure per layer and direction:
//   1. input projection for all T steps at once (skinny GEMM, dense.cu): gx = x W_ih^T + b_ih
//   2. T recurrent steps; one launch serves both directions (blockIdx.y).  A step CTA keeps 8 batch rows
//      of h_{t-1} in shared memory; each warp owns one hidden unit and streams the G gate rows of W_hh
//      (coalesced along k), reduces with shuffles and applies the cell non-linearity in registers.
//   3. backward: a point-wise kernel turns (dh, dc) into gate pre-activation gradients, a skinny GEMM
//      carries them through W_hh to dh_{t-1}; weight / bias / input gradients are four GEMMs over all
//      B*T rows after the time loop.
// The weight matrices are the dominant traffic (Odom-LSTM: 143 MB per pass) and are read once per step.
#include <vector>

#include "common.cuh"
#include "dense.cuh"

namespace dlio {

constexpr int RB = 8;  // batch rows per step CTA
constexpr int RJ = 8;  // hidden units per step CTA (one per warp)

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

struct StepDir {
    const float *w_hh, *b_hh, *gx;
    const float *hprev;  // [B rows, stride hprev_ld] or nullptr (zeros)
    const float *cprev;  // LSTM
    long long hprev_ld, cprev_ld;
    float *gates, *cst, *hprev_save, *cprev_save;
    float *h_out;  // + b * h_out_ld + j
    long long h_out_ld;
    float *hn, *cn;  // optional final-state outputs [B, H]
    int t;
};
struct StepArgs {
    StepDir d[2];
    int B, T, H;
};

template <int KIND>  // 0 LSTM, 1 GRU
__global__ void __launch_bounds__(256) rnn_step_fwd_kernel(StepArgs a) {
    extern __shared__ float hs[];  // [RB][H]
    constexpr int G = KIND == 0 ? 4 : 3;
    const StepDir &s = a.d[blockIdx.y];
    const int H = a.H, T = a.T;
    const int b0 = blockIdx.z * RB;
    const int nb = min(RB, a.B - b0);
    for (int i = threadIdx.x; i < RB * H; i += 256) {
        int b = i / H, k = i - b * H;
        hs[i] = (b < nb && s.hprev) ? s.hprev[(size_t)(b0 + b) * s.hprev_ld + k] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * RJ + warp;
    if (j >= H) return;
    float acc[G][RB];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int b = 0; b < RB; ++b) acc[g][b] = 0.f;
    for (int k = lane; k < H; k += 32) {
        float wv[G];
#pragma unroll
        for (int g = 0; g < G; ++g) wv[g] = s.w_hh[((size_t)g * H + j) * H + k];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            float hv = hs[b * H + k];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g][b] = fmaf(wv[g], hv, acc[g][b]);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int b = 0; b < RB; ++b) acc[g][b] = warp_sum(acc[g][b]);
    float pre[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float v = 0.f;
#pragma unroll
        for (int b = 0; b < RB; ++b)
            if (lane == b) v = acc[g][b];
        pre[g] = v;
    }
    if (lane >= nb) return;
    const int b = b0 + lane;
    const size_t row = (size_t)b * T + s.t;
    const float hp = hs[lane * H + j];
    float h;
    if (KIND == 0) {
        float gi = sigmoidf_(s.gx[row * 4 * H + j] + pre[0] + s.b_hh[j]);
        float gf = sigmoidf_(s.gx[row * 4 * H + H + j] + pre[1] + s.b_hh[H + j]);
        float gg = tanhf(s.gx[row * 4 * H + 2 * H + j] + pre[2] + s.b_hh[2 * H + j]);
        float go = sigmoidf_(s.gx[row * 4 * H + 3 * H + j] + pre[3] + s.b_hh[3 * H + j]);
        float cp = s.cprev ? s.cprev[(size_t)b * s.cprev_ld + j] : 0.f;
        float c = gf * cp + gi * gg;
        h = go * tanhf(c);
        s.gates[row * 4 * H + j] = gi;
        s.gates[row * 4 * H + H + j] = gf;
        s.gates[row * 4 * H + 2 * H + j] = gg;
        s.gates[row * 4 * H + 3 * H + j] = go;
        s.cst[row * H + j] = c;
        s.cprev_save[row * H + j] = cp;
        if (s.cn) s.cn[(size_t)b * H + j] = c;
    } else {
        float ghn = pre[2] + s.b_hh[2 * H + j];
        float r = sigmoidf_(s.gx[row * 3 * H + j] + pre[0] + s.b_hh[j]);
        float z = sigmoidf_(s.gx[row * 3 * H + H + j] + pre[1] + s.b_hh[H + j]);
        float n = tanhf(s.gx[row * 3 * H + 2 * H + j] + r * ghn);
        h = (1.f - z) * n + z * hp;
        s.gates[row * 4 * H + j] = r;
        s.gates[row * 4 * H + H + j] = z;
        s.gates[row * 4 * H + 2 * H + j] = n;
        s.gates[row * 4 * H + 3 * H + j] = ghn;
    }
    s.hprev_save[row * H + j] = hp;
    s.h_out[(size_t)b * s.h_out_ld + j] = h;
    if (s.hn) s.hn[(size_t)b * H + j] = h;
}

struct BwdDir {
    const float *gates, *cst, *hprev_save, *cprev_save;
    const float *dout;  // + b * dout_ld + j, or nullptr
    long long dout_ld;
    float *dh_rec, *dc_rec;  // [B, H], read then overwritten with the direct recurrent term
    float *dgx, *dgh;        // [B*T, G*H]
    int t;
};
struct BwdArgs {
    BwdDir d[2];
    int B, T, H;
};

template <int KIND>
__global__ void __launch_bounds__(256) rnn_step_bwd_kernel(BwdArgs a) {
    const BwdDir &s = a.d[blockIdx.y];
    const int H = a.H, T = a.T;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B * H) return;
    int b = i / H, j = i - b * H;
    const size_t row = (size_t)b * T + s.t;
    float dh = s.dh_rec[i] + (s.dout ? s.dout[(size_t)b * s.dout_ld + j] : 0.f);
    if (KIND == 0) {
        float gi = s.gates[row * 4 * H + j], gf = s.gates[row * 4 * H + H + j];
        float gg = s.gates[row * 4 * H + 2 * H + j], go = s.gates[row * 4 * H + 3 * H + j];
        float tc = tanhf(s.cst[row * H + j]);
        float cp = s.cprev_save[row * H + j];
        float dc = dh * go * (1.f - tc * tc) + s.dc_rec[i];
        float d_o = dh * tc;
        s.dgx[row * 4 * H + j] = dc * gg * gi * (1.f - gi);
        s.dgx[row * 4 * H + H + j] = dc * cp * gf * (1.f - gf);
        s.dgx[row * 4 * H + 2 * H + j] = dc * gi * (1.f - gg * gg);
        s.dgx[row * 4 * H + 3 * H + j] = d_o * go * (1.f - go);
        s.dc_rec[i] = dc * gf;
        s.dh_rec[i] = 0.f;
    } else {
        float r = s.gates[row * 4 * H + j], z = s.gates[row * 4 * H + H + j];
        float n = s.gates[row * 4 * H + 2 * H + j], ghn = s.gates[row * 4 * H + 3 * H + j];
        float hp = s.hprev_save[row * H + j];
        float dn_pre = dh * (1.f - z) * (1.f - n * n);
        float dr_pre = dn_pre * ghn * r * (1.f - r);
        float dz_pre = dh * (hp - n) * z * (1.f - z);
        s.dgx[row * 3 * H + j] = dr_pre;
        s.dgx[row * 3 * H + H + j] = dz_pre;
        s.dgx[row * 3 * H + 2 * H + j] = dn_pre;
        s.dgh[row * 3 * H + j] = dr_pre;
        s.dgh[row * 3 * H + H + j] = dz_pre;
        s.dgh[row * 3 * H + 2 * H + j] = dn_pre * r;
        s.dh_rec[i] = dh * z;
    }
}

__global__ void mul_inplace_kernel(float *x, const float *__restrict__ m, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] *= m[i];
}

struct RnnLayout {
    int kind, L, D, B, T, I, H, G;
    size_t bt;
    size_t per_dir;    // gx + gates + cst + hprev_save + cprev_save
    size_t per_layer;  // D * per_dir + layer output
    RnnLayout(int kind_, int L_, int D_, int B_, int T_, int I_, int H_)
        : kind(kind_), L(L_), D(D_), B(B_), T(T_), I(I_), H(H_), G(kind_ == 0 ? 4 : 3) {
        bt = (size_t)B * T;
        per_dir = bt * ((size_t)G * H + 4 * (size_t)H + 3 * (size_t)H);
        per_layer = D * per_dir + bt * D * H;
    }
    size_t total() const { return per_layer * L; }
    float *gx(float *r, int l, int d) const { return r + l * per_layer + d * per_dir; }
    float *gates(float *r, int l, int d) const { return gx(r, l, d) + bt * G * H; }
    float *cst(float *r, int l, int d) const { return gates(r, l, d) + bt * 4 * H; }
    float *hps(float *r, int l, int d) const { return cst(r, l, d) + bt * H; }
    float *cps(float *r, int l, int d) const { return hps(r, l, d) + bt * H; }
    float *yout(float *r, int l) const { return r + l * per_layer + D * per_dir; }
    int in_size(int l) const { return l == 0 ? I : D * H; }
};

static int check_rnn(int kind, int L, int D, int B, int T, int I, int H) {
    DLIO_CHECK_ARG((kind == 0 || kind == 1) && L >= 1 && (D == 1 || D == 2) && B > 0 && T > 0 && I > 0 && H > 0,
                   "rnn: bad dimensions");
    DLIO_CHECK_ARG((size_t)RB * H * sizeof(float) <= 200 * 1024, "rnn: hidden size %d too large", H);
    return DLIO_OK;
}

}  // namespace dlio

using namespace dlio;

extern "C" size_t dlio_rnn_reserve_floats(int kind, int L, int D, int B, int T, int I, int H) {
    return RnnLayout(kind, L, D, B, T, I, H).total();
}
extern "C" size_t dlio_rnn_bwd_scratch_floats(int kind, int L, int D, int B, int T, int I, int H) {
    RnnLayout lay(kind, L, D, B, T, I, H);
    size_t widest = (size_t)(D * H > I ? D * H : I);
    return (size_t)D * 2 * lay.bt * lay.G * H + 2 * lay.bt * widest;
}

extern "C" int dlio_rnn_fwd(int kind, int L, int D, int B, int T, int I, int H, const float *const *weights,
                            const float *x, const float *h0, const float *c0, const float *drop_mask, float *out,
                            float *hn, float *cn, float *reserve, void *stream) {
    ProfScope prof_(DLIO_PROF_RNN, (cudaStream_t)stream);
    int rc = check_rnn(kind, L, D, B, T, I, H);
    if (rc) return rc;
    DLIO_CHECK_ARG(weights && x && out && hn && reserve && (kind == 1 || cn), "rnn_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    RnnLayout lay(kind, L, D, B, T, I, H);
    const int G = lay.G;
    const size_t smem = (size_t)RB * H * sizeof(float);
    if (smem > 48 * 1024) {
        DLIO_CUDA(cudaFuncSetAttribute(rnn_step_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DLIO_CUDA(cudaFuncSetAttribute(rnn_step_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const float *xin = x;
    for (int l = 0; l < L; ++l) {
        const int Il = lay.in_size(l);
        float *yl = (l == L - 1) ? out : lay.yout(reserve, l);
        const long long DH = (long long)D * H;
        for (int d = 0; d < D; ++d) {
            const float *const *w = weights + 4 * (l * D + d);
            if ((rc = linear_fwd_launch(xin, Il, w[0], w[2], nullptr, B * T, G * H, Il, DLIO_ACT_NONE,
                                        lay.gx(reserve, l, d), G * H, st)))
                return rc;
        }
        for (int s = 0; s < T; ++s) {
            StepArgs a;
            a.B = B; a.T = T; a.H = H;
            for (int d = 0; d < D; ++d) {
                const float *const *w = weights + 4 * (l * D + d);
                const int t = d == 0 ? s : T - 1 - s;
                const int tp = d == 0 ? t - 1 : t + 1;
                const size_t slot = (size_t)(l * D + d) * B * H;
                StepDir &q = a.d[d];
                q.w_hh = w[1]; q.b_hh = w[3]; q.gx = lay.gx(reserve, l, d);
                if (s == 0) {
                    q.hprev = h0 ? h0 + slot : nullptr; q.hprev_ld = H;
                    q.cprev = c0 ? c0 + slot : nullptr; q.cprev_ld = H;
                } else {
                    q.hprev = yl + (size_t)tp * DH + (size_t)d * H; q.hprev_ld = (long long)T * DH;
                    q.cprev = lay.cst(reserve, l, d) + (size_t)tp * H; q.cprev_ld = (long long)T * H;
                }
                q.gates = lay.gates(reserve, l, d); q.cst = lay.cst(reserve, l, d);
                q.hprev_save = lay.hps(reserve, l, d); q.cprev_save = lay.cps(reserve, l, d);
                q.h_out = yl + (size_t)t * DH + (size_t)d * H; q.h_out_ld = (long long)T * DH;
                q.hn = (s == T - 1) ? hn + slot : nullptr;
                q.cn = (s == T - 1 && kind == 0) ? cn + slot : nullptr;
                q.t = t;
            }
            dim3 grid(ceil_div(H, RJ), D, ceil_div(B, RB));
            if (kind == 0) rnn_step_fwd_kernel<0><<<grid, 256, smem, st>>>(a);
            else rnn_step_fwd_kernel<1><<<grid, 256, smem, st>>>(a);
            DLIO_LAUNCH_CHECK();
        }
        if (l < L - 1 && drop_mask) {
            long long n = (long long)B * T * DH;
            mul_inplace_kernel<<<ceil_div(n, 256) > 1184 ? 1184 : ceil_div(n, 256), 256, 0, st>>>(
                yl, drop_mask + (size_t)l * n, n);
            DLIO_LAUNCH_CHECK();
        }
        xin = yl;
    }
    return DLIO_OK;
}

extern "C" int dlio_rnn_bwd(int kind, int L, int D, int B, int T, int I, int H, const float *const *weights,
                            const float *x, const float *drop_mask, const float *dout, const float *dhn,
                            const float *dcn, const float *reserve_c, float *const *grads, float *dx, float *dh0,
                            float *dc0, float *scratch, size_t scratch_floats, void *stream) {
    ProfScope prof_(DLIO_PROF_RNN, (cudaStream_t)stream);
    int rc = check_rnn(kind, L, D, B, T, I, H);
    if (rc) return rc;
    DLIO_CHECK_ARG(weights && x && reserve_c && grads && dh0 && scratch && (kind == 1 || dc0), "rnn_bwd: null pointer");
    DLIO_CHECK_ARG(scratch_floats >= dlio_rnn_bwd_scratch_floats(kind, L, D, B, T, I, H), "rnn_bwd: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    RnnLayout lay(kind, L, D, B, T, I, H);
    float *reserve = const_cast<float *>(reserve_c);
    const int G = lay.G;
    const size_t bt = lay.bt;
    const long long DH = (long long)D * H;
    const size_t state = (size_t)L * D * B * H;
    // recurrent gradient accumulators live in dh0 / dc0
    if (dhn) DLIO_CUDA(cudaMemcpyAsync(dh0, dhn, state * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else DLIO_CUDA(cudaMemsetAsync(dh0, 0, state * sizeof(float), st));
    if (kind == 0) {
        if (dcn) DLIO_CUDA(cudaMemcpyAsync(dc0, dcn, state * sizeof(float), cudaMemcpyDeviceToDevice, st));
        else DLIO_CUDA(cudaMemsetAsync(dc0, 0, state * sizeof(float), st));
    }
    float *dg[2][2];  // [dir][0: dgx, 1: dgh]
    float *p = scratch;
    for (int d = 0; d < D; ++d) {
        dg[d][0] = p; p += bt * G * H;
        dg[d][1] = kind == 0 ? dg[d][0] : p; p += bt * G * H;
    }
    const size_t widest = (size_t)(DH > I ? DH : I);
    float *dbuf[2] = {p, p + bt * widest};
    const float *dy = dout;  // gradient w.r.t. the current layer's output [B, T, D*H]
    for (int l = L - 1; l >= 0; --l) {
        const int Il = lay.in_size(l);
        const float *xin = l == 0 ? x : lay.yout(reserve, l - 1);
        for (int s = T - 1; s >= 0; --s) {
            BwdArgs a;
            a.B = B; a.T = T; a.H = H;
            for (int d = 0; d < D; ++d) {
                const int t = d == 0 ? s : T - 1 - s;
                const size_t slot = (size_t)(l * D + d) * B * H;
                BwdDir &q = a.d[d];
                q.gates = lay.gates(reserve, l, d); q.cst = lay.cst(reserve, l, d);
                q.hprev_save = lay.hps(reserve, l, d); q.cprev_save = lay.cps(reserve, l, d);
                q.dout = dy ? dy + (size_t)t * DH + (size_t)d * H : nullptr; q.dout_ld = (long long)T * DH;
                q.dh_rec = dh0 + slot; q.dc_rec = kind == 0 ? dc0 + slot : nullptr;
                q.dgx = dg[d][0]; q.dgh = dg[d][1]; q.t = t;
            }
            dim3 grid(ceil_div((long long)B * H, 256), D);
            if (kind == 0) rnn_step_bwd_kernel<0><<<grid, 256, 0, st>>>(a);
            else rnn_step_bwd_kernel<1><<<grid, 256, 0, st>>>(a);
            DLIO_LAUNCH_CHECK();
            for (int d = 0; d < D; ++d) {
                const float *const *w = weights + 4 * (l * D + d);
                const int t = d == 0 ? s : T - 1 - s;
                // dh_{t-1} += dgh_t W_hh        (rows b of dgh_t are T*G*H apart)
                if ((rc = linear_dx_launch(dg[d][1] + (size_t)t * G * H, T * G * H, w[1], B, G * H, H,
                                           dh0 + (size_t)(l * D + d) * B * H, H, st)))
                    return rc;
            }
        }
        float *dxl = (l == 0) ? dx : dbuf[l & 1];
        if (dxl) DLIO_CUDA(cudaMemsetAsync(dxl, 0, bt * Il * sizeof(float), st));
        for (int d = 0; d < D; ++d) {
            const float *const *w = weights + 4 * (l * D + d);
            float *const *g = grads + 4 * (l * D + d);
            if ((rc = linear_dw_launch(dg[d][0], G * H, xin, Il, (int)bt, G * H, Il, g[0], st))) return rc;
            if ((rc = linear_dw_launch(dg[d][1], G * H, lay.hps(reserve, l, d), H, (int)bt, G * H, H, g[1], st))) return rc;
            if ((rc = colsum_launch(dg[d][0], G * H, (int)bt, G * H, g[2], st))) return rc;
            if ((rc = colsum_launch(dg[d][1], G * H, (int)bt, G * H, g[3], st))) return rc;
            if (dxl && (rc = linear_dx_launch(dg[d][0], G * H, w[0], (int)bt, G * H, Il, dxl, Il, st))) return rc;
        }
        if (l > 0 && drop_mask) {
            long long n = (long long)bt * DH;
            mul_inplace_kernel<<<ceil_div(n, 256) > 1184 ? 1184 : ceil_div(n, 256), 256, 0, st>>>(
                dxl, drop_mask + (size_t)(l - 1) * n, n);
            DLIO_LAUNCH_CHECK();
        }
        dy = dxl;
    }
    return DLIO_OK;
}
