// Dense layers with a small row count (M = B*S or B*T rows, N up to 4096 features, K up to 2048):
// weight-bandwidth-bound "skinny" GEMMs.  Every kernel streams the weight matrix exactly once per
// 16-row slab of activations, with coalesced accesses along the contiguous dimension of W [N, K].
//
//   forward  y[m, n]  = act(sum_k x[m, k] w[n, k] + b[n])      warp per output feature, lanes along k
//   dgrad    dx[m, k] += sum_n dz[m, n] w[n, k]                thread per k, n split across CTAs (atomics)
//   wgrad    dw[n, k] = sum_m dz[m, n] x[m, k]                 thread per (n, k), loop over m
#include "common.cuh"
#include "dense.cuh"

namespace dlio {

constexpr int LIN_MB = 16;   // activation rows per CTA
constexpr int LIN_NB = 32;   // output features per CTA (4 per warp)
constexpr int LIN_KC = 128;  // k chunk staged in shared memory

__global__ void __launch_bounds__(256) linear_fwd_kernel(const float *__restrict__ x, int ldx,
                                                         const float *__restrict__ w, const float *__restrict__ b,
                                                         const float *__restrict__ b2, int M, int N, int K, int act,
                                                         float *__restrict__ y, int ldy) {
    __shared__ float xs[LIN_MB][LIN_KC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * LIN_NB + warp * 4;
    const int m0 = blockIdx.y * LIN_MB;
    float acc[4][LIN_MB];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int m = 0; m < LIN_MB; ++m) acc[f][m] = 0.f;
    for (int k0 = 0; k0 < K; k0 += LIN_KC) {
        __syncthreads();
        for (int i = threadIdx.x; i < LIN_MB * LIN_KC; i += 256) {
            int m = i / LIN_KC, k = i - m * LIN_KC;
            xs[m][k] = (m0 + m < M && k0 + k < K) ? x[(size_t)(m0 + m) * ldx + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < LIN_KC / 32; ++j) {
            const int kk = j * 32 + lane, k = k0 + kk;
            float wv[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) wv[f] = (n0 + f < N && k < K) ? w[(size_t)(n0 + f) * K + k] : 0.f;
#pragma unroll
            for (int m = 0; m < LIN_MB; ++m) {
                float xv = xs[m][kk];
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[f][m] = fmaf(wv[f], xv, acc[f][m]);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int m = 0; m < LIN_MB; ++m) acc[f][m] = warp_sum(acc[f][m]);
    // lane l < 16 writes row m0 + l for the warp's 4 features
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        float v = 0.f;
#pragma unroll
        for (int m = 0; m < LIN_MB; ++m)
            if (lane == m) v = acc[f][m];
        int n = n0 + f, m = m0 + lane;
        if (lane < LIN_MB && m < M && n < N) {
            if (b) v += b[n];
            if (b2) v += b2[n];
            y[(size_t)m * ldy + n] = act_apply(v, act);
        }
    }
}

// dz = dy * act'(y)
__global__ void act_bwd_kernel(const float *__restrict__ y, int ldy, const float *__restrict__ dy, int lddy, int M,
                               int N, int act, float *__restrict__ dz) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    int m = (int)(i / N), n = (int)(i - (long long)m * N);
    dz[i] = dy[(size_t)m * lddy + n] * act_grad_from_out(y[(size_t)m * ldy + n], act);
}

// out[n] = sum_m dz[m, n]   (8 independent loads in flight per thread: the loop is latency-bound otherwise)
__global__ void colsum_kernel(const float *__restrict__ dz, int lddz, int M, int N, float *__restrict__ out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int m = 0;
    for (; m + 8 <= M; m += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += dz[(size_t)(m + u) * lddz + n];
    }
    for (; m < M; ++m) s[0] += dz[(size_t)m * lddz + n];
    out[n] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

constexpr int DX_NC = 128;  // n rows per CTA (at most; `nc` rows when the problem is too small to fill the GPU)

// The n loop keeps 8 independent weight loads in flight per thread: with one load per iteration a 4-CTA launch
// (the IMU LSTM's recurrent dgrad, 120 launches per step) took 38 us of pure load latency.
__global__ void __launch_bounds__(128) linear_dx_kernel(const float *__restrict__ dz, int lddz,
                                                        const float *__restrict__ w, int M, int N, int K, int nc,
                                                        float *__restrict__ dx, int lddx) {
    __shared__ float zs[LIN_MB][DX_NC];
    const int k = blockIdx.x * 128 + threadIdx.x;
    const int nb = blockIdx.y * nc;
    const int m0 = blockIdx.z * LIN_MB;
    for (int i = threadIdx.x; i < LIN_MB * nc; i += 128) {
        int m = i / nc, n = i - m * nc;
        zs[m][n] = (m0 + m < M && nb + n < N) ? dz[(size_t)(m0 + m) * lddz + nb + n] : 0.f;
    }
    __syncthreads();
    float acc[LIN_MB];
#pragma unroll
    for (int m = 0; m < LIN_MB; ++m) acc[m] = 0.f;
    if (k < K) {
        const int nend = min(nc, N - nb);
        const float *wp = w + (size_t)nb * K + k;
        int n = 0;
        for (; n + 8 <= nend; n += 8) {
            float wv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) wv[u] = wp[(size_t)(n + u) * K];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int m = 0; m < LIN_MB; ++m) acc[m] = fmaf(zs[m][n + u], wv[u], acc[m]);
        }
        for (; n < nend; ++n) {
            float wv = wp[(size_t)n * K];
#pragma unroll
            for (int m = 0; m < LIN_MB; ++m) acc[m] = fmaf(zs[m][n], wv, acc[m]);
        }
#pragma unroll
        for (int m = 0; m < LIN_MB; ++m)
            if (m0 + m < M) atomicAdd(dx + (size_t)(m0 + m) * lddx + k, acc[m]);
    }
}

constexpr int DW_NR = 8;  // n rows per CTA

// db (optional): the bias gradient, out[n] = sum_m dz[m, n] -- taken by the CTAs of the first k block from the dz
// tiles they stage anyway (one launch less per linear layer than a separate column-sum kernel)
__global__ void __launch_bounds__(128) linear_dw_kernel(const float *__restrict__ dz, int lddz,
                                                        const float *__restrict__ x, int ldx, int M, int N, int K,
                                                        float *__restrict__ dw, float *__restrict__ db) {
    __shared__ float zs[64][DW_NR];
    const int k = blockIdx.x * 128 + threadIdx.x;
    const int n0 = blockIdx.y * DW_NR;
    float acc[DW_NR];
#pragma unroll
    for (int r = 0; r < DW_NR; ++r) acc[r] = 0.f;
    const bool sums = db != nullptr && blockIdx.x == 0 && threadIdx.x < DW_NR;
    float bsum = 0.f;
    for (int mb = 0; mb < M; mb += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * DW_NR; i += 128) {
            int m = i / DW_NR, r = i - m * DW_NR;
            zs[m][r] = (mb + m < M && n0 + r < N) ? dz[(size_t)(mb + m) * lddz + n0 + r] : 0.f;
        }
        __syncthreads();
        if (sums) {
            float s8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // the same association as colsum_kernel
#pragma unroll
            for (int m = 0; m < 64; m += 8)
#pragma unroll
                for (int u = 0; u < 8; ++u) s8[u] += zs[m + u][threadIdx.x];
            bsum += ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
        }
        if (k < K) {
            const int mend = min(64, M - mb);
            const float *xp = x + (size_t)mb * ldx + k;
            int m = 0;
            for (; m + 8 <= mend; m += 8) {       // 8 independent loads in flight (latency-bound otherwise)
                float xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) xv[u] = xp[(size_t)(m + u) * ldx];
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int r = 0; r < DW_NR; ++r) acc[r] = fmaf(zs[m + u][r], xv[u], acc[r]);
            }
            for (; m < mend; ++m) {
                float xv = xp[(size_t)m * ldx];
#pragma unroll
                for (int r = 0; r < DW_NR; ++r) acc[r] = fmaf(zs[m][r], xv, acc[r]);
            }
        }
    }
    if (k < K) {
#pragma unroll
        for (int r = 0; r < DW_NR; ++r)
            if (n0 + r < N) dw[(size_t)(n0 + r) * K + k] = acc[r];
    }
    if (sums && n0 + (int)threadIdx.x < N) db[n0 + threadIdx.x] = bsum;
}

int linear_fwd_launch(const float *x, int ldx, const float *w, const float *b, const float *b2, int M, int N, int K,
                      int act, float *y, int ldy, cudaStream_t st) {
    dim3 grid(ceil_div(N, LIN_NB), ceil_div(M, LIN_MB));
    linear_fwd_kernel<<<grid, 256, 0, st>>>(x, ldx, w, b, b2, M, N, K, act, y, ldy);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
int linear_dx_launch(const float *dz, int lddz, const float *w, int M, int N, int K, float *dx, int lddx,
                     cudaStream_t st) {
    // fewer n rows per CTA when 128-row slabs would leave most SMs idle
    int nc = DX_NC;
    while (nc > 16 && (long long)ceil_div(K, 128) * ceil_div(N, nc) * ceil_div(M, LIN_MB) < 148) nc >>= 1;
    dim3 grid(ceil_div(K, 128), ceil_div(N, nc), ceil_div(M, LIN_MB));
    linear_dx_kernel<<<grid, 128, 0, st>>>(dz, lddz, w, M, N, K, nc, dx, lddx);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
int linear_dw_launch(const float *dz, int lddz, const float *x, int ldx, int M, int N, int K, float *dw,
                     cudaStream_t st, float *db) {
    dim3 grid(ceil_div(K, 128), ceil_div(N, DW_NR));
    linear_dw_kernel<<<grid, 128, 0, st>>>(dz, lddz, x, ldx, M, N, K, dw, db);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
int colsum_launch(const float *dz, int lddz, int M, int N, float *out, cudaStream_t st) {
    colsum_kernel<<<ceil_div(N, 128), 128, 0, st>>>(dz, lddz, M, N, out);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_linear_fwd(const float *x, int ldx, const float *w, const float *b, int m, int n, int k, int act,
                               float *y, int ldy, void *stream) {
    ProfScope prof_(DLIO_PROF_DENSE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(x && w && y && m > 0 && n > 0 && k > 0 && ldx >= k && ldy >= n, "linear_fwd: bad argument");
    return linear_fwd_launch(x, ldx, w, b, nullptr, m, n, k, act, y, ldy, (cudaStream_t)stream);
}

extern "C" int dlio_linear_bwd(const float *x, int ldx, const float *w, const float *y, int ldy, const float *dy,
                               int lddy, int m, int n, int k, int act, float *dx, int lddx, float *dw, float *db,
                               float *scratch, void *stream) {
    ProfScope prof_(DLIO_PROF_DENSE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(x && w && dy && m > 0 && n > 0 && k > 0, "linear_bwd: bad argument");
    DLIO_CHECK_ARG(act == DLIO_ACT_NONE || (y && scratch), "linear_bwd: activation backward needs y and scratch");
    cudaStream_t st = (cudaStream_t)stream;
    const float *dz = dy;
    int lddz = lddy;
    if (act != DLIO_ACT_NONE) {
        act_bwd_kernel<<<ceil_div((long long)m * n, 256), 256, 0, st>>>(y, ldy, dy, lddy, m, n, act, scratch);
        DLIO_LAUNCH_CHECK();
        dz = scratch;
        lddz = n;
    }
    int rc;
    if (db && !dw && (rc = colsum_launch(dz, lddz, m, n, db, st))) return rc;
    if (dw && (rc = linear_dw_launch(dz, lddz, x, ldx, m, n, k, dw, st, db))) return rc;
    if (dx) {
        // dx rows may be strided (lddx > k): zero row by row through a 2-D memset
        DLIO_CUDA(cudaMemset2DAsync(dx, (size_t)lddx * sizeof(float), 0, (size_t)k * sizeof(float), m, st));
        if ((rc = linear_dx_launch(dz, lddz, w, m, n, k, dx, lddx, st))) return rc;
    }
    return DLIO_OK;
}
