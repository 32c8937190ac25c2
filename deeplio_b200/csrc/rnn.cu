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

// ------------------------------------------------------------------ whole-sequence kernels (short hidden size)
// For the IMU nets (H = 128, T = 15 .. 50) one launch runs ALL T steps of a layer.  A CLUSTER OF FOUR CTAs serves
// one (direction, chunk of RB batch rows): each CTA keeps the W_hh rows of a quarter of the hidden units in shared
// memory for the whole sequence (4 * 32 * 128 floats = 64 KB at H = 128), h_{t-1} in shared memory and c_{t-1} in
// registers of the thread that owns (unit, batch row).  Every CTA writes its part of h_t into the NEXT h buffer of
// all four CTAs (its peers' through distributed shared memory); one cluster barrier per step.  60 step launches per
// IMU window become 4.
constexpr int SEQ_WARPS = 16;
constexpr int SEQ_CLUSTER = 4;     // CTAs per (direction, batch chunk): each owns H / 4 hidden units

__device__ __forceinline__ float warp_colsum32_rnn(float (&v)[32], int lane) {
#pragma unroll
    for (int n = 16; n >= 1; n >>= 1) {
        const bool up = (lane & n) != 0;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const float keep = up ? v[j + n] : v[j];
            const float send = up ? v[j] : v[j + n];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return v[0];
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of both CTAs; orders the global-memory writes before it against the reads after it
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared-memory address of `p` in the peer CTA `rank` of the cluster, and a store through it (DSMEM)
__device__ __forceinline__ uint32_t peer_smem(const void *p, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_peer(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct SeqDir {
    const float *w_hh, *b_hh, *gx;
    const float *h0, *c0;          // [B, H] or nullptr (zeros)
    float *gates, *cst, *hprev_save, *cprev_save;
    float *y;                      // layer output [B, T, DH], this direction's columns start at y
    float *hn, *cn;                // [B, H]
};
struct SeqArgs {
    SeqDir d[2];
    int B, T, H, DH;
};
constexpr int SEQ_ITEMS = 2;       // (batch row, unit) pairs per thread in the point-wise phase: RB * H / 2 <= 1024
constexpr int SEQ_ROWT = 128;      // row threads of the recurrent product (x 4 k groups = 512 threads)
constexpr int SEQ_MAXR = 3;        // W_hh rows per row thread: G * H / 2 <= 384

// Step structure: (A) the recurrent pre-activations as a register-blocked product: thread (k group, row thread)
// owns up to three W_hh rows x all RB batch rows for a quarter of the k range -- per k two 128-bit loads of h
// (batch-contiguous, broadcast) and one conflict-free load per row feed 8 FMAs per row, no shuffles; the four
// partial sums meet in shared memory.  (B) one thread per (batch row, unit) applies the cell non-linearity -- its
// input-projection values gx were prefetched before (A), so their L2 latency is hidden -- keeps c_t in a register,
// writes the saved tensors, and stores h_t into the NEXT h buffer of BOTH CTAs (the peer's through distributed
// shared memory); one cluster barrier per step.  (A first version with one warp per hidden unit and a shuffle
// transpose-reduce per unit spent 77 % of its 2200 instructions per warp and step outside the FMAs.)
// part[(kg * RB + b) * R + r] = sum over k in [k_begin, k_end) of W_hh^T[k][r] * h[k][b] for this thread's NROW rows
template <int NROW>
__device__ __forceinline__ void seq_partial_product(const float *ws, const float *hc, float *part, int k_begin, int k_end,
                                                    int RP, int rp, int kg, int R, const bool (&rowok)[SEQ_MAXR]) {
    float acc[NROW][RB];
#pragma unroll
    for (int m = 0; m < NROW; ++m)
#pragma unroll
        for (int b = 0; b < RB; ++b) acc[m][b] = 0.f;
    const float *wr = ws + (size_t)k_begin * RP + rp;
    const float *hp4 = hc + k_begin * RB;
#pragma unroll 4
    for (int k = k_begin; k < k_end; ++k, wr += RP, hp4 += RB) {
        const float4 h0 = *reinterpret_cast<const float4 *>(hp4);
        const float4 h1 = *reinterpret_cast<const float4 *>(hp4 + 4);
        const float hv[RB] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int m = 0; m < NROW; ++m) {
            const float w = rowok[m] ? wr[m * SEQ_ROWT] : 0.f;  // select, not a branch
#pragma unroll
            for (int b = 0; b < RB; ++b) acc[m][b] = fmaf(w, hv[b], acc[m][b]);
        }
    }
#pragma unroll
    for (int m = 0; m < NROW; ++m) {
        const int r = rp + m * SEQ_ROWT;
        if (r < R) {
#pragma unroll
            for (int b = 0; b < RB; ++b) part[(kg * RB + b) * R + r] = acc[m][b];
        }
    }
}

template <int KIND>
__global__ void __cluster_dims__(SEQ_CLUSTER, 1, 1) __launch_bounds__(SEQ_WARPS * 32) rnn_seq_fwd_kernel(SeqArgs a) {
    extern __shared__ float sm[];
    constexpr int G = KIND == 0 ? 4 : 3;
    const uint32_t rank = cluster_rank();
    const SeqDir &s = a.d[blockIdx.x / SEQ_CLUSTER];
    const bool reverse = blockIdx.x / SEQ_CLUSTER == 1;
    const int H = a.H, T = a.T, HU = H / SEQ_CLUSTER, j0 = (int)rank * HU;
    const int R = G * HU, RP = R + 1;          // own W_hh rows (g, u); padded row count of the transposed copy
    const int b0 = blockIdx.y * RB;
    const int nb = min(RB, a.B - b0);
    float *ws = sm;                            // [H][RP]: ws[k][g * HU + u] = W_hh[g * H + j0 + u][k]
    float *hs = ws + (((size_t)H * RP + 3) & ~(size_t)3);   // [2][H][RB], 16-byte aligned
    float *part = hs + 2 * H * RB;             // [4 k groups][RB][R]
    float *bhh = part + 4 * RB * R;            // [R]
    // W_hh slice -> shared memory, 8 independent loads in flight per thread (one load per iteration made this
    // prologue ~60 us of pure L2 latency per launch)
    for (int i0 = threadIdx.x; i0 < R * H; i0 += 8 * blockDim.x) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * blockDim.x;
            const int k = i % H, r = i / H, g = r / HU, u = r - g * HU;      // coalesced along k
            v[q] = i < R * H ? __ldg(s.w_hh + ((size_t)g * H + j0 + u) * H + k) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * blockDim.x;
            if (i < R * H) ws[(size_t)(i % H) * RP + i / H] = v[q];
        }
    }
    for (int i = threadIdx.x; i < R; i += blockDim.x) bhh[i] = __ldg(s.b_hh + (size_t)(i / HU) * H + j0 + i % HU);
    for (int i = threadIdx.x; i < RB * H; i += blockDim.x) {
        int b = i / H, k = i - b * H;
        hs[k * RB + b] = (b < nb && s.h0) ? s.h0[(size_t)(b0 + b) * H + k] : 0.f;
    }
    // point-wise role: items it = threadIdx.x + q * blockDim.x  <->  (batch row it / HU, unit it % HU)
    float creg[SEQ_ITEMS];
#pragma unroll
    for (int q = 0; q < SEQ_ITEMS; ++q) {
        const int it = threadIdx.x + q * blockDim.x, bl = it / HU, u = it - bl * HU;
        creg[q] = (KIND == 0 && it < RB * HU && bl < nb && s.c0) ? s.c0[(size_t)(b0 + bl) * H + j0 + u] : 0.f;
    }
    uint32_t peer_hs[SEQ_CLUSTER - 1];         // the other CTAs' h buffers
#pragma unroll
    for (int p = 0; p < SEQ_CLUSTER - 1; ++p) peer_hs[p] = peer_smem(hs, (rank + 1u + p) % SEQ_CLUSTER);
    const int nrow = (R + SEQ_ROWT - 1) / SEQ_ROWT;         // W_hh rows per row thread actually in use (uniform)
    const int kg = threadIdx.x / SEQ_ROWT, rp = threadIdx.x % SEQ_ROWT;
    const int kper = (H + 3) / 4, k_begin = kg * kper, k_end = min(H, k_begin + kper);
    bool rowok[SEQ_MAXR];
#pragma unroll
    for (int m = 0; m < SEQ_MAXR; ++m) rowok[m] = rp + m * SEQ_ROWT < R;
    cluster_barrier();                         // both CTAs are set up before any remote store
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = reverse ? T - 1 - step : step;
        const float *hc = hs + cur * H * RB;
        // prefetch the input projections of this step for the point-wise role
        float gxr[SEQ_ITEMS][G];
#pragma unroll
        for (int q = 0; q < SEQ_ITEMS; ++q) {
            const int it = threadIdx.x + q * blockDim.x, bl = it / HU, u = it - bl * HU;
            if (it < RB * HU && bl < nb) {
                const float *gx = s.gx + ((size_t)(b0 + bl) * T + t) * G * H + j0 + u;
#pragma unroll
                for (int g = 0; g < G; ++g) gxr[q][g] = __ldg(gx + (size_t)g * H);
            }
        }
        // (A) recurrent pre-activations, partial over this thread's k range.  The number of W_hh rows per row thread
        // (1 at H = 128) is a compile-time constant of the specialised loop: with a run-time bound the inner loop
        // carried a branch per row and k -- 27 % of the kernel's instructions (ncu source view), and the kernel is
        // issue-bound at four warps per scheduler.
        if (nrow == 1) seq_partial_product<1>(ws, hc, part, k_begin, k_end, RP, rp, kg, R, rowok);
        else if (nrow == 2) seq_partial_product<2>(ws, hc, part, k_begin, k_end, RP, rp, kg, R, rowok);
        else seq_partial_product<3>(ws, hc, part, k_begin, k_end, RP, rp, kg, R, rowok);
        __syncthreads();
        // (B) cell update, one (batch row, unit) per thread.  The new h goes to the NEXT h buffer of all four CTAs
        // first; the tensors saved for the backward pass (gates, c, h_{t-1}, y) are stored to global memory AFTER the
        // barrier arrive: its release (a MEMBAR.ALL.GPU) then waits for the shared-memory exchange only, and the
        // global stores drain while the CTA waits for its peers (they are ordered by the next step's arrive).
        float *hnext = hs + (cur ^ 1) * H * RB;
        float sv[SEQ_ITEMS][8];        // LSTM: gi gf gg go c cp hp h;  GRU: r z n ghn hp h
        bool live[SEQ_ITEMS];
#pragma unroll
        for (int q = 0; q < SEQ_ITEMS; ++q) {
            const int it = threadIdx.x + q * blockDim.x, bl = it / HU, u = it - bl * HU;
            live[q] = it < RB * HU && bl < nb;
            if (!live[q]) continue;
            const int j = j0 + u;
            const float hp = hc[j * RB + bl];
            float pre[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int r = g * HU + u;
                pre[g] = ((part[(0 * RB + bl) * R + r] + part[(1 * RB + bl) * R + r]) +
                          (part[(2 * RB + bl) * R + r] + part[(3 * RB + bl) * R + r])) + bhh[r];
            }
            float h;
            if (KIND == 0) {
                const float gi = sigmoidf_(gxr[q][0] + pre[0]), gf = sigmoidf_(gxr[q][1] + pre[1]);
                const float gg = tanhf(gxr[q][2] + pre[2]), go = sigmoidf_(gxr[q][3] + pre[3]);
                const float cp = creg[q];
                const float c = gf * cp + gi * gg;
                h = go * tanhf(c);
                creg[q] = c;
                sv[q][0] = gi; sv[q][1] = gf; sv[q][2] = gg; sv[q][3] = go; sv[q][4] = c; sv[q][5] = cp;
            } else {
                const float ghn = pre[2];
                const float r = sigmoidf_(gxr[q][0] + pre[0]), z = sigmoidf_(gxr[q][1] + pre[1]);
                const float n = tanhf(gxr[q][2] + r * ghn);
                h = (1.f - z) * n + z * hp;
                sv[q][0] = r; sv[q][1] = z; sv[q][2] = n; sv[q][3] = ghn;
            }
            sv[q][6] = hp; sv[q][7] = h;
            hnext[j * RB + bl] = h;
#pragma unroll
            for (int p = 0; p < SEQ_CLUSTER - 1; ++p)
                st_peer(peer_hs[p] + (uint32_t)(((cur ^ 1) * H * RB + j * RB + bl) * sizeof(float)), h);
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < SEQ_ITEMS; ++q) {
            if (!live[q]) continue;
            const int it = threadIdx.x + q * blockDim.x, bl = it / HU, u = it - bl * HU;
            const int b = b0 + bl, j = j0 + u;
            const size_t row = (size_t)b * T + t;
            float *gt = s.gates + row * 4 * H;
            gt[j] = sv[q][0]; gt[H + j] = sv[q][1]; gt[2 * H + j] = sv[q][2]; gt[3 * H + j] = sv[q][3];
            if (KIND == 0) {
                s.cst[row * H + j] = sv[q][4];
                s.cprev_save[row * H + j] = sv[q][5];
                if (step == T - 1 && s.cn) s.cn[(size_t)b * H + j] = sv[q][4];
            }
            s.hprev_save[row * H + j] = sv[q][6];
            s.y[((size_t)b * T + t) * a.DH + j] = sv[q][7];
            if (step == T - 1 && s.hn) s.hn[(size_t)b * H + j] = sv[q][7];
        }
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        // h_t complete in all CTAs; everybody is done reading h_{t-1} and the partial sums
        cur ^= 1;
    }
}

// Backward through time in one launch (same clusters).  Per step: (1) point-wise gate gradients of the CTA's own
// units from (dh, dc) -- written to global memory for the weight-gradient GEMMs that follow the time loop and kept
// in shared memory (the saved gates of the NEXT step are prefetched meanwhile); (2) the recurrent product over
// the CTA's own W_hh rows, dh_{t-1}[b][k] += sum_r dg[b][r] W_hh[r][k] for ALL k: the half that belongs to the
// peer's units is stored straight into the peer's shared memory (double-buffered), one cluster barrier per step.
struct SeqBwdDir {
    const float *w_hh;
    const float *gates, *cst, *hprev_save, *cprev_save;
    const float *dout;             // gradient of the layer output, this direction's columns ([B, T, DH]) or nullptr
    float *dh_rec, *dc_rec;        // [B, H]: in: gradient of (h_n, c_n); out: gradient of (h_0, c_0)
    float *dgx, *dgh;              // [B*T, G*H]
};
struct SeqBwdArgs {
    SeqBwdDir d[2];
    int B, T, H, DH;
};

template <int KIND>
__global__ void __cluster_dims__(SEQ_CLUSTER, 1, 1) __launch_bounds__(SEQ_WARPS * 32) rnn_seq_bwd_kernel(SeqBwdArgs a) {
    extern __shared__ float sm[];
    constexpr int G = KIND == 0 ? 4 : 3;
    const uint32_t rank = cluster_rank();
    const SeqBwdDir &s = a.d[blockIdx.x / SEQ_CLUSTER];
    const bool reverse = blockIdx.x / SEQ_CLUSTER == 1;
    const int H = a.H, T = a.T, HU = H / SEQ_CLUSTER, j0 = (int)rank * HU, GU = G * HU;
    const int b0 = blockIdx.y * RB;
    const int nb = min(RB, a.B - b0);
    const int ngroups = blockDim.x / H;
    float *ws = sm;                          // [G * HU][H]
    float *dg = ws + (size_t)GU * H;         // [G * HU][RB] recurrent gate gradients of this step (own units), batch-contiguous
    float *dh = dg + RB * GU;                // [RB][HU]
    float *dc = dh + RB * HU;                // [RB][HU]
    float *part = dc + RB * HU;              // [ngroups][RB][H] partial products of the row groups
    float *xin = part + ngroups * RB * H;    // [2][SEQ_CLUSTER - 1][RB][HU] contributions received from the peers (by step parity)
    for (int i0 = threadIdx.x; i0 < GU * H; i0 += 8 * blockDim.x) {      // 8 independent loads in flight
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * blockDim.x;
            const int k = i % H, ru = i / H, g = ru / HU, u = ru - g * HU;
            v[q] = i < GU * H ? __ldg(s.w_hh + ((size_t)g * H + j0 + u) * H + k) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * blockDim.x;
            if (i < GU * H) ws[i] = v[q];
        }
    }
    for (int i = threadIdx.x; i < RB * HU; i += blockDim.x) {
        int b = i / HU, u = i - b * HU;
        dh[i] = b < nb ? s.dh_rec[(size_t)(b0 + b) * H + j0 + u] : 0.f;
        dc[i] = (KIND == 0 && b < nb) ? s.dc_rec[(size_t)(b0 + b) * H + j0 + u] : 0.f;
    }
    // CTA `owner` receives the contribution of sender `rank` in slot (rank - owner - 1) mod CLUSTER of its xin
    uint32_t peer_xin[SEQ_CLUSTER];
#pragma unroll
    for (int o = 0; o < SEQ_CLUSTER; ++o)
        peer_xin[o] = peer_smem(xin, (uint32_t)o) +
                      (uint32_t)((((int)rank - o - 1 + SEQ_CLUSTER) % SEQ_CLUSTER) * RB * HU * sizeof(float));
    cluster_barrier();
    const int kq = threadIdx.x % H, rg = threadIdx.x / H;
    // saved tensors of one (batch row, unit) at one time step: LSTM gi gf gg go c cp dout; GRU r z n ghn hp dout.
    // A thread's (first) item of the NEXT step is loaded while the recurrent product of this step runs, so the L2
    // round trip of seven dependent-free loads is off the per-step critical path.
    auto load_saved = [&](int i, int tt, float (&v)[7]) {
        const int bl = i / HU, u = i - bl * HU, j = j0 + u;
        if (bl >= nb) return;
        const int b = b0 + bl;
        const size_t row = (size_t)b * T + tt;
        const float *gt = s.gates + row * 4 * H;
        v[0] = __ldg(gt + j); v[1] = __ldg(gt + H + j); v[2] = __ldg(gt + 2 * H + j); v[3] = __ldg(gt + 3 * H + j);
        if (KIND == 0) {
            v[4] = __ldg(s.cst + row * H + j);
            v[5] = __ldg(s.cprev_save + row * H + j);
        } else {
            v[4] = __ldg(s.hprev_save + row * H + j);
            v[5] = 0.f;
        }
        v[6] = s.dout ? __ldg(s.dout + ((size_t)b * T + tt) * a.DH + j) : 0.f;
    };
    float pf[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if ((int)threadIdx.x < RB * HU) load_saved(threadIdx.x, reverse ? 0 : T - 1, pf);
    for (int step = T - 1; step >= 0; --step) {
        const int t = reverse ? T - 1 - step : step;
        for (int i = threadIdx.x; i < RB * HU; i += blockDim.x) {
            const int bl = i / HU, u = i - bl * HU, j = j0 + u;
            float *dgr = dg + u * RB + bl;       // + g * HU * RB
            if (bl >= nb) {
#pragma unroll
                for (int g = 0; g < G; ++g) dgr[g * HU * RB] = 0.f;
                continue;
            }
            const int b = b0 + bl;
            const size_t row = (size_t)b * T + t;
            float sv[7];
            if (i == (int)threadIdx.x) {
#pragma unroll
                for (int q = 0; q < 7; ++q) sv[q] = pf[q];
            } else {
                load_saved(i, t, sv);
            }
            const float dhv = dh[i] + sv[6];
            if (KIND == 0) {
                const float gi = sv[0], gf = sv[1], gg = sv[2], go = sv[3];
                const float tc = tanhf(sv[4]);
                const float cp = sv[5];
                const float dcv = dhv * go * (1.f - tc * tc) + dc[i];
                const float d0 = dcv * gg * gi * (1.f - gi), d1 = dcv * cp * gf * (1.f - gf);
                const float d2 = dcv * gi * (1.f - gg * gg), d3 = dhv * tc * go * (1.f - go);
                float *o = s.dgx + row * 4 * H;
                o[j] = d0; o[H + j] = d1; o[2 * H + j] = d2; o[3 * H + j] = d3;
                dgr[0] = d0; dgr[HU * RB] = d1; dgr[2 * HU * RB] = d2; dgr[3 * HU * RB] = d3;
                dc[i] = dcv * gf;
                dh[i] = 0.f;
            } else {
                const float r = sv[0], z = sv[1], n = sv[2], ghn = sv[3];
                const float hp = sv[4];
                const float dn_pre = dhv * (1.f - z) * (1.f - n * n);
                const float dr_pre = dn_pre * ghn * r * (1.f - r);
                const float dz_pre = dhv * (hp - n) * z * (1.f - z);
                float *ox = s.dgx + row * 3 * H, *oh = s.dgh + row * 3 * H;
                ox[j] = dr_pre; ox[H + j] = dz_pre; ox[2 * H + j] = dn_pre;
                oh[j] = dr_pre; oh[H + j] = dz_pre; oh[2 * H + j] = dn_pre * r;
                dgr[0] = dr_pre; dgr[HU * RB] = dz_pre; dgr[2 * HU * RB] = dn_pre * r;
                dh[i] = dhv * z;
            }
        }
        if (step > 0 && (int)threadIdx.x < RB * HU) load_saved(threadIdx.x, reverse ? T - step : step - 1, pf);
        __syncthreads();
        // partial dh_{t-1}[b][k] over this CTA's rows; row group rg handles rows rg, rg + ngroups, ...
        if (rg < ngroups) {
            float acc[RB];
#pragma unroll
            for (int b = 0; b < RB; ++b) acc[b] = 0.f;
#pragma unroll 4
            for (int r = rg; r < GU; r += ngroups) {
                const float wv = ws[(size_t)r * H + kq];
                const float4 g0 = *reinterpret_cast<const float4 *>(dg + r * RB);       // broadcast loads
                const float4 g1 = *reinterpret_cast<const float4 *>(dg + r * RB + 4);
                acc[0] = fmaf(g0.x, wv, acc[0]); acc[1] = fmaf(g0.y, wv, acc[1]);
                acc[2] = fmaf(g0.z, wv, acc[2]); acc[3] = fmaf(g0.w, wv, acc[3]);
                acc[4] = fmaf(g1.x, wv, acc[4]); acc[5] = fmaf(g1.y, wv, acc[5]);
                acc[6] = fmaf(g1.z, wv, acc[6]); acc[7] = fmaf(g1.w, wv, acc[7]);
            }
#pragma unroll
            for (int b = 0; b < RB; ++b) part[(rg * RB + b) * H + kq] = acc[b];
        }
        __syncthreads();
        const int par = step & 1;
        for (int i = threadIdx.x; i < RB * H; i += blockDim.x) {
            const int bl = i / H, k = i - bl * H;
            float v = 0.f;
            for (int q = 0; q < ngroups; ++q) v += part[(q * RB + bl) * H + k];
            const int owner = k / HU, u = k - owner * HU;
            if (owner == (int)rank) dh[bl * HU + u] += v;
            else st_peer(peer_xin[owner] + (uint32_t)((par * (SEQ_CLUSTER - 1) * RB * HU + bl * HU + u) * sizeof(float)), v);
        }
        cluster_barrier();
        for (int i = threadIdx.x; i < RB * HU; i += blockDim.x) {
            float v = 0.f;
#pragma unroll
            for (int p = 0; p < SEQ_CLUSTER - 1; ++p) v += xin[(par * (SEQ_CLUSTER - 1) + p) * RB * HU + i];
            dh[i] += v;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < RB * HU; i += blockDim.x) {
        int b = i / HU, u = i - b * HU;
        if (b < nb) {
            s.dh_rec[(size_t)(b0 + b) * H + j0 + u] = dh[i];
            if (KIND == 0) s.dc_rec[(size_t)(b0 + b) * H + j0 + u] = dc[i];
        }
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

// whole-sequence kernels: the W_hh rows of half the hidden units must fit in one CTA's shared memory (H <= 152 for
// an LSTM, 176 for a GRU), and a single step is not worth a cluster launch
static size_t seq_fwd_smem(int kind, int H) {
    const size_t G = kind == 0 ? 4 : 3, R = G * (H / SEQ_CLUSTER);
    return ((((size_t)H * (R + 1) + 3) & ~(size_t)3) + 2 * (size_t)RB * H + 4 * RB * R + R) * sizeof(float);
}
static size_t seq_bwd_smem(int kind, int H) {
    const size_t G = kind == 0 ? 4 : 3, HU = H / SEQ_CLUSTER, threads = SEQ_WARPS * 32;
    return (G * HU * H + RB * G * HU + 2 * RB * HU + (threads / H) * RB * H + 2 * (SEQ_CLUSTER - 1) * RB * HU) * sizeof(float);
}
static bool use_seq_kernels(int kind, int H, int T) {
    return H % SEQ_CLUSTER == 0 && H >= 16 && H <= SEQ_WARPS * 32 &&
           (size_t)RB * (H / SEQ_CLUSTER) <= (size_t)SEQ_ITEMS * SEQ_WARPS * 32 && T >= 2 &&
           (kind == 0 ? 4 : 3) * (H / SEQ_CLUSTER) <= SEQ_ROWT * SEQ_MAXR &&
           seq_bwd_smem(kind, H) <= 200 * 1024 && seq_fwd_smem(kind, H) <= 200 * 1024;
}

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
        if (use_seq_kernels(kind, H, T)) {
            // one launch for all T steps (short hidden size: the IMU nets)
            SeqArgs a;
            a.B = B; a.T = T; a.H = H; a.DH = (int)DH;
            for (int d = 0; d < D; ++d) {
                const float *const *w = weights + 4 * (l * D + d);
                const size_t slot = (size_t)(l * D + d) * B * H;
                SeqDir &q = a.d[d];
                q.w_hh = w[1]; q.b_hh = w[3]; q.gx = lay.gx(reserve, l, d);
                q.h0 = h0 ? h0 + slot : nullptr;
                q.c0 = c0 ? c0 + slot : nullptr;
                q.gates = lay.gates(reserve, l, d); q.cst = lay.cst(reserve, l, d);
                q.hprev_save = lay.hps(reserve, l, d); q.cprev_save = lay.cps(reserve, l, d);
                q.y = yl + (size_t)d * H;
                q.hn = hn + slot;
                q.cn = kind == 0 ? cn + slot : nullptr;
            }
            dim3 grid(D * SEQ_CLUSTER, ceil_div(B, RB));
            const size_t sm = seq_fwd_smem(kind, H);
            DLIO_CUDA(cudaFuncSetAttribute(rnn_seq_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            DLIO_CUDA(cudaFuncSetAttribute(rnn_seq_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            if (kind == 0) rnn_seq_fwd_kernel<0><<<grid, SEQ_WARPS * 32, sm, st>>>(a);
            else rnn_seq_fwd_kernel<1><<<grid, SEQ_WARPS * 32, sm, st>>>(a);
            DLIO_LAUNCH_CHECK();
        } else
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
        if (use_seq_kernels(kind, H, T)) {
            SeqBwdArgs a;
            a.B = B; a.T = T; a.H = H; a.DH = (int)DH;
            for (int d = 0; d < D; ++d) {
                const float *const *w = weights + 4 * (l * D + d);
                const size_t slot = (size_t)(l * D + d) * B * H;
                SeqBwdDir &q = a.d[d];
                q.w_hh = w[1];
                q.gates = lay.gates(reserve, l, d); q.cst = lay.cst(reserve, l, d);
                q.hprev_save = lay.hps(reserve, l, d); q.cprev_save = lay.cps(reserve, l, d);
                q.dout = dy ? dy + (size_t)d * H : nullptr;
                q.dh_rec = dh0 + slot; q.dc_rec = kind == 0 ? dc0 + slot : nullptr;
                q.dgx = dg[d][0]; q.dgh = dg[d][1];
            }
            const int threads = SEQ_WARPS * 32;
            const size_t sm = seq_bwd_smem(kind, H);
            DLIO_CUDA(cudaFuncSetAttribute(rnn_seq_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            DLIO_CUDA(cudaFuncSetAttribute(rnn_seq_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            dim3 grid(D * SEQ_CLUSTER, ceil_div(B, RB));
            if (kind == 0) rnn_seq_bwd_kernel<0><<<grid, threads, sm, st>>>(a);
            else rnn_seq_bwd_kernel<1><<<grid, threads, sm, st>>>(a);
            DLIO_LAUNCH_CHECK();
        } else
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
            // weight gradients; the bias gradients (column sums of the same gate gradients) ride along
            if ((rc = linear_dw_launch(dg[d][0], G * H, xin, Il, (int)bt, G * H, Il, g[0], st, g[2]))) return rc;
            if ((rc = linear_dw_launch(dg[d][1], G * H, lay.hps(reserve, l, d), H, (int)bt, G * H, H, g[1], st, g[3]))) return rc;
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
