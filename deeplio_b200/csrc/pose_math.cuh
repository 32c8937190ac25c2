// SO(3) / SE(3) per-sample math shared by the pose kernels (pose.cu) and the host-side self-check
// (tests/native/pose_host.cu compiles these same functions for the CPU, so the arithmetic the GPU runs is checked
// against the oracle in the CPU test tier as well).
#pragma once
#include <math.h>

#include "common.cuh"

namespace dlio {

constexpr float SO3_TOL = 1e-6f;       // liegroups.torch.utils.isclose
constexpr int CHAIN_MAX_S = 64;

__host__ __device__ __forceinline__ void m3_mul(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A^T B
__host__ __device__ __forceinline__ void m3_tmul(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// C = A B^T
__host__ __device__ __forceinline__ void m3_mult(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
__host__ __device__ __forceinline__ float m3_det(const float *A) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
__host__ __device__ __forceinline__ bool finite3(const float *v) { return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]); }

// R = exp(phi)
__host__ __device__ inline void so3_exp(const float *phi, float *R) {
    const float th = sqrtf(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
    if (fabsf(th) < SO3_TOL) {      // I + wedge(phi)
        R[0] = 1.f; R[1] = -phi[2]; R[2] = phi[1];
        R[3] = phi[2]; R[4] = 1.f; R[5] = -phi[0];
        R[6] = -phi[1]; R[7] = phi[0]; R[8] = 1.f;
        return;
    }
    const float a0 = phi[0] / th, a1 = phi[1] / th, a2 = phi[2] / th;
    float s, c;
    sincosf(th, &s, &c);
    const float k = 1.f - c;
    R[0] = c + k * a0 * a0; R[1] = k * a0 * a1 - s * a2; R[2] = k * a0 * a2 + s * a1;
    R[3] = k * a1 * a0 + s * a2; R[4] = c + k * a1 * a1; R[5] = k * a1 * a2 - s * a0;
    R[6] = k * a2 * a0 - s * a1; R[7] = k * a2 * a1 + s * a0; R[8] = c + k * a2 * a2;
}
// dphi = (d exp / d phi)^T G
__host__ __device__ inline void so3_exp_bwd(const float *phi, const float *G, float *dphi) {
    const float v0 = G[7] - G[5], v1 = G[2] - G[6], v2 = G[3] - G[1];       // sum_ij G_ij d wedge(a)_ij / d a_k
    const float th = sqrtf(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
    if (fabsf(th) < SO3_TOL) {
        dphi[0] = v0; dphi[1] = v1; dphi[2] = v2;
        return;
    }
    const float a[3] = {phi[0] / th, phi[1] / th, phi[2] / th};
    float s, c;
    sincosf(th, &s, &c);
    const float tr = G[0] + G[4] + G[8];
    float aGa = 0.f, Ga[3];
    for (int i = 0; i < 3; ++i) {
        float r = 0.f, cc = 0.f;
        for (int j = 0; j < 3; ++j) {
            r += G[3 * i + j] * a[j];
            cc += G[3 * j + i] * a[j];
        }
        Ga[i] = r + cc;                    // ((G + G^T) a)_i
        aGa += a[i] * r;
    }
    const float dth = -s * tr + s * aGa + c * (a[0] * v0 + a[1] * v1 + a[2] * v2);
    const float da[3] = {(1.f - c) * Ga[0] + s * v0, (1.f - c) * Ga[1] + s * v1, (1.f - c) * Ga[2] + s * v2};
    const float ada = a[0] * da[0] + a[1] * da[1] + a[2] * da[2];
    for (int m = 0; m < 3; ++m) dphi[m] = dth * a[m] + (da[m] - a[m] * ada) / th;
}
// liegroups is_valid_matrix: |det - 1| < 1e-6 and every entry of R^T R within 1e-6 of the identity
__host__ __device__ inline bool so3_valid(const float *R) {
    if (!(fabsf(m3_det(R) - 1.f) < SO3_TOL)) return false;
    float P[9];
    m3_tmul(R, R, P);
    for (int i = 0; i < 9; ++i)
        if (!(fabsf(P[i] - ((i % 4 == 0) ? 1.f : 0.f)) < SO3_TOL)) return false;
    return true;
}
// nearest rotation (U V^T of the SVD): Newton iteration X <- (X + X^-T) / 2 on the polar decomposition
__host__ __device__ inline void so3_project(float *R) {
    for (int it = 0; it < 6; ++it) {
        const float d = m3_det(R);
        if (!(fabsf(d) > 1e-20f)) return;
        const float inv = 1.f / d;
        float C[9];                       // cofactor matrix = det * R^-T
        C[0] = R[4] * R[8] - R[5] * R[7]; C[1] = R[5] * R[6] - R[3] * R[8]; C[2] = R[3] * R[7] - R[4] * R[6];
        C[3] = R[2] * R[7] - R[1] * R[8]; C[4] = R[0] * R[8] - R[2] * R[6]; C[5] = R[1] * R[6] - R[0] * R[7];
        C[6] = R[1] * R[5] - R[2] * R[4]; C[7] = R[2] * R[3] - R[0] * R[5]; C[8] = R[0] * R[4] - R[1] * R[3];
        for (int i = 0; i < 9; ++i) R[i] = 0.5f * (R[i] + C[i] * inv);
    }
}
// quaternion (w, x, y, z) of a rotation matrix; returns the branch: 0 regular (qw not close to 0), 1..3 the
// near-zero branches keyed by the largest diagonal entry
__host__ __device__ inline int so3_to_quat(const float *R, float *q) {
    const float qw = 0.5f * sqrtf(fmaxf(1.f + R[0] + R[4] + R[8], 0.f));
    if (!(fabsf(qw) < SO3_TOL)) {
        const float d = 4.f * qw;
        q[0] = qw; q[1] = (R[7] - R[5]) / d; q[2] = (R[2] - R[6]) / d; q[3] = (R[3] - R[1]) / d;
        return 0;
    }
    const int i = (R[0] > R[4] && R[0] > R[8]) ? 0 : (R[4] > R[8] ? 1 : 2);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    const float d = 2.f * sqrtf(fmaxf(1.f + R[4 * i] - R[4 * j] - R[4 * k], 1e-30f));
    q[0] = (R[3 * k + j] - R[3 * j + k]) / d;
    q[1 + i] = 0.25f * d;
    q[1 + j] = (R[3 * j + i] + R[3 * i + j]) / d;
    q[1 + k] = (R[3 * i + k] + R[3 * k + i]) / d;
    return 1 + i;
}
// G += (d quat / d R)^T g, same branch as the forward conversion
__host__ __device__ inline void so3_to_quat_bwd(const float *R, int branch, const float *q, const float *g, float *G) {
    if (branch == 0) {
        const float qw = q[0], d = 4.f * qw;
        const float dd = -(g[1] * q[1] + g[2] * q[2] + g[3] * q[3]) / d;
        const float du = g[0] / (8.f * qw) + dd / (2.f * qw);
        G[0] += du; G[4] += du; G[8] += du;
        G[7] += g[1] / d; G[5] -= g[1] / d;
        G[2] += g[2] / d; G[6] -= g[2] / d;
        G[3] += g[3] / d; G[1] -= g[3] / d;
        return;
    }
    const int i = branch - 1, j = (i + 1) % 3, k = (i + 2) % 3;
    const float d = 4.f * q[1 + i];
    const float dd = 0.25f * g[1 + i] - (g[0] * q[0] + g[1 + j] * q[1 + j] + g[1 + k] * q[1 + k]) / d;
    const float du = dd * 2.f / d;
    G[4 * i] += du; G[4 * j] -= du; G[4 * k] -= du;
    G[3 * k + j] += g[0] / d; G[3 * j + k] -= g[0] / d;
    G[3 * j + i] += g[1 + j] / d; G[3 * i + j] += g[1 + j] / d;
    G[3 * i + k] += g[1 + k] / d; G[3 * k + i] += g[1 + k] / d;
}
__host__ __device__ inline void so3_log(const float *R, float *phi) {
    const float ca = fminf(fmaxf(0.5f * (R[0] + R[4] + R[8]) - 0.5f, -1.f), 1.f);
    const float th = acosf(ca);
    if (fabsf(th) < SO3_TOL) {          // vee(R - I)
        phi[0] = R[7]; phi[1] = R[2]; phi[2] = R[3];
        return;
    }
    const float f = 0.5f * th / sinf(th);
    phi[0] = f * (R[7] - R[5]); phi[1] = f * (R[2] - R[6]); phi[2] = f * (R[3] - R[1]);
}

// status bits
constexpr int ST_NONFINITE = 1, ST_DET_STEP = 2, ST_DET_CHAIN = 4, ST_INVALID_GT = 8;
// torch.isclose(det, 1) with its defaults rtol 1e-5, atol 1e-8 (trainer.py:341,347)
__host__ __device__ __forceinline__ bool det_close(float d) { return fabsf(d - 1.f) <= 1e-8f + 1e-5f; }

// one chain step shared by the forward kernel and the backward kernel's recomputation; Rn = normalised R_new
__host__ __device__ __forceinline__ void chain_step(const float *R_old, const float *phi, float *Rc, float *R_new, float *Rn,
                                           bool &normalised) {
    so3_exp(phi, Rc);
    m3_mul(R_old, Rc, R_new);
    for (int i = 0; i < 9; ++i) Rn[i] = R_new[i];
    normalised = !so3_valid(Rn);
    if (normalised) so3_project(Rn);
}


// forward chain of one sample: x, w [S,3] -> ox [S,3], oq [S,4]; returns the status flags
__host__ __device__ inline int chain_fwd_sample(const float *x, const float *w, int S, float *ox, float *oq) {
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, t[3] = {0.f, 0.f, 0.f};
    int st = 0;
    for (int s = 0; s < S; ++s) {
        const float *xs = x + s * 3, *ws = w + s * 3;
        if (!finite3(xs) || !finite3(ws)) st |= ST_NONFINITE;
        float Rc[9], Rnew[9], Rn[9], q[4];
        bool normalised;
        chain_step(R, ws, Rc, Rnew, Rn, normalised);
        if (!det_close(m3_det(Rc))) st |= ST_DET_STEP;
        const float t0 = R[0] * xs[0] + R[1] * xs[1] + R[2] * xs[2] + t[0];
        const float t1 = R[3] * xs[0] + R[4] * xs[1] + R[5] * xs[2] + t[1];
        const float t2 = R[6] * xs[0] + R[7] * xs[1] + R[8] * xs[2] + t[2];
        t[0] = t0; t[1] = t1; t[2] = t2;
        for (int i = 0; i < 9; ++i) R[i] = Rnew[i];
        if (!det_close(m3_det(R))) st |= ST_DET_CHAIN;
        so3_to_quat(Rn, q);
        ox[s * 3] = t[0]; ox[s * 3 + 1] = t[1]; ox[s * 3 + 2] = t[2];
        oq[s * 4] = q[0]; oq[s * 4 + 1] = q[1]; oq[s * 4 + 2] = q[2]; oq[s * 4 + 3] = q[3];
    }
    return st;
}

// backward of one sample: gx [S,3], gq [S,4] (either may be NULL) -> dx, dw [S,3]; recomputes the chain
__host__ __device__ inline void chain_bwd_sample(const float *x, const float *w, int S, const float *gx, const float *gq,
                                                 float *dx, float *dw) {
    float Rs[CHAIN_MAX_S][9];             // R_prev BEFORE step s (local memory; S <= CHAIN_MAX_S)
    {
        float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        for (int s = 0; s < S; ++s) {
            for (int i = 0; i < 9; ++i) Rs[s][i] = R[i];
            float Rc[9], Rnew[9];
            so3_exp(w + s * 3, Rc);
            m3_mul(R, Rc, Rnew);
            for (int i = 0; i < 9; ++i) R[i] = Rnew[i];
        }
    }
    float GR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, Gt[3] = {0.f, 0.f, 0.f};
    for (int s = S - 1; s >= 0; --s) {
        const float *xs = x + s * 3, *ws = w + s * 3, *R_old = Rs[s];
        float Rc[9], Rnew[9], Rn[9], q[4];
        bool normalised;
        chain_step(R_old, ws, Rc, Rnew, Rn, normalised);
        const int branch = so3_to_quat(Rn, q);
        float Gq[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (gq) so3_to_quat_bwd(Rn, branch, q, gq + s * 4, Gq);
        if (normalised) {
            // differential of the projection at a (nearly) orthogonal matrix: dRn = Rn skew(Rn^T dR); self-adjoint
            float A[9], K[9];
            m3_tmul(Rn, Gq, A);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) K[3 * i + j] = 0.5f * (A[3 * i + j] - A[3 * j + i]);
            m3_mul(Rn, K, Gq);
        }
        for (int i = 0; i < 9; ++i) GR[i] += Gq[i];
        if (gx) {
            Gt[0] += gx[s * 3];
            Gt[1] += gx[s * 3 + 1];
            Gt[2] += gx[s * 3 + 2];
        }
        // t_new = R_old x + t_old;  R_new = R_old Rc
        dx[s * 3 + 0] = R_old[0] * Gt[0] + R_old[3] * Gt[1] + R_old[6] * Gt[2];
        dx[s * 3 + 1] = R_old[1] * Gt[0] + R_old[4] * Gt[1] + R_old[7] * Gt[2];
        dx[s * 3 + 2] = R_old[2] * Gt[0] + R_old[5] * Gt[1] + R_old[8] * Gt[2];
        float dRc[9], GRold[9], dphi[3];
        m3_tmul(R_old, GR, dRc);
        so3_exp_bwd(ws, dRc, dphi);
        dw[s * 3 + 0] = dphi[0]; dw[s * 3 + 1] = dphi[1]; dw[s * 3 + 2] = dphi[2];
        m3_mult(GR, Rc, GRold);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) GR[3 * i + j] = GRold[3 * i + j] + Gt[i] * xs[j];
    }
}

// ground truth of one (sample, pair): gi, gj, g0 = the 15-float records of frames i, j and 0
__host__ __device__ inline int gt_relative_sample(const float *gi, const float *gj, const float *g0, float *o, float *p) {
    int st = 0;
    for (int i = 0; i < 12; ++i)
        if (!isfinite(gi[i]) || !isfinite(gj[i]) || !isfinite(g0[i])) st |= ST_NONFINITE;
    float R[9], d[3], phi[3], q[4];
    // frame to frame: T_i^-1 T_j
    m3_tmul(gi + 3, gj + 3, R);
    d[0] = gj[0] - gi[0]; d[1] = gj[1] - gi[1]; d[2] = gj[2] - gi[2];
    o[0] = gi[3] * d[0] + gi[6] * d[1] + gi[9] * d[2];
    o[1] = gi[4] * d[0] + gi[7] * d[1] + gi[10] * d[2];
    o[2] = gi[5] * d[0] + gi[8] * d[1] + gi[11] * d[2];
    if (!so3_valid(R)) st |= ST_INVALID_GT;      // the reference's from_matrix(normalize=False) raises
    so3_log(R, phi);
    o[3] = phi[0]; o[4] = phi[1]; o[5] = phi[2];
    // frame to start: T_0^-1 T_j
    m3_tmul(g0 + 3, gj + 3, R);
    d[0] = gj[0] - g0[0]; d[1] = gj[1] - g0[1]; d[2] = gj[2] - g0[2];
    p[0] = g0[3] * d[0] + g0[6] * d[1] + g0[9] * d[2];
    p[1] = g0[4] * d[0] + g0[7] * d[1] + g0[10] * d[2];
    p[2] = g0[5] * d[0] + g0[8] * d[1] + g0[11] * d[2];
    if (!so3_valid(R)) st |= ST_INVALID_GT;
    so3_to_quat(R, q);
    p[3] = q[0]; p[4] = q[1]; p[5] = q[2]; p[6] = q[3];
    return st;
}

}  // namespace dlio
