// Fused Adam with L2 weight decay over one flat parameter arena (torch.optim.Adam semantics,
// reference: deeplio/models/optimizer.py:10).  One launch updates every parameter of the model:
// 16 B read + 12 B written per element, HBM-bound.
#include "common.cuh"

namespace dlio {

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                   float *__restrict__ m, float *__restrict__ v, long long n,
                                                   float lr, float b1, float b2, float eps, float wd,
                                                   float bc1, float bc2_sqrt, float grad_scale) {
    long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    const float step = lr / bc1;
    for (; i < n; i += stride) {
        if (i + 3 < n) {
            float4 pp = ld4(p + i), gg = ld4(g + i), mm = ld4(m + i), vv = ld4(v + i);
            float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
            float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float gr = ga[j] * grad_scale + wd * pa[j];
                ma[j] = b1 * ma[j] + (1.f - b1) * gr;
                va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
                pa[j] -= step * ma[j] / (sqrtf(va[j]) / bc2_sqrt + eps);
            }
            st4(p + i, make_float4(pa[0], pa[1], pa[2], pa[3]));
            st4(m + i, make_float4(ma[0], ma[1], ma[2], ma[3]));
            st4(v + i, make_float4(va[0], va[1], va[2], va[3]));
        } else {
            for (long long k = i; k < n; ++k) {
                float gr = g[k] * grad_scale + wd * p[k];
                float mk = b1 * m[k] + (1.f - b1) * gr;
                float vk = b2 * v[k] + (1.f - b2) * gr * gr;
                m[k] = mk;
                v[k] = vk;
                p[k] -= step * mk / (sqrtf(vk) / bc2_sqrt + eps);
            }
        }
    }
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                              float grad_scale, void *stream) {
    ProfScope prof_(DLIO_PROF_OPTIM, (cudaStream_t)stream);
    DLIO_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad argument");
    DLIO_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                   "adam_step: pointers must be 16-byte aligned");
    float bc1 = 1.f - powf(beta1, (float)step);
    float bc2 = sqrtf(1.f - powf(beta2, (float)step));
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                               eps, weight_decay, bc1, bc2, grad_scale);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

// ------------------------------------------------------------------ pose loss (HWSLoss, frame-to-frame part)
namespace dlio {
// loss = mse(pos, gt_pos) * exp(-sx) + sx + mse(ori, gt_ori) * exp(-sq) + sq   (losses.py:68-86, "local" terms)
// single block: the tensors are [B, S, 3]
__global__ void __launch_bounds__(256) hws_loss_kernel(const float *__restrict__ pos, const float *__restrict__ ori,
                                                       const float *__restrict__ gpos, const float *__restrict__ gori,
                                                       int n, float sx, float sq, float *loss, float *dpos, float *dori) {
    __shared__ float red[2][8];
    const float wx = expf(-sx), wq = expf(-sq);
    float st = 0.f, sw = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float a = pos[i] - gpos[i], b = ori[i] - gori[i];
        st += a * a;
        sw += b * b;
        if (dpos) dpos[i] = 2.f * a * wx / (float)n;
        if (dori) dori[i] = 2.f * b * wq / (float)n;
    }
    st = warp_sum(st);
    sw = warp_sum(sw);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = st;
        red[1][threadIdx.x >> 5] = sw;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) {
            a += red[0][w];
            b += red[1][w];
        }
        loss[0] = a / (float)n * wx + sx + b / (float)n * wq + sq;
    }
}
}  // namespace dlio

extern "C" int dlio_hws_loss(const float *pos, const float *ori, const float *gt_pos, const float *gt_ori, int n,
                             float sx, float sq, float *loss, float *dpos, float *dori, void *stream) {
    ProfScope prof_(DLIO_PROF_OPTIM, (cudaStream_t)stream);
    DLIO_CHECK_ARG(pos && ori && gt_pos && gt_ori && loss && n > 0, "hws_loss: bad argument");
    dlio::hws_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pos, ori, gt_pos, gt_ori, n, sx, sq, loss, dpos, dori);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
