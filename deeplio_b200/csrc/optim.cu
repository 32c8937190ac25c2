// Fused Adam with L2 weight decay over one flat parameter arena (torch.optim.Adam semantics,
// reference: deeplio/models/optimizer.py:10).  One launch updates every parameter of the model:
// 16 B read + 12 B written per element, HBM-bound.
#include "common.cuh"

namespace dlio {

// The update follows torch.optim.Adam's single-tensor path operation by operation (torch/optim/adam.py):
//   g' = g * grad_scale + wd * p;  m <- m + (1 - b1) (g' - m)  [lerp];  v <- v * b2 + ((1 - b2) g') g'  [mul, addcmul];
//   p <- p + (-step_size) * (m / (sqrt(v) / bc2_sqrt + eps))   [addcdiv]
// with 1 - b1, 1 - b2, step_size = lr / (1 - b1^t) and bc2_sqrt = sqrt(1 - b2^t) evaluated in double on the host and
// rounded once (torch evaluates them in Python floats): computing 1.f - 0.999f on the device is off by 4.7e-5.
struct AdamK {
    float b2, omb1, omb2, step, bc2_sqrt, eps, wd, grad_scale;
};
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const AdamK &k) {
    const float gr = g * k.grad_scale + k.wd * p;
    m = m + k.omb1 * (gr - m);
    v = v * k.b2 + (k.omb2 * gr) * gr;
    p = p - k.step * (m / (sqrtf(v) / k.bc2_sqrt + k.eps));
}
__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                   float *__restrict__ m, float *__restrict__ v, long long n, AdamK k) {
    long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        if (i + 3 < n) {
            float4 pp = ld4(p + i), gg = ld4(g + i), mm = ld4(m + i), vv = ld4(v + i);
            adam_one(pp.x, gg.x, mm.x, vv.x, k);
            adam_one(pp.y, gg.y, mm.y, vv.y, k);
            adam_one(pp.z, gg.z, mm.z, vv.z, k);
            adam_one(pp.w, gg.w, mm.w, vv.w, k);
            st4(p + i, pp);
            st4(m + i, mm);
            st4(v + i, vv);
        } else {
            for (long long j = i; j < n; ++j) adam_one(p[j], g[j], m[j], v[j], k);
        }
    }
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n,
                              double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                              double grad_scale, void *stream) {
    ProfScope prof_(DLIO_PROF_OPTIM, (cudaStream_t)stream);
    DLIO_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad argument");
    DLIO_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                   "adam_step: pointers must be 16-byte aligned");
    AdamK k;
    k.b2 = (float)beta2;
    k.omb1 = (float)(1.0 - beta1);
    k.omb2 = (float)(1.0 - beta2);
    k.step = (float)(lr / (1.0 - pow(beta1, (double)step)));
    k.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    k.eps = (float)eps;
    k.wd = (float)weight_decay;
    k.grad_scale = (float)grad_scale;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, k);
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
