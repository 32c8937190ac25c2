// Caller-side glue of the train step, on the device: what the reference does
// around `model(...)` in Python loops over (batch, pair) with ~20 tiny kernels and two host synchronisations per
// pair -- pose chaining se(3) -> SE(3) (trainer.py:324-351), ground-truth pairing (misc.py:83-125), the NaN / Inf
// guards (trainer.py:221-243), HWSLoss / LWSLoss (losses.py:11-96) -- as a handful of single launches.  All of it
// is tiny (B x S <= a few hundred 3-vectors): one thread per sample walks the S pairs; nothing here is
// bandwidth- or FLOP-relevant, the point is launch count and the absence of host round trips.
//
// SO(3) maps follow liegroups.torch.SO3 (what the reference calls; absent from this image, restated in
// oracle/pose_oracle.py): exp = Rodrigues with a first-order branch for |phi| < 1e-6, log through acos of the trace,
// matrix -> quaternion (wxyz) in four branches, SVD projection of matrices that fail the 1e-6 validity test
// (here: Newton iteration on the polar decomposition, which converges to the same U V^T).
#include "pose_math.cuh"

namespace dlio {

__global__ void se3_chain_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, int B, int S,
                                     float *__restrict__ ox, float *__restrict__ oq, int *status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t o = (size_t)b * S;
    const int st = chain_fwd_sample(x + o * 3, w + o * 3, S, ox + o * 3, oq + o * 4);
    if (st && status) atomicOr(status, st);
}

__global__ void se3_chain_bwd_kernel(const float *__restrict__ x, const float *__restrict__ w, int B, int S,
                                     const float *__restrict__ gx, const float *__restrict__ gq,
                                     float *__restrict__ dx, float *__restrict__ dw) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t o = (size_t)b * S;
    chain_bwd_sample(x + o * 3, w + o * 3, S, gx ? gx + o * 3 : nullptr, gq ? gq + o * 4 : nullptr, dx + o * 3, dw + o * 3);
}

// ground-truth pairing (misc.py:83-125): gts [B, F, 15] = (t 3, R 9 row-major, v 3)
struct Combos {
    int n;
    int idx[2 * CHAIN_MAX_S];
};
__global__ void gt_relative_kernel(const float *__restrict__ gts, int B, int F, Combos cb, float *__restrict__ f2f,
                                   float *__restrict__ f2g, int *status) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= B * cb.n) return;
    const int b = id / cb.n, s = id - b * cb.n;
    const float *g0 = gts + (size_t)b * F * 15;
    const int st = gt_relative_sample(g0 + (size_t)cb.idx[2 * s] * 15, g0 + (size_t)cb.idx[2 * s + 1] * 15, g0,
                                      f2f + (size_t)id * 6, f2g + (size_t)id * 7);
    if (st && status) atomicOr(status, st);
}

// HWSLoss / LWSLoss and their gradients, one block (losses.py:11-96).  Four squared-error terms (t, w: frame to frame;
// p, q: frame to start), each a strided 3-D view [B, G, C] of its prediction and ground-truth tensors, so the slices
// the trainer takes (trainer.py:246-263) need no copies.
struct PoseLoss {
    dlio_loss_term term[4];
    const float *sx, *sq, *upstream;
    int lws;
    float beta;
    float *loss, *dsx, *dsq;
};
__global__ void __launch_bounds__(256) pose_loss_kernel(PoseLoss a) {
    __shared__ float red[4][8];
    __shared__ float tot[4];
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const dlio_loss_term &t = a.term[k];
        if (!t.pred) continue;
        const int gc = t.G * t.C, n = t.B * gc;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int b = e / gc, r = e - b * gc, g = r / t.C, c = r - g * t.C;
            const float d = t.pred[b * t.pred_sb + g * t.pred_ss + c] - t.gt[b * t.gt_sb + g * t.gt_ss + c];
            s[k] += d * d;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float v = warp_sum(s[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[threadIdx.x][i];
        // mean over the element count torch.nn.functional.mse_loss sees (an empty slice gives 0 / 0 = NaN there too)
        const dlio_loss_term &t = a.term[threadIdx.x];
        tot[threadIdx.x] = t.pred ? v / (float)(t.B * t.G * t.C) : 0.f;
    }
    __syncthreads();
    const float Lx = tot[0] + tot[2], Lr = tot[1] + tot[3];      // (L_t + L_p), (L_w + L_q)
    const float wx = a.lws ? 1.f : expf(-*a.sx), wq = a.lws ? a.beta : expf(-*a.sq);
    const float up = a.upstream ? *a.upstream : 1.f;
    if (threadIdx.x == 0) {
        if (a.loss) a.loss[0] = a.lws ? Lx + a.beta * Lr : Lx * wx + *a.sx + Lr * wq + *a.sq;
        if (!a.lws && a.dsx) a.dsx[0] = up * (1.f - Lx * wx);
        if (!a.lws && a.dsq) a.dsq[0] = up * (1.f - Lr * wq);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const dlio_loss_term &t = a.term[k];
        if (!t.pred || !t.dpred) continue;
        const int gc = t.G * t.C, n = t.B * gc;
        const float f = up * 2.f * ((k & 1) ? wq : wx) / (float)n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int b = e / gc, r = e - b * gc, g = r / t.C, c = r - g * t.C;
            t.dpred[b * t.d_sb + g * t.d_ss + c] =
                f * (t.pred[b * t.pred_sb + g * t.pred_ss + c] - t.gt[b * t.gt_sb + g * t.gt_ss + c]);
        }
    }
}

// one pass over up to 8 tensors: flag bit t is raised when tensor t holds a NaN or an infinity
struct FiniteArgs {
    const float *p[8];
    long long n[8];
};
__global__ void __launch_bounds__(256) finite_check_kernel(FiniteArgs a, int *flags) {
    const int t = blockIdx.y;
    const float *p = a.p[t];
    const long long n = a.n[t];
    bool bad = false;
    const long long stride = (long long)gridDim.x * blockDim.x, i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if ((((uintptr_t)p) & 15) == 0) {
        const long long n4 = n >> 2;
        for (long long i = i0; i < n4; i += stride) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + 4 * i);
            bad |= ((v.x & 0x7f800000u) == 0x7f800000u) | ((v.y & 0x7f800000u) == 0x7f800000u) |
                   ((v.z & 0x7f800000u) == 0x7f800000u) | ((v.w & 0x7f800000u) == 0x7f800000u);
        }
        for (long long i = n4 * 4 + i0; i < n; i += stride) bad |= (__float_as_uint(p[i]) & 0x7f800000u) == 0x7f800000u;
    } else {
        for (long long i = i0; i < n; i += stride) bad |= (__float_as_uint(p[i]) & 0x7f800000u) == 0x7f800000u;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 1 << t);
}

// pairing gather (misc.py:37-42,65-69 + lidar_feat_nets.py:216-218): frames [B, F, *, H, W] (element strides sb, sf,
// sc; H and W contiguous) -> padded NHWC [B*S, hp, wp, dst.c] with channels (t0: c0..c0+C-1, t1: c0..c0+C-1, zeros),
// the frames of pair s being combinations[s]; optional TF32 lo plane; optional NaN / Inf flag
struct PairGather {
    const float *src;
    long long sb, sf, sc;
    int S, C, c0;
    Combos cb;
    Geo d;
    float *hi, *lo;
    __half *h2;            // optional packed fp16 planes, per pixel [c hi | c lo], scaled from *bound
    const float *bound;
    int *flags;
    int flag_bit;
};
__global__ void __launch_bounds__(256) pair_gather_kernel(PairGather a) {
    const Geo &d = a.d;
    const float s16 = a.h2 ? f16_scale_from_bound(*a.bound) : 1.f;
    const unsigned total = (unsigned)d.n * d.hp * d.wp;
    bool bad = false;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int xx = (int)(i % (unsigned)d.wp);
        const unsigned r = i / (unsigned)d.wp;
        const int yy = (int)(r % (unsigned)d.hp), n = (int)(r / (unsigned)d.hp);
        const int h = yy - d.ph, w = xx - d.pw;
        const bool in = h >= 0 && h < d.h && w >= 0 && w < d.w;
        const int b = n / a.S, s = n - b * a.S;
        const float *base = a.src + (size_t)b * a.sb + (size_t)h * d.w + w;
        for (int c0 = 0; c0 < d.c; c0 += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                float x = 0.f;
                if (in && c < 2 * a.C) {
                    const int tt = c / a.C, cc = c - tt * a.C;
                    x = base[(size_t)a.cb.idx[2 * s + tt] * a.sf + (size_t)(a.c0 + cc) * a.sc];
                    bad |= (__float_as_uint(x) & 0x7f800000u) == 0x7f800000u;
                }
                v[j] = x;
            }
            const float4 v4 = make_float4(v[0], v[1], v[2], v[3]);
            if (a.hi) st4_split(a.hi, a.lo, (size_t)i * d.c + c0, v4);
            if (a.h2) st4_h2(a.h2, i, d.c, c0, v4, s16, 1);
        }
    }
    if (a.flags && __any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(a.flags, a.flag_bit);
}

}  // namespace dlio

using namespace dlio;

extern "C" int dlio_se3_chain_fwd(const float *f2f_x, const float *f2f_w, int B, int S, float *f2g_x, float *f2g_q,
                                  int *status, void *stream) {
    ProfScope prof_(DLIO_PROF_OPTIM, (cudaStream_t)stream);
    DLIO_CHECK_ARG(f2f_x && f2f_w && f2g_x && f2g_q && B > 0 && S > 0, "se3_chain_fwd: bad argument");
    se3_chain_fwd_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(f2f_x, f2f_w, B, S, f2g_x, f2g_q, status);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_se3_chain_bwd(const float *f2f_x, const float *f2f_w, int B, int S, const float *g_x,
                                  const float *g_q, float *d_x, float *d_w, void *stream) {
    ProfScope prof_(DLIO_PROF_OPTIM, (cudaStream_t)stream);
    DLIO_CHECK_ARG(f2f_x && f2f_w && d_x && d_w && B > 0 && S > 0, "se3_chain_bwd: bad argument");
    DLIO_CHECK_ARG(S <= CHAIN_MAX_S, "se3_chain_bwd: at most %d pairs per sample", CHAIN_MAX_S);
    se3_chain_bwd_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(f2f_x, f2f_w, B, S, g_x, g_q, d_x, d_w);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

static int fill_combos(Combos &cb, const int *combinations, int S, int frames, const char *who) {
    DLIO_CHECK_ARG(combinations && S > 0 && S <= CHAIN_MAX_S, "%s: 1 .. %d frame pairs per sample", who, CHAIN_MAX_S);
    cb.n = S;
    for (int i = 0; i < 2 * S; ++i) {
        DLIO_CHECK_ARG(combinations[i] >= 0 && combinations[i] < frames, "%s: combination index %d outside the %d frames",
                       who, combinations[i], frames);
        cb.idx[i] = combinations[i];
    }
    return DLIO_OK;
}

extern "C" int dlio_gt_relative(const float *gts, int B, int F, const int *combinations, int S, float *gt_f2f,
                                float *gt_f2g, int *status, void *stream) {
    ProfScope prof_(DLIO_PROF_OPTIM, (cudaStream_t)stream);
    DLIO_CHECK_ARG(gts && gt_f2f && gt_f2g && B > 0 && F > 0, "gt_relative: bad argument");
    Combos cb;
    int rc = fill_combos(cb, combinations, S, F, "gt_relative");
    if (rc) return rc;
    gt_relative_kernel<<<ceil_div((long long)B * S, 128), 128, 0, (cudaStream_t)stream>>>(gts, B, F, cb, gt_f2f, gt_f2g, status);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_pose_loss(dlio_loss_term t, dlio_loss_term w, dlio_loss_term p, dlio_loss_term q, const float *sx,
                              const float *sq, int lws, float beta, const float *upstream, float *loss, float *d_sx,
                              float *d_sq, void *stream) {
    ProfScope prof_(DLIO_PROF_OPTIM, (cudaStream_t)stream);
    PoseLoss a;
    a.term[0] = t; a.term[1] = w; a.term[2] = p; a.term[3] = q;
    bool any = false;
    for (int k = 0; k < 4; ++k) {
        const dlio_loss_term &x = a.term[k];
        if (!x.pred) continue;
        any = true;
        DLIO_CHECK_ARG(x.gt && x.B >= 0 && x.G >= 0 && x.C > 0 && (long long)x.B * x.G * x.C < (1LL << 30),
                       "pose_loss: bad term %d", k);
    }
    DLIO_CHECK_ARG(any && (loss || d_sx || d_sq || t.dpred || w.dpred || p.dpred || q.dpred), "pose_loss: nothing to do");
    DLIO_CHECK_ARG(lws || (sx && sq), "pose_loss: HWSLoss needs sx and sq (device scalars)");
    a.sx = sx; a.sq = sq; a.upstream = upstream; a.lws = lws; a.beta = beta; a.loss = loss; a.dsx = d_sx; a.dsq = d_sq;
    pose_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_finite_check(const float *const *tensors, const long long *sizes, int count, int *flags,
                                 void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(tensors && sizes && flags && count >= 1 && count <= 8, "finite_check: 1 .. 8 tensors");
    FiniteArgs a;
    long long nmax = 0;
    for (int i = 0; i < count; ++i) {
        DLIO_CHECK_ARG(tensors[i] && sizes[i] > 0, "finite_check: tensor %d is empty", i);
        a.p[i] = tensors[i];
        a.n[i] = sizes[i];
        nmax = sizes[i] > nmax ? sizes[i] : nmax;
    }
    long long blocks = (nmax / 4 + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    finite_check_kernel<<<dim3((unsigned)blocks, (unsigned)count), 256, 0, (cudaStream_t)stream>>>(a, flags);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}

extern "C" int dlio_pair_gather(const float *frames, long long sb, long long sf, long long sc, int B, int F,
                                const int *combinations, int S, int c0, int C, dlio_tensor4 dst, float *dst_ptr,
                                float *dst_lo, void *dst_h2, const float *bound, int *flags, int flag_bit,
                                void *stream) {
    ProfScope prof_(DLIO_PROF_ELEMENTWISE, (cudaStream_t)stream);
    DLIO_CHECK_ARG(frames && (dst_ptr || dst_h2) && valid_t4(dst) && dst.c % 4 == 0 && dst.c >= 2 * C && C > 0 &&
                       c0 >= 0 && dst.n == B * S && B > 0 && F > 0,
                   "pair_gather: bad argument");
    DLIO_CHECK_ARG((dst_ptr || !dst_lo) && (!dst_h2 || (bound && (((uintptr_t)dst_h2) & 15) == 0)),
                   "pair_gather: dst_lo needs dst_ptr, fp16 planes need their bound");
    PairGather a;
    int rc = fill_combos(a.cb, combinations, S, F, "pair_gather");
    if (rc) return rc;
    a.src = frames; a.sb = sb; a.sf = sf; a.sc = sc; a.S = S; a.C = C; a.c0 = c0;
    a.d = Geo(dst); a.hi = dst_ptr; a.lo = dst_lo; a.flags = flags; a.flag_bit = flag_bit;
    a.h2 = (__half *)dst_h2; a.bound = bound;
    const long long total = (long long)a.d.n * a.d.hp * a.d.wp;
    DLIO_CHECK_ARG(total < (1LL << 32), "pair_gather: tensor too large");
    long long blocks = (total + 255) / 256;
    blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
    pair_gather_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    DLIO_LAUNCH_CHECK();
    return DLIO_OK;
}
