/*
 * deeplio_b200 -- C ABI of the B200-native (sm_100a) DeepLIO training hot path.
 *
 * The reference (ArashJavan/DeepLIO) has no FFI: its hot path is four nn.Module subsystems under
 * deeplio/models/nets that dispatch to ATen / cuDNN / cuBLAS.  Each entry point below replaces one
 * family of those implicit library calls; the reference call sites are cited per function.  The
 * Python host side (deeplio_b200/nets/) mirrors the reference's module interface on top of this
 * ABI; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - all floating-point data is IEEE fp32; statistics accumulators are fp64; the operands of the fp16
 *     tensor-core kernels are "packed split planes" derived from fp32 tensors (see "fp16 operand split" below);
 *   - activations are "padded NHWC": memory [n][h + 2*ph][w + 2*pw][c] with ZERO pad rows / columns,
 *     described by dlio_tensor4 (the pads are those of the convolution that consumes the tensor);
 *   - convolution weights are OHWI: [cout][kh][kw][cin] (a channels_last nn.Conv2d weight, zero-copy);
 *   - no allocation, no ownership transfer, no host synchronisation inside any call: work is enqueued
 *     on `stream` (a cudaStream_t passed as void*); scratch memory is passed in by the caller;
 *   - return value: 0 on success, a negative dlio_status otherwise; dlio_last_error() gives the text.
 */
#ifndef DEEPLIO_B200_H
#define DEEPLIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLIO_ABI_VERSION 7

typedef enum {
    DLIO_OK = 0,
    DLIO_ERR_INVALID = -1,     /* bad argument / unsupported shape */
    DLIO_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
    DLIO_ERR_WORKSPACE = -3,   /* caller-provided scratch too small */
    DLIO_ERR_UNSUPPORTED = -4  /* device is not sm_100 */
} dlio_status;

/* activation / epilogue functions (torch.nn.functional names) */
typedef enum {
    DLIO_ACT_NONE = 0,
    DLIO_ACT_RELU = 1,
    DLIO_ACT_LEAKY = 2,   /* leaky_relu, slope 0.01 */
    DLIO_ACT_SIGMOID = 3,
    DLIO_ACT_TANH = 4
} dlio_act;

/* padded NHWC activation descriptor */
typedef struct {
    int n, h, w, c; /* logical extent */
    int ph, pw;     /* zero rows / columns stored on each side */
} dlio_tensor4;

/* 2-D convolution (cross-correlation, dilation 1, groups 1) */
typedef struct {
    int kh, kw; /* kernel */
    int sh, sw; /* stride */
    int ph, pw; /* zero padding */
} dlio_conv;

/* ------------------------------------------------------------------ fp16 operand split ("3xF16")
 * The tensor-core convolutions compute fp32-accurate products from fp16 operands: a tensor v[rows][C] (rows =
 * padded pixels, or output channels for weights) is stored as ONE plane of halves, row r = [C hi | C lo],
 *     hi = fp16(s*v),  lo = fp16((s*v - hi) * 2^11),  s = 2^(14 - ceil(log2(bound)))   (|v| <= bound)
 * and the kernels accumulate hi*hi and (lo*hi + hi*lo) in separate fp32 accumulators, then undo s and 2^11.
 * `bound` is ONE float in device memory per tensor, written by the kernel that produces the tensor's statistics
 * (dlio_bn_finalize, dlio_bn_bwd_apply, dlio_weight_pack_f16) and read by the kernels that write or consume the
 * planes -- the scale never visits the host.  A void* named *_h2 is such a plane; *_bound is its bound. */

/* ------------------------------------------------------------------ library */
int dlio_abi_version(void);
const char *dlio_last_error(void);                 /* host string, thread-local */
int dlio_device_check(int device);                 /* DLIO_OK iff compute capability 10.x */
/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
long long dlio_launch_count(void);
/* Per-kernel-class device timing for bench.py's roofline: while enabled, every launch of a profiled class
 * is bracketed by CUDA events on its stream.  dlio_profile_enable clears earlier records;
 * dlio_profile_read synchronises the recorded events of one class and returns their summed duration. */
typedef enum {
    DLIO_PROF_CONV_FWD_SIMT = 0, DLIO_PROF_CONV_DGRAD_SIMT = 1, DLIO_PROF_CONV_WGRAD_SIMT = 2,
    DLIO_PROF_CONV_FWD_TC = 3, DLIO_PROF_CONV_DGRAD_TC = 4, DLIO_PROF_CONV_WGRAD_TC = 5,
    DLIO_PROF_ELEMENTWISE = 6, /* BN / pool / SE / packing / element-wise passes (norm_pool.cu) */
    DLIO_PROF_DENSE = 7, DLIO_PROF_RNN = 8, DLIO_PROF_OPTIM = 9
} dlio_prof_kind;
/* Tuning switches (A/B measurements and tests; the defaults are what ships): "pool_tma" 0/1 -- pooling passes staged
 * through shared memory by bulk copies (default 1; environment DLIO_POOL_TMA), "ew_block" 64..256 -- block-size cap of
 * the element-wise passes (default 256; environment DLIO_EW_BLOCK), "conv_cg2" 0/1 -- stride-1 fp16 convolutions with 128-
 * channel output tiles on CTA pairs (tcgen05 cta_group::2; default 1; environment DLIO_CONV_CG2), "nvtx" 0/1 -- an NVTX
 * range named dlio/<kernel class> around the launches of every entry point (default 0; environment DLIO_NVTX; the
 * Python engine adds a range per layer, e.g. encoder1.conv3, when it is on), "bwd_single_pass" 0/1 -- MEASUREMENT ONLY:
 * dgrad and wgrad issue the hi*hi product alone (plain fp16 / TF32 operand accuracy, the reference's own cudnn.allow_tf32
 * level) instead of the three products of the split scheme; default 0, the gradient parity tests fail with it on;
 * "apply_rows" 0/1 -- row-structured kernel for the BN-apply passes without pooling (default 1). */
int dlio_set_option(const char *name, int value);
/* Caller-owned scratch of one call, in bytes -- the library never allocates device memory.  One
 * query for every entry point that takes scratch:
 *   "conv2d_fwd"   {Cout}                  the fp64 `stats` buffer (sum | sum of squares per channel; zeroed by the caller)
 *   "bn_bwd"       {C}                     the fp64 `sums` buffer of dlio_bn_act_pool_bwd_reduce / dlio_pool_bwd_sums
 *   "rnn_fwd"      {kind,L,D,B,T,I,H}      `reserve` of dlio_rnn_fwd (= 4 * dlio_rnn_reserve_floats)
 *   "rnn_bwd"      {kind,L,D,B,T,I,H}      `scratch` of dlio_rnn_bwd
 *   "scan_project" {H,W}                   `scratch` of dlio_scan_project
 * Returns (size_t)-1 and sets dlio_last_error for an unknown op or a wrong number of dims. */
size_t dlio_workspace_bytes(const char *op, const long long *dims, int ndims);
int dlio_profile_enable(int on);
int dlio_profile_read(int kind, double *total_ms, long long *launches);

/* ------------------------------------------------------------------ input staging
 * Replaces imgs.reshape(b*s, t*c, h, w) (lidar_feat_nets.py:216-218): gathers the strided
 * [N, T, C, H, W] view (element strides sn, st, sc; h and w contiguous) into padded NHWC with
 * `dst.c` >= T*C channels (extra channels zero), channel order (t0:c0..c2, t1:c0..c2).  dst_lo (optional): the
 * low-order TF32 plane for the tensor-core first layer. */
int dlio_pack_input(const float *src, long long sn, long long st, long long sc, int T, int C,
                    dlio_tensor4 dst, float *dst_ptr, float *dst_lo, void *stream);

/* ------------------------------------------------------------------ convolution
 * Replaces aten::conv2d forward / dgrad / wgrad behind every nn.Conv2d on the path
 * (lidar_feat_nets.py:248-257,279-301; pointseg_net.py:18; pointseg_modules.py:96-104; resnet.py:36;
 * torchvision resnet.py:40-56).
 *
 * forward:  y = act(conv(x, w) + bias); optionally accumulates per-channel sum / sum-of-squares of y
 *           over the logical (n,h,w) extent into stats[0:cout] / stats[cout:2*cout] (fp64, caller
 *           zeroes) -- the batch statistics nn.BatchNorm2d needs (train mode).
 *           x_lo / w_lo: optional low-order TF32 planes (v - trunc_tf32(v); the *_hi planes always hold the
 *           full fp32 values).  When both are given, the stride is 1, cin % 32 == 0, cout % 16 == 0, the
 *           kernel is "same" (k = 2*pad + 1) and x is stored with pads >= the conv pads, the tcgen05
 *           (3xTF32) implicit-GEMM kernel is used; otherwise the generic fp32 kernel runs on x_hi and w_hi.
 */
int dlio_conv2d_fwd(dlio_tensor4 x, const float *x_hi, const float *x_lo,
                    const float *w_hi, const float *w_lo, const float *bias,
                    dlio_conv cv, int act,
                    dlio_tensor4 y, float *y_ptr, double *stats, void *stream);

/* dgrad: dx = conv_transpose(dy, w).  dx pads are written as zeros when dx is padded.
 * wt_hi / wt_lo (optional): the flipped / transposed weights of dlio_weight_flip_transpose as TF32 split
 * planes; when they and dy_lo are given, the stride is 1, dy is stored with pads >= (kh-1-ph, kw-1-pw) and
 * cout % 32 == 0, cin % 16 == 0, the dgrad runs on the tcgen05 kernel as a convolution of dy with wt. */
/* accumulate != 0: dx += the gradient (a tensor with several consumers collects its contributions without a temporary
 * and an add pass); 0: dx is overwritten. */
int dlio_conv2d_bwd_data(dlio_tensor4 dy, const float *dy_hi, const float *dy_lo,
                         const float *w_hi, const float *w_lo, const float *wt_hi, const float *wt_lo,
                         dlio_conv cv, dlio_tensor4 dx, float *dx_ptr, int accumulate, void *stream);

/* wgrad: dw[cout][kh][kw][cin] = sum_{n,h,w} dy * x (overwrites dw). */
int dlio_conv2d_bwd_weight(dlio_tensor4 x, const float *x_hi, const float *x_lo,
                           dlio_tensor4 dy, const float *dy_hi, const float *dy_lo,
                           dlio_conv cv, float *dw, void *stream);

/* OIHW (torch default) <-> OHWI with channel padding cin -> cin_pad (zeros); optional hi/lo split */
int dlio_weight_to_ohwi(const float *w_oihw, int cout, int cin, int kh, int kw, int cin_pad,
                        float *w_hi, float *w_lo, void *stream);
int dlio_weight_grad_to_oihw(const float *dw_ohwi, int cout, int cin, int kh, int kw, int cin_pad,
                             float *dw_oihw, void *stream);
/* First-layer convolutions (cin <= 8, kw <= 7 odd, stride_w 1 or 2) as a stride-1 kh x 3 convolution over the
 * space-to-depth views [n, h, w/4, 32] -> [n, h, w/4, R*cout], R = 4/stride_w, of the same memory (the input must
 * be stored with 8 channels and row pads of 4 pixels; see csrc/conv_s2d.cu): weights rearranged to
 * [R*cout][kh][3][32] (+ lo plane), their gradient folded back to OIHW, and the R-fold BN statistics summed. */
int dlio_weight_to_s2d(const float *w_oihw, int cout, int cin, int kh, int kw, int sw, float *w4_hi, float *w4_lo,
                       void *stream);
int dlio_weight_grad_from_s2d(const float *dw4, int cout, int cin, int kh, int kw, int sw, float *dw_oihw,
                              void *stream);
int dlio_fold_stats(const double *in, int r, int c, double *out, void *stream);
/* dgrad operand for the tcgen05 path: wt[cin][kh'][kw'][cout] = w[cout][kh-1-kh'][kw-1-kw'][cin] */
int dlio_weight_flip_transpose(const float *w_ohwi, int cout, int cin, int kh, int kw,
                               float *wt_hi, float *wt_lo, void *stream);

/* fp16 tensor-core path (tcgen05 kind::f16, three products per fp32 product).  Same semantics as the three
 * calls above for stride-1 "same" convolutions with cin % 64 == 0 (wgrad also: cout % 64 == 0; dgrad: cout % 64
 * == 0, cin % 16 == 0), operands given as packed split planes; there is no fallback: an inapplicable problem
 * returns DLIO_ERR_INVALID.  dlio_conv2d_fwd_f16 also takes cv.sh == 2 (cv.sw == 1): the convolution is computed
 * over the whole input grid and only the even rows are stored and counted in the statistics (y.h = (x.h + 2 ph -
 * kh) / 2 + 1) -- with the pixel-pair view below this is how the reference's stride-(1,2) and (2,2) convolutions
 * (FlowNet lidar_feat_nets.py:248-257, ResNet first blocks resnet.py:40-48) reach the tensor cores.
 *   dlio_weight_pack_f16: OIHW fp32 -> packed OHWI rows [cout][2][kh*kw*cin_pad] (transpose_flip == 0), or the
 *   dgrad operand [cin_pad][2][kh*kw*cout], wt[ci][kh'][kw'][co] = w[co][ci][kh-1-kh'][kw-1-kw'] (transpose_flip
 *   != 0).  compute_bound != 0: first reduces max |w| into *w_bound; otherwise *w_bound is read. */
int dlio_weight_pack_f16(const float *w_oihw, int cout, int cin, int kh, int kw, int cin_pad, int transpose_flip,
                         int compute_bound, float *w_bound, void *w_h2, void *stream);
/* Pixel-pair view of a W-stride-2 convolution (kw in {1,3,5}, pad (kw-1)/2): on the SAME memory, x [n,h,w,c] with
 * even w and even row pads is x2 [n,h,w/2,2c] (fp16 planes in the pixel-pair layout, dlio_bnpool.out_group == 2) and
 * the layer is a stride-1 convolution with kernel kh x kw2 (kw2 = 3, or 1 for kw == 1) and weights
 *     w2[co][dy][t][p*c + ci] = w[co][ci][dy][2 (t - (kw2-1)/2) + p + (kw-1)/2]   (zero where that column is outside)
 * dlio_weight_pack_pair_f16 packs w2 as dlio_weight_pack_f16 packs w (rows [cout][2][kh*kw2*2cin], or the dgrad
 * operand [2cin][2][kh*kw2*cout]); dlio_weight_grad_from_pair folds dw2 [cout][kh][kw2][2cin] back to OIHW. */
int dlio_weight_pack_pair_f16(const float *w_oihw, int cout, int cin, int kh, int kw, int transpose_flip,
                              int compute_bound, float *w_bound, void *w_h2, void *stream);
int dlio_weight_grad_from_pair(const float *dw2, int cout, int cin, int kh, int kw, float *dw_oihw, void *stream);
int dlio_conv2d_fwd_f16(dlio_tensor4 x, const void *x_h2, const float *x_bound, const void *w_h2,
                        const float *w_bound, const float *bias, dlio_conv cv, int act, dlio_tensor4 y,
                        float *y_ptr, double *stats, void *stream);
int dlio_conv2d_bwd_data_f16(dlio_tensor4 dy, const void *dy_h2, const float *dy_bound, const void *wt_h2,
                             const float *w_bound, dlio_conv cv, dlio_tensor4 dx, float *dx_ptr, int accumulate,
                             void *stream);
int dlio_conv2d_bwd_weight_f16(dlio_tensor4 x, const void *x_h2, const float *x_bound, dlio_tensor4 dy,
                               const void *dy_h2, const float *dy_bound, dlio_conv cv, float *dw, void *stream);
/* fp32 [rows][c] -> packed split plane, computing the bound (max |v|) first; test / staging helper */
int dlio_pack_f16(const float *src, long long rows, int c, float *bound, void *dst_h2, void *stream);
/* max |x| over n floats -> *bound (device float; written, not accumulated). */
int dlio_absmax(const float *x, long long n, float *bound, void *stream);

/* First layer on the fp16 tensor-core path ("folded split"; replaces the same reference calls as dlio_conv2d_fwd for
 * the first convolution of every encoder: lidar_feat_nets.py:306 conv1, :248 FlowNet conv1, pointseg_net.py:20 conv1a).
 * The 8-channel input planes hold per pixel [8 hi | 8 lo] halves, so FOUR consecutive pixels are one 128-byte row of
 * 64 halves with K index k = p * 16 + part * 8 + c.  With weight planes B_main[k] = (part == 0 ? w_hi : 0) and
 * B_corr[k] = (part == 0 ? w_lo : w_hi) (dlio_weight_to_s2d_f16; rows (r, co) of the R = 4 / sw outputs computed from
 * a group, columns (dy, t, k)), x . B_main = x_hi w_hi and x . B_corr = x_hi w_lo + x_lo w_hi: the two accumulators of
 * the split scheme from two products and one plane of x.
 *   x: the [n, h, w/4, 64] view of the input planes (pads ph, 1: the planes are stored with row pads of 4 pixels);
 *   y: the [n, h, w/4, R * cout] view of the fp32 output [n, h, w / sw, cout]; stats: 2 * R * cout (dlio_fold_stats).
 * Weight gradient: dy is the [n, h, w/4, R * 64] view (pads as x) of planes stored per OUTPUT pixel as [64 hi | 64 lo]
 * (cout == 64), dw64 [R * 64][kh][3][64] fp32, folded back to OIHW by dlio_weight_grad_from_s2d_f16. */
int dlio_weight_to_s2d_f16(const float *w_oihw, int cout, int cin, int kh, int kw, int sw, float *w_bound, void *w4_h2,
                           void *stream);
int dlio_weight_grad_from_s2d_f16(const float *dw64, int cout, int cin, int kh, int kw, int sw, float *dw_oihw,
                                  void *stream);
int dlio_conv2d_fwd_f16_folded(dlio_tensor4 x, const void *x_h2, const float *x_bound, const void *w4_h2,
                               const float *w_bound, const float *bias, dlio_conv cv, int act, dlio_tensor4 y,
                               float *y_ptr, double *stats, void *stream);
int dlio_conv2d_bwd_weight_f16_folded(dlio_tensor4 x, const void *x_h2, const float *x_bound, dlio_tensor4 dy,
                                      const void *dy_h2, const float *dy_bound, dlio_conv cv, float *dw64, void *stream);

/* ------------------------------------------------------------------ batch-norm / activation / pooling
 * Replaces aten::batch_norm (train + eval), relu, max_pool2d_with_indices, adaptive_avg_pool2d and the
 * residual / bypass adds around them (lidar_feat_nets.py:306-342; base_net.py:55-71; pointseg_modules.py
 * :110-141; torchvision resnet.py:88-103).
 *
 * dlio_bn_finalize: from the fp64 sums over `count` elements per channel computes
 *   mean, invstd = 1/sqrt(biased_var + eps), scale = gamma*invstd, shift = beta - mean*scale and updates
 *   running_mean / running_var (unbiased var, `momentum`).  With use_running != 0 (eval mode) scale/shift
 *   come from the running statistics and nothing is updated.  Outputs are [c] fp32 each.
 *   out_bound (optional, needs stats): an upper bound of |scale*y + shift| over the tensor (+ *res_bound when
 *   a residual will be added), from |y - mean_b| <= sqrt(count * var_b); it scales the fp16 planes that
 *   dlio_bn_act_pool_fwd writes.
 *   num_batches_tracked (optional, device int64, the BatchNorm2d buffer): incremented in train mode by the kernel
 *   itself (no separate add launch); momentum < 0 means nn.BatchNorm2d(momentum=None): the cumulative average
 *   with factor 1 / num_batches_tracked.
 */
int dlio_bn_finalize(const double *stats, long long count, int c, const float *gamma, const float *beta,
                     float *running_mean, float *running_var, float momentum, float eps, int use_running,
                     float *mean, float *invstd, float *scale, float *shift, const float *res_bound,
                     float *out_bound, long long *num_batches_tracked, void *stream);

typedef struct {
    int relu;       /* 1: apply ReLU after scale*y + shift (+ residual if res_mode == 1) */
    int res_mode;   /* 0 none; 1 residual added BEFORE the ReLU (torchvision BasicBlock, resnet.py:101);
                       2 residual added AFTER it (Fire bypass, pointseg_modules.py:138-140) */
    int pool_k;     /* 1 = no pooling, 3 = 3x3 max pool (pad 1; the output extent carries ceil_mode) */
    int pool_sh, pool_sw;
    int c_off;      /* channel offset inside the output (and residual) tensor (Fire concat) */
    int out_group;  /* layout of out_h2: 0 / 1 plain rows [C hi | C lo] per pixel; 2 pixel pairs
                       [p0 C hi | p1 C hi | p0 C lo | p1 C lo] -- the operand layout of a W-stride-2 convolution run
                       through the pixel-pair view [n, h, w/2, 2C] (dlio_weight_pack_pair_f16); needs an even padded width */
    int zero_tail;  /* dlio_bn_act_pool_fwd, pool_k == 1, c_off == 0: the output has more channels than y (a narrow
                       Fire squeeze output allocated with a multiple of 64 channels); write zeros to channels
                       [y.c, out.c) in the same pass instead of a memset of the whole tensor before it */
} dlio_bnpool;

/* out[n,ho,wo,c_off+c] = maxpool(act(scale[c]*y + shift[c] (+res)) (+res)); writes out's pads as zeros for
 * the channel range.  scale == shift == NULL means identity (plain max-pool / copy into a padded tensor).
 * out_lo (optional): low-order TF32 plane, v - trunc_tf32(v) (out_hi always receives the full fp32 value).
 * out_h2 / out_bound (optional): packed fp16 split planes on out's padded grid, scaled from *out_bound; out_hi
 * may then be NULL (no fp32 copy is written).  pool_idx (uint8 [n,ho,wo,c], required when pooling is
 * differentiated): window-relative arg-max, first maximum wins (torch tie-break).  pool_ymax (optional, fp32
 * [n,ho,wo,c], 3x3 pool without residual): the conv output y at the arg-max, for dlio_pool_bwd_sums. */
int dlio_bn_act_pool_fwd(dlio_tensor4 y, const float *y_ptr, const float *scale, const float *shift,
                         dlio_tensor4 res, const float *res_ptr, dlio_bnpool p, dlio_tensor4 out,
                         float *out_hi, float *out_lo, void *out_h2, const float *out_bound,
                         uint8_t *pool_idx, float *pool_ymax, void *stream);

typedef enum { DLIO_GRAD_DIRECT = 0, DLIO_GRAD_POOL = 1, DLIO_GRAD_AVG = 2 } dlio_grad_src;

/* backward pass 1: dz = relu'(.) * pool_backward(dout) at every (n,h,w,c) of y; writes dz (unpadded
 * [n,h,w,c]) and accumulates sums[0:c] += sum dz, sums[c:2c] += sum dz * yhat (fp64, caller zeroes; NULL
 * when there is no BN).  grad_src: DIRECT / POOL  dout is an (unpadded or padded) NHWC gradient of the
 * op's output, read at channel c_off; AVG  dout is [n, ld_dout] read at c_off and spread as 1/(h*w).
 * dres (optional, unpadded [n,h,w,dres_c], written at channel c_off): gradient of the residual input;
 * overwritten, or added to when dres_accumulate != 0.  sums_absmax != 0: sums has 2c+1 entries and sums[2c]
 * receives max |dz| (what dlio_bn_bwd_apply needs to scale fp16 planes of dy). */
int dlio_bn_act_pool_bwd_reduce(dlio_tensor4 y, const float *y_ptr, const float *scale, const float *shift,
                                const float *mean, const float *invstd,
                                dlio_tensor4 res, const float *res_ptr, dlio_bnpool p, int grad_src,
                                dlio_tensor4 dout, const float *dout_ptr, int ld_dout, const uint8_t *pool_idx,
                                float *dz, float *dres, int dres_c, int dres_accumulate, double *sums,
                                int sums_absmax, void *stream);

/* backward pass 2: dy = scale * (dz - mean(dz) - yhat * mean(dz*yhat)) (batch_stats != 0) or scale * dz
 * (eval-mode BN); if pre_relu (conv -> ReLU -> BN, lidar_feat_nets.py:308-309) dy *= (y > 0).  Writes dy
 * (geometry dy_t, zero pads, optional TF32 split), dgamma[c], dbeta[c] and accumulates
 * dbias_sums[c] += sum dy (fp64, caller zeroes; may be NULL).  dy_h2 / dy_bound (optional; needs sums[2c] =
 * max |dz|): dy as packed fp16 split planes, and the bound it was scaled from (OUTPUT, for dgrad / wgrad);
 * dy_hi may then be NULL.  dy_t.h == y.h: dy on y's own grid.  dy_t.h == the input height of an H-stride-2
 * convolution ((dy_t.h - 1) / 2 + 1 == y.h): y row i is written to dy row 2 i and the odd rows are zero, which is
 * the dy operand of that convolution's backward run as a stride-1 convolution over its input grid.
 * post_relu != 0 (needs `shift`): `dz` is the gradient of the ReLU OUTPUT (BN -> ReLU layers, base_net.py:55-71) and
 * the kernel masks it with scale*y + shift > 0 itself -- pass 1 then runs with dz == NULL (sums only) and the masked
 * gradient never goes through HBM.  dy_t.c > y.c (multiple of 4): channels [y.c, dy_t.c) of dy are written as zeros,
 * so that the backward of a narrow convolution (Fire squeeze, 16 / 48 / 80 output channels) runs on the tensor cores
 * against weights padded with zero rows. */
int dlio_bn_bwd_apply(dlio_tensor4 y, const float *y_ptr, const float *dz, const double *sums,
                      long long count, const float *scale, const float *mean, const float *invstd,
                      int pre_relu, int batch_stats, dlio_tensor4 dy_t, float *dy_hi, float *dy_lo,
                      void *dy_h2, float *dy_bound, float *dgamma, float *dbeta, double *dbias_sums,
                      const float *shift, int post_relu, void *stream);
/* The two backward passes of a layer  conv -> (ReLU) -> BN -> 3x3 max-pool  (no ReLU / residual between BN and pool:
 * Simple-1, lidar_feat_nets.py:306-322) without materialising dz:
 *   dlio_pool_bwd_sums: the BN-backward sums from the POOLED side -- every dout is routed to exactly one input, so
 *   sums[0:c] += sum dout, sums[c:2c] += sum dout * yhat(arg-max) with y at the arg-max saved by the forward pass
 *   (pool_ymax); sums[2c] = windows * max |dout| >= max |dz| (`windows`: how many pooling windows can contain one
 *   input position: 3 or 2 per axis for stride 1 or 2).  8 bytes per pooled element.
 *   dlio_bn_pool_bwd_apply: dlio_bn_bwd_apply with dz un-pooled on the fly from (dout, pool_idx) through pool p.
 * Against pass 1 + pass 2 above this saves writing and re-reading dz and one read of y per conv-output element. */
int dlio_pool_bwd_sums(dlio_tensor4 dout, const float *dout_ptr, int c_off, int c, const float *ymax,
                       const float *mean, const float *invstd, int windows, double *sums, void *stream);
int dlio_bn_pool_bwd_apply(dlio_tensor4 y, const float *y_ptr, dlio_bnpool p, dlio_tensor4 dout,
                           const float *dout_ptr, const uint8_t *pool_idx, const double *sums, long long count,
                           const float *scale, const float *mean, const float *invstd, int pre_relu,
                           int batch_stats, dlio_tensor4 dy_t, float *dy_hi, float *dy_lo, void *dy_h2,
                           float *dy_bound, float *dgamma, float *dbeta, double *dbias_sums, void *stream);
int dlio_f64_to_f32(const double *src, float *dst, int n, void *stream);

/* global average pooling (adaptive_avg_pool2d((1,1)), lidar_feat_nets.py:84-85,131,175,340; SE squeeze,
 * pointseg_modules.py:217): out[n, c_off + c] = mean_{h,w} act(scale*x + shift)  (scale/shift may be NULL) */
int dlio_spatial_mean_fwd(dlio_tensor4 x, const float *x_ptr, const float *scale, const float *shift,
                          int relu, float *out, int ld_out, int c_off, void *stream);
/* out[n, c] = sum_{h,w} a * b   (gradient of the SE gate) */
int dlio_spatial_dot(dlio_tensor4 a, const float *a_ptr, dlio_tensor4 b, const float *b_ptr, float *out,
                     void *stream);
/* SE layer (pointseg_modules.py:203-221): out = x * gate[n,c];
 * backward: dx = dout * gate[n,c] + dmean[n,c] / hw  (dout, dx unpadded [n,hw,c]; dmean may be NULL) */
int dlio_channel_scale_fwd(dlio_tensor4 x, const float *x_ptr, const float *gate, dlio_tensor4 out,
                           float *out_hi, float *out_lo, void *stream);
int dlio_channel_scale_bwd(const float *dout, const float *gate, const float *dmean, int n, int hw, int c,
                           float *dx, void *stream);
/* element-wise helpers: out = alpha*a + beta*b (16-byte aligned; out may alias a or b), out = a * b,
 * out[a,c] = sum_t x[a,t,c] (ImuFeatFC time sum, imu_feat_nets.py:50), Bernoulli keep mask scaled by 1/(1-p)
 * (counter-based generator; seed_epoch, optional, is a counter in DEVICE memory mixed into the seed at run time, so a
 * launch captured in a CUDA graph draws a fresh mask on every replay once the graph's owner advances it) */
int dlio_axpby(const float *a, float alpha, const float *b, float beta, float *out, long long n, void *stream);
int dlio_sum_mid(const float *x, float *out, long long a, int t, int c, void *stream);
int dlio_mul(const float *a, const float *b, float *out, long long n, void *stream);
/* dst[r * ldd + c] = src[r * lds + c] for a [rows, cols] block (element strides): torch.cat along the last axis
 * (fusion_nets.py:27,70,75), torch.stack of the per-window IMU features (imu_feat_nets.py:81-83) and their backward. */
int dlio_copy2d(const float *src, long long lds, float *dst, long long ldd, long long rows, int cols, void *stream);
int dlio_dropout_mask(float *mask, long long n, float p, unsigned long long seed,
                      const unsigned long long *seed_epoch, void *stream);

/* ------------------------------------------------------------------ dense layers (small M)
 * Replaces aten::linear / addmm behind nn.Linear (fc1, IMU-FC stack, soft fusion, Odom-FC, heads:
 * lidar_feat_nets.py:97,144,185,233; imu_feat_nets.py:47; fusion_nets.py:68-69; odom_feat_nets.py:34;
 * deeplio_nets.py:88-89).
 *   y[m, ldy] = act(x[m, ldx] . w[n, k]^T + b)                        (forward)
 *   dz = dy * act'(y);  dx = dz . w;  dw = dz^T . x;  db = sum_m dz   (backward; dx/dw/db may be NULL)
 * `scratch` for the backward: m*n floats. */
int dlio_linear_fwd(const float *x, int ldx, const float *w, const float *b, int m, int n, int k, int act,
                    float *y, int ldy, void *stream);
int dlio_linear_bwd(const float *x, int ldx, const float *w, const float *y, int ldy, const float *dy, int lddy,
                    int m, int n, int k, int act, float *dx, int lddx, float *dw, float *db,
                    float *scratch, void *stream);

/* ------------------------------------------------------------------ recurrent layers
 * Replaces aten::lstm / aten::gru (cuDNN RNN) behind nn.LSTM / nn.GRU, batch_first, multi-layer,
 * optionally bidirectional (imu_feat_nets.py:64-82; odom_feat_nets.py:61-80).
 *
 * kind: 0 = LSTM (gates i,f,g,o), 1 = GRU (gates r,z,n).  G = 4 or 3, D = 1 or 2.
 * weights: array (HOST memory) of 4*L*D device pointers in torch order
 *          [w_ih, w_hh, b_ih, b_hh] for (l0,fwd), (l0,rev), (l1,fwd), ...
 * x [B,T,I]; h0/c0 [L*D,B,H] (NULL = zeros); out [B,T,D*H]; hn/cn [L*D,B,H].
 * reserve: caller scratch kept from forward to backward, dlio_rnn_reserve_floats() floats.
 * drop_mask: optional [L-1, B, T, D*H] pre-scaled keep mask applied to inter-layer activations.
 */
size_t dlio_rnn_reserve_floats(int kind, int L, int D, int B, int T, int I, int H);
int dlio_rnn_fwd(int kind, int L, int D, int B, int T, int I, int H, const float *const *weights,
                 const float *x, const float *h0, const float *c0, const float *drop_mask,
                 float *out, float *hn, float *cn, float *reserve, void *stream);
/* grads: array (HOST) of 4*L*D device pointers, same order as weights; each is OVERWRITTEN.
 * dout [B,T,D*H] (NULL = zeros); dhn/dcn [L*D,B,H] (NULL = zeros); dx [B,T,I] (may be NULL);
 * dh0/dc0 [L*D,B,H] are outputs (and serve as the recurrent accumulators). */
int dlio_rnn_bwd(int kind, int L, int D, int B, int T, int I, int H, const float *const *weights,
                 const float *x, const float *drop_mask, const float *dout, const float *dhn,
                 const float *dcn, const float *reserve, float *const *grads, float *dx, float *dh0,
                 float *dc0, float *scratch, size_t scratch_floats, void *stream);
size_t dlio_rnn_bwd_scratch_floats(int kind, int L, int D, int B, int T, int I, int H);


/* ------------------------------------------------------------------ optimizer
 * Fused Adam with L2 weight decay over a flat fp32 arena (torch.optim.Adam semantics; the reference builds it
 * at deeplio/models/optimizer.py:10): g = grad*grad_scale + wd*p; m,v moments; bias-corrected update, operation by
 * operation as torch's single-tensor Adam; the hyper-parameters are doubles (Python floats) so that 1 - beta and the
 * bias corrections are rounded once, as torch does.  `step` is the 1-based step count. */
int dlio_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n,
                   double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                   double grad_scale, void *stream);
/* Frame-to-frame part of HWSLoss (deeplio/losses/losses.py:68-86) with fixed sx, sq, and its gradient:
 * loss = mse(pos, gt_pos) e^-sx + sx + mse(ori, gt_ori) e^-sq + sq over n = B*S*3 elements; dpos / dori
 * (optional) receive d loss / d pos, d loss / d ori. */
int dlio_hws_loss(const float *pos, const float *ori, const float *gt_pos, const float *gt_ori, int n,
                  float sx, float sq, float *loss, float *dpos, float *dori, void *stream);

/* Full HWSLoss / LWSLoss and their gradients (deeplio/losses/losses.py:11-96 as trainer.py:246-263 calls them).
 * Four mean-squared-error terms: t, w (frame to frame: translation, so(3) rotation) and p, q (frame to start:
 * position, quaternion), each described by a dlio_loss_term -- a strided 3-D view [B, G, C] of the prediction and of
 * the ground truth (last dimension contiguous), so the slices the trainer takes need no copies; pred == NULL
 * switches a term off (the reference's loss_Types).
 *   lws == 0 (HWSLoss): loss = (L_p + L_t) e^-sx + sx + (L_q + L_w) e^-sq + sq, sx / sq DEVICE scalars (the
 *   nn.Parameters); lws != 0 (LWSLoss): loss = (L_p + L_t) + beta (L_q + L_w).
 * Outputs, each optional: loss[1]; term.dpred = upstream * d loss / d pred (written through d_sb / d_ss); d_sx, d_sq.
 * upstream (optional DEVICE scalar, default 1): the gradient flowing into the loss. */
typedef struct {
    const float *pred, *gt;
    float *dpred;
    int B, G, C;
    int pred_sb, pred_ss, gt_sb, gt_ss, d_sb, d_ss;   /* element strides of the first two dimensions */
} dlio_loss_term;
int dlio_pose_loss(dlio_loss_term t, dlio_loss_term w, dlio_loss_term p, dlio_loss_term q, const float *sx,
                   const float *sq, int lws, float beta, const float *upstream, float *loss, float *d_sx, float *d_sq,
                   void *stream);

/* ------------------------------------------------------------------ caller-side glue (N1, N2)
 * Pose chaining, Trainer.se3_to_SE3 (deeplio/models/trainer.py:324-351): per sample b, R_0 = I, t_0 = 0 and for
 * s = 0 .. S-1:  t <- R t_s + t,  R <- R exp(w_s);  f2g_x[b,s] = t,  f2g_q[b,s] = quaternion (w,x,y,z) of R (projected
 * onto SO(3) when it fails liegroups' 1e-6 validity test).  The reference runs a Python loop over B x S with ~20
 * tiny kernels and two torch.det host synchronisations per pair; here one launch.  *status (optional, device int,
 * caller zeroes) receives OR-ed flags instead of the reference's exceptions: 1 a non-finite input, 2 det(exp(w))
 * not close to 1 (trainer.py:341), 4 det(R) not close to 1 (:347).  S <= 64.
 * bwd: d_x, d_w [B,S,3] from g_x [B,S,3], g_q [B,S,4] (either may be NULL = zeros); recomputes the chain. */
int dlio_se3_chain_fwd(const float *f2f_x, const float *f2f_w, int B, int S, float *f2g_x, float *f2g_q, int *status,
                       void *stream);
int dlio_se3_chain_bwd(const float *f2f_x, const float *f2f_w, int B, int S, const float *g_x, const float *g_q,
                       float *d_x, float *d_w, void *stream);
/* Ground-truth pairing, DataCombiCreater.process_ground_turth (deeplio/models/misc.py:83-125): gts [B,F,15] =
 * (t 3, R 9 row-major, v 3) per frame (kitti.py:292-301), combinations: HOST array of S (i, j) frame-index pairs.
 * gt_f2f[b,s] = (R_i^T (t_j - t_i), log(R_i^T R_j)), gt_f2g[b,s] = (R_0^T (t_j - t_0), quaternion(R_0^T R_j)).
 * *status flags: 1 non-finite, 8 a relative rotation fails the validity test (the reference raises there). */
int dlio_gt_relative(const float *gts, int B, int F, const int *combinations, int S, float *gt_f2f, float *gt_f2g,
                     int *status, void *stream);
/* The NaN / Inf guards of the train loop (trainer.py:221-229,240-243: six isnan().any() / isinf().any() pairs, each
 * a reduction plus a host synchronisation) as ONE pass: tensors / sizes are HOST arrays of `count` <= 8 device
 * pointers / element counts; bit t of *flags (device int, caller zeroes) is raised when tensor t is not finite. */
int dlio_finite_check(const float *const *tensors, const long long *sizes, int count, int *flags, void *stream);
/* Pairing gather, DataCombiCreater.process_images (misc.py:65-69: imgs[:, combinations], channel split) fused
 * with the encoders' input reshape (lidar_feat_nets.py:216-218): frames [B, F, *, H, W] (element strides sb, sf,
 * sc; H, W contiguous) -> padded NHWC dst with dst.n = B*S images and channels (frame combinations[s][0]: c0 ..
 * c0+C-1, frame combinations[s][1]: c0 .. c0+C-1, zeros up to dst.c).  No [B,S,2,C,H,W] copy of the pairs is made.
 * dst_ptr / dst_lo (optional): fp32 plane and its TF32 low-order plane.  dst_h2 (optional): packed fp16 planes, per
 * pixel [dst.c hi | dst.c lo], scaled from *bound (an upper bound of max|frames| in device memory: dlio_absmax) --
 * the operand of dlio_conv2d_fwd_f16_folded.  flags (optional): bit `flag_bit` is raised on a NaN / Inf input. */
int dlio_pair_gather(const float *frames, long long sb, long long sf, long long sc, int B, int F,
                     const int *combinations, int S, int c0, int C, dlio_tensor4 dst, float *dst_ptr, float *dst_lo,
                     void *dst_h2, const float *bound, int *flags, int flag_bit, void *stream);

/* ------------------------------------------------------------------ LiDAR / IMU preprocessing (N4)
 * dlio_scan_project: one velodyne frame -> range image channels, replacing LaserScan.open_scan's depth filter +
 * do_range_projection + do_normal_projection (deeplio/common/laserscan.py:86-91,122-191,215-248) and the image assembly
 * / mean subtraction / channel selection of KittiRawData.get_velo_image + Kitti.transform_images (deeplio/datasets/
 * kitti.py:83-97,345-364).  points4 [n_points, 4] = (x, y, z, remission) as in the .bin files (utils.py:168-171),
 * 16-byte aligned.  The 8 channels are (x, y, z) / max_depth, remission, normal (3), range; `channels` (HOST array)
 * selects n_channels of them (config.yaml `channels`), mean8 (HOST, 8 floats, may be NULL) is subtracted in
 * out_normed.  Outputs, each optional: out_org / out_normed [n_channels, H, W] planar; out_idx [H, W] = index of the
 * winning point in points4 (-1: empty pixel).  The nearest point wins a pixel (lower index on an exact depth tie).
 * scratch: dlio_scan_scratch_bytes(H, W) bytes (the 64-bit z-buffer), 8-byte aligned. */
size_t dlio_scan_scratch_bytes(int H, int W);
int dlio_scan_project(const float *points4, int n_points, int H, int W, float fov_up_deg, float fov_down_deg,
                      float min_depth, float max_depth, const int *channels, int n_channels, const float *mean8,
                      void *scratch, float *out_org, float *out_normed, int *out_idx, void *stream);
/* IMU windowing, Kitti.load_imus + transform_imus (kitti.py:317-343,366-368): ts [m] (sorted, seconds, fp64) and imu
 * [m, 6] = (ax, ay, az, wx, wy, wz) of the OXTS stream, velo_ts [n_frames] the LiDAR timestamps; window w holds the
 * samples with velo_ts[w] <= ts < velo_ts[w+1], zero-padded / truncated to T rows, then (v - mean6) / std6 (HOST
 * arrays, may both be NULL) applied to every row, padding included, as the reference does.  out [n_frames-1, T, 6];
 * valid [n_frames-1] (optional): 0 where a window is empty. */
int dlio_imu_windows(const double *ts, const float *imu, int m, const double *velo_ts, int n_frames, int T,
                     const float *mean6, const float *std6, float *out, int *valid, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPLIO_B200_H */
