// tcgen05 implicit-GEMM convolution for sm_100a: stride-1 convolutions with Cin % 32 == 0, computed in
// error-compensated TF32 ("3xTF32": hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM), which keeps the
// result within ~1e-6 of an fp32 convolution (the 1e-4 pose gate rules out single-pass TF32 / BF16).
//
// Formulation ("shifted GEMM over the padded grid").  The input is a padded NHWC tensor whose zero pads are
// real memory, viewed as a matrix X[R = n*hp*wp, Cin].  For EVERY padded position q the kernel computes
//     Y[q, co] = sum_{tap=(dy,dx)} sum_ci X[q + (dy-ph)*wp + (dx-pw), ci] * W[co, tap, ci]
// which is the convolution at interior positions; rows that fall on pads are discarded in the epilogue
// (2-15 % extra MMA work, no im2col, no gather).  Each (tap, 32-channel chunk) is one pipeline stage: a plain
// 2-D TMA box [128 rows x 32 floats] of X at row offset q0 + shift (out-of-range rows are zero-filled by TMA
// and only feed discarded rows) and a [BN x 32] box of W, both 128B-swizzled, hi and lo planes.
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 tcgen05.mma issuer (one elected lane) + TMEM owner,
// warps 2-5 epilogue (TMEM -> registers -> bias / ReLU -> global, per-channel BN statistics).
// One 128 x BN output tile per CTA.
//
// Accumulation accuracy.  The tensor core adds each K=8 MMA into the fp32 accumulator with truncation, so the
// error of ONE accumulator grows linearly with the number of MMA steps (measured: 8e-9 * K relative, 3.7e-5 at
// K = 4608 -- too much for gradient parity).  The tile therefore owns FOUR accumulators of BN columns: the
// hi*hi products go round-robin (by K chunk) into three of them and the small lo*hi + hi*lo corrections into
// the fourth; the epilogue adds the four in fp32.  Each main accumulator sees 1/9 of the steps.
#include <cuda.h>

#include "common.cuh"
#include "conv.cuh"

namespace dlio {

constexpr int TC_BM = 128;       // output rows (padded pixels) per CTA
constexpr int TC_BK = 32;        // fp32 elements per K chunk = one 128-byte swizzle row (64 halves in f16 mode)
constexpr int TC_BK16 = 64;
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM_LIMIT = 200 * 1024;

struct TcArgs {
    Geo x;             // padded input geometry (pads are memory)
    Geo o;             // output geometry
    int kh, kw, ph, pw;
    int cin, cout, bn;  // bn: output channels per CTA (multiple of 16, <= 256)
    int act;
    int stages;
    int cluster;       // 1, or 2: CTA pairs along M share the weight tile through TMA multicast
    const float *bias;
    float *out;
    double *stats;
    long long rows;    // R = n * hp * wp
    int f16;           // 0: 3xTF32 on fp32 planes; 1: 3xF16 on packed half planes (K chunk = 64 halves)
    const float *xb, *wb;   // f16: device bounds of the two operands (-> power-of-two scales)
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask`, and each of those
// CTAs' mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// MMA kind and A-collector usage are compile-time; `f16` selects kind::f16 (fp16 operands, K = 16 per
// instruction) over kind::tf32 (K = 8) -- both consume 32 bytes of every operand row per instruction.
#define DLIO_UMMA_ASM(KIND, COLL)                                                                         \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                      \
                 "tcgen05.mma.cta_group::1.kind::" KIND COLL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), \
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)                                             \
                 : "memory")
// coll: 0 plain, 1 collector::a::fill, 2 collector::a::lastuse
template <int COLL>
__device__ __forceinline__ void umma(bool f16, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (f16) {
        if (COLL == 0) DLIO_UMMA_ASM("f16", "");
        else if (COLL == 1) DLIO_UMMA_ASM("f16", ".collector::a::fill");
        else DLIO_UMMA_ASM("f16", ".collector::a::lastuse");
    } else {
        if (COLL == 0) DLIO_UMMA_ASM("tf32", "");
        else if (COLL == 1) DLIO_UMMA_ASM("tf32", ".collector::a::fill");
        else DLIO_UMMA_ASM("tf32", ".collector::a::lastuse");
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile whose rows are 128 bytes: 8-row groups are 1024 bytes apart (SBO),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).  Advancing by one UMMA K step (8 tf32 = 32 B)
// adds 2 to the 16-byte-granular start address.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                        // version
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// fp32 accumulate, both operands K-major, M = 128; operand format TF32 (2) or F16 (0)
__device__ __forceinline__ uint32_t make_idesc(int n, bool f16) {
    const uint32_t fmt = f16 ? 0u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
               const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * 8 + 1];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red[2][256];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = TC_BM * TC_BK * 4;             // 16 KB per plane
    const uint32_t b_bytes = (uint32_t)a.bn * TC_BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const int S = a.stages;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), tfull = smem_u32(&bars[16]);

    const bool f16 = a.f16 != 0;
    const int bk = f16 ? TC_BK16 : TC_BK;                   // elements per 128-byte K chunk
    const int cchunks = a.cin / bk;
    const int iters = a.kh * a.kw * cchunks;
    const long long q0 = (long long)blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * a.bn;
    const int tmem_cols = 4 * a.bn <= 64 ? 64 : (4 * a.bn <= 128 ? 128 : (4 * a.bn <= 256 ? 256 : 512));
    const int nmain = iters < 3 ? iters : 3;
    uint32_t cta_rank = 0;
    if (a.cluster == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, (uint32_t)a.cluster);   // freed when every CTA of the pair has consumed it
        }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 512; i += TC_THREADS) red[i >> 8][i & 255] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (a.cluster == 2) cluster_sync_all();   // peer barriers must exist before multicast traffic / remote arrives
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const int tap = it / cchunks, cc = it - tap * cchunks;
                const int dy = tap / a.kw, dx = tap - dy * a.kw;
                const long long row = q0 + (long long)(dy - a.ph) * a.x.wp + (dx - a.pw);
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, stage_bytes);
                tma_load_2d(sa, &tm_xhi, fb, cc * bk, (int)row);
                tma_load_2d(sa + a_bytes, &tm_xlo, fb, cc * bk, (int)row);
                if (a.cluster == 2) {   // each CTA fetches half of the weight tile for both
                    const uint32_t half = b_bytes / 2;
                    const int nh = n0 + (int)cta_rank * (a.bn / 2);
                    tma_load_2d_mc(sa + 2 * a_bytes + cta_rank * half, &tm_whi, fb, tap * a.cin + cc * bk, nh, 3);
                    tma_load_2d_mc(sa + 2 * a_bytes + b_bytes + cta_rank * half, &tm_wlo, fb, tap * a.cin + cc * bk, nh, 3);
                } else {
                    tma_load_2d(sa + 2 * a_bytes, &tm_whi, fb, tap * a.cin + cc * bk, n0);
                    tma_load_2d(sa + 2 * a_bytes + b_bytes, &tm_wlo, fb, tap * a.cin + cc * bk, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(a.bn, f16);
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t d_ahi = make_kmajor_desc(sa), d_alo = make_kmajor_desc(sa + a_bytes);
                const uint64_t d_bhi = make_kmajor_desc(sa + 2 * a_bytes), d_blo = make_kmajor_desc(sa + 2 * a_bytes + b_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {                // 128-byte rows / 32 bytes per MMA K step
                    const uint64_t ko = (uint64_t)(k * 2);   // 32 bytes per K step, in 16-byte units
                    umma<0>(f16, tmem_base + 3 * a.bn, d_alo + ko, d_bhi + ko, idesc, (it | k) ? 1u : 0u);
                    umma<1>(f16, tmem_base + 3 * a.bn, d_ahi + ko, d_blo + ko, idesc, 1u);
                    umma<2>(f16, tmem_base + (it % 3) * a.bn, d_ahi + ko, d_bhi + ko, idesc, (it >= 3 || k) ? 1u : 0u);
                }
                // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                if (a.cluster == 2) umma_commit_mc(empty0 + 8 * s, 3);
                else umma_commit(empty0 + 8 * s);
            }
            umma_commit(tfull);                // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int quad = warp & 3;
        const int r = quad * 32 + lane;               // row of the tile == TMEM lane
        const long long q = q0 + r;
        bool valid = q < a.rows;
        size_t obase = 0;
        if (valid) {
            int xx = (int)(q % a.x.wp);
            long long t = q / a.x.wp;
            int yy = (int)(t % a.x.hp);
            int n = (int)(t / a.x.hp);
            int h = yy - a.x.ph, w = xx - a.x.pw;
            valid = h >= 0 && h < a.x.h && w >= 0 && w < a.x.w;
            if (valid) obase = a.o.off(n, h, w);
        }
        float *tr = reinterpret_cast<float *>(smem) + (size_t)(warp - 2) * 32 * 33;   // pipeline smem is free now
        // f16: undo the two power-of-two operand scales; the correction accumulator carries another 2^-11
        float inv_x = 1.f, inv_w = 1.f, corr = 1.f;
        if (f16) {
            inv_x = 1.f / f16_scale_from_bound(*a.xb);
            inv_w = 1.f / f16_scale_from_bound(*a.wb);
            corr = 1.f / 2048.f;
        }
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < a.bn; c0 += 32) {
            uint32_t v[32], u[32];
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
            tmem_ld32(trow, v);
            for (int m = 1; m < nmain; ++m) {
                tmem_ld32(trow + m * a.bn, u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
            tmem_ld32(trow + 3 * a.bn, u);
            const int ncol = min(32, a.bn - c0);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = fmaf(__uint_as_float(u[j]), corr, __uint_as_float(v[j])) * inv_x * inv_w;
                if (a.bias && j < ncol) x += a.bias[n0 + c0 + j];
                if (a.act == DLIO_ACT_RELU) x = fmaxf(x, 0.f);
                f[j] = valid ? x : 0.f;
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (j < ncol) st4(a.out + obase + n0 + c0 + j, make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
            }
            if (a.stats) {
                // column sums over the warp's 32 rows through a padded shared-memory transpose
#pragma unroll
                for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = f[j];
                __syncwarp();
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                    float x = tr[rr * 33 + lane];
                    s1 += x;
                    s2 = fmaf(x, x, s2);
                }
                __syncwarp();
                if (lane < ncol) {
                    atomicAdd(&red[0][c0 + lane], s1);
                    atomicAdd(&red[1][c0 + lane], s2);
                }
            }
        }
        if (a.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps
            for (int c = threadIdx.x - 64; c < a.bn; c += 128) {
                atomicAdd(a.stats + n0 + c, (double)red[0][c]);
                atomicAdd(a.stats + a.cout + n0 + c, (double)red[1][c]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (a.cluster == 2) cluster_sync_all();   // no CTA may exit while its peer can still signal its barriers
    else __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D tensor map over `rows` x `cols` elements (cols contiguous, rows `row_stride_elems` apart), box
// box_rows x (128 bytes of elements), zero OOB fill.  half_elems: fp16 elements (64 per box row) instead of fp32 (32).
static int make_map(CUtensorMap *m, const void *base, bool half_elems, long long rows, long long cols,
                    long long row_stride_elems, int box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("conv_tc: cuTensorMapEncodeTiled is not available from the driver");
        return DLIO_ERR_CUDA;
    }
    const size_t esz = half_elems ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, half_elems ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d half=%d", (int)r, rows, cols,
                  box_rows, (int)half_elems);
        return DLIO_ERR_CUDA;
    }
    return DLIO_OK;
}
// hi / lo maps of one operand: two fp32 planes, or the two halves of every row of a packed fp16 plane
static int make_map_pair(CUtensorMap *hi, CUtensorMap *lo, const float *p_hi, const float *p_lo, const __half *p_h2,
                         long long rows, long long cols, int box_rows,
                         CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    int rc;
    if (p_h2) {
        if ((rc = make_map(hi, p_h2, true, rows, cols, 2 * cols, box_rows, swizzle))) return rc;
        return make_map(lo, p_h2 + cols, true, rows, cols, 2 * cols, box_rows, swizzle);
    }
    if ((rc = make_map(hi, p_hi, false, rows, cols, cols, box_rows, swizzle))) return rc;
    return make_map(lo, p_lo, false, rows, cols, cols, box_rows, swizzle);
}

static int pick_bn(int cout) {
    for (int bn : {128, 64, 32, 16})   // four accumulators of bn columns must fit the 512 TMEM columns
        if (cout % bn == 0) return bn;
    return 0;
}

int conv_tc_fwd(const ConvArgs &a, int prof_kind, cudaStream_t st) {
    // applicability: split planes present, stride 1, Cin in 128-byte chunks, pads held in memory
    const bool f16 = a.x_h2 != nullptr;
    if (f16 ? !(a.w_h2 && a.x_bound && a.w_bound) : !(a.x_lo && a.w_lo)) return 0;
    if (a.sh != 1 || a.sw != 1) return 0;
    if (a.cin % (f16 ? TC_BK16 : TC_BK) != 0 || a.cout % 16 != 0) return 0;
    if (a.x.ph < a.ph || a.x.pw < a.pw) return 0;
    if (a.kh != 2 * a.ph + 1 || a.kw != 2 * a.pw + 1) return 0;   // "same" convolution: output extent == input extent
    const int bn = pick_bn(a.cout);
    if (!bn) return 0;
    const long long rows = (long long)a.x.n * a.x.hp * a.x.wp;
    if (rows >= (1LL << 31) - 4096) return 0;
    if ((((uintptr_t)a.x_hi | (uintptr_t)a.x_lo | (uintptr_t)a.w_hi | (uintptr_t)a.w_lo | (uintptr_t)a.out |
          (uintptr_t)a.x_h2 | (uintptr_t)a.w_h2) & 15) != 0) return 0;

    TcArgs t;
    t.f16 = f16 ? 1 : 0; t.xb = a.x_bound; t.wb = a.w_bound;
    t.x = a.x; t.o = a.o;
    t.kh = a.kh; t.kw = a.kw; t.ph = a.ph; t.pw = a.pw;
    t.cin = a.cin; t.cout = a.cout; t.bn = bn; t.act = a.act;
    t.bias = a.bias; t.out = a.out; t.stats = a.stats; t.rows = rows;
    const int stage_bytes = 2 * TC_BM * TC_BK * 4 + 2 * bn * TC_BK * 4;
    int stages = TC_SMEM_LIMIT / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) return 0;
    t.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;

    CUtensorMap mxh, mxl, mwh, mwl;
    const long long K = (long long)a.kh * a.kw * a.cin;
    int rc;
    if ((rc = make_map_pair(&mxh, &mxl, a.x_hi, a.x_lo, a.x_h2, rows, a.cin, TC_BM))) return rc;
    const long long mtiles = (rows + TC_BM - 1) / TC_BM;
    // CTA pairs that share the weight tile through TMA multicast are implemented and tested (set to 2), but
    // measured 6 % SLOWER (fwd 8.32 -> 8.81 ms / step): ncu sampling shows the producer waiting on `empty`, i.e.
    // the loop is bound by the tensor pipe reading its SS operands from shared memory (8 KB per 64-cycle MMA =
    // the full 128 B/clk), not by L2 -> SM traffic, and pairing adds lock-step between the two CTAs.
    constexpr bool kPairMulticast = false;
    t.cluster = (kPairMulticast && mtiles >= 2 && bn % 16 == 0) ? 2 : 1;
    if ((rc = make_map_pair(&mwh, &mwl, a.w_hi, a.w_lo, a.w_h2, a.cout, K, bn / t.cluster))) return rc;

    static bool attr_set = false;
    if (!attr_set) {
        DLIO_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
        attr_set = true;
    }
    ProfScope prof(prof_kind, st);
    if (a.o.ph > 0 || a.o.pw > 0) DLIO_CUDA(cudaMemsetAsync(a.out, 0, a.o.numel() * sizeof(float), st));
    dim3 grid((unsigned)((mtiles + t.cluster - 1) / t.cluster * t.cluster), (unsigned)(a.cout / bn));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = t.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DLIO_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel, mxh, mxl, mwh, mwl, t));
    DLIO_LAUNCH_CHECK();
    return 1;
}

// ==================================================================== wgrad
// dw[co, tap, ci] = sum_q dy[q, co] * x[q + shift(tap), ci] over the padded grid that x and dy share (dy's pad
// rows are zero, so they add nothing).  GEMM with M = co (128 per CTA), N = ci (BN per CTA), K = q.  Both
// operands are read pixel-major exactly as they lie in HBM, i.e. "MN-major" for the tensor core.  MN-major TF32
// operands exist only in the "128-byte swizzle with 32-byte atoms" layout (UMMA layout type 1, TMA swizzle
// 128B_ATOM_32B): an atom is 32 channels x 4 pixel rows (512 bytes).  A stage holds 32 pixel rows; each TMA box
// is [32 rows x 32 channels]; boxes of consecutive 32-channel groups are 4096 bytes apart (LBO), 4-row groups
// 512 bytes apart (SBO); one K = 8 MMA step spans two row groups (1024 bytes).
// MN-major FP16 operands (f16 mode) use the plain 128-byte swizzle (layout type 2 <-> TMA SWIZZLE_128B): an atom
// is 64 channels x 8 pixel rows (1024 bytes).  A stage holds 64 pixel rows; each TMA box is [64 rows x 64
// channels] = 8192 bytes (LBO), 8-row groups 1024 bytes apart (SBO); one K = 16 MMA step spans two row groups.
// grid = (K splits, taps * Cin/BN, Cout/128); partial tiles are combined with fp32 atomic adds into dw.
struct TcWgradArgs {
    int kh, kw, ph, pw, wp;
    int cin, cout, bn, stages;
    long long rows, rows_per_split;
    float *dw;
    int f16;
    const float *xb, *yb;   // f16: device bounds of x and dy
};

__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, bool f16) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((f16 ? 8192 : 4096) >> 4) << 16;   // leading byte offset: next channel atom along M / N
    d |= (uint64_t)((f16 ? 1024 : 512) >> 4) << 32;    // stride byte offset: next row group along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(f16 ? 2 : 1) << 61;                // SWIZZLE_128B / SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dyhi, const __grid_constant__ CUtensorMap tm_dylo,
                const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo, TcWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * 8 + 1];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = TC_BM * TC_BK * 4;             // dy tile: 32 rows x 128 channels
    const uint32_t b_bytes = (uint32_t)a.bn * TC_BK * 4;    // x tile:  32 rows x BN channels
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const int S = a.stages;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), tfull = smem_u32(&bars[16]);

    const int nblk = a.cin / a.bn;
    const int tap = blockIdx.y / nblk, ci0 = (blockIdx.y - tap * nblk) * a.bn;
    const int co0 = blockIdx.z * TC_BM;
    const int dy_ = tap / a.kw, dx_ = tap - dy_ * a.kw;
    const long long shift = (long long)(dy_ - a.ph) * a.wp + (dx_ - a.pw);
    const long long k_begin = (long long)blockIdx.x * a.rows_per_split;
    long long k_end = k_begin + a.rows_per_split;
    if (k_end > a.rows) k_end = a.rows;
    const bool f16 = a.f16 != 0;
    const int krows = f16 ? 64 : 32;                        // pixel rows (K) per stage
    const int bc = f16 ? 64 : 32;                           // channels per TMA box (128 bytes)
    const uint32_t box_bytes = (uint32_t)krows * 128;
    const int iters = k_end > k_begin ? (int)((k_end - k_begin + krows - 1) / krows) : 0;
    const int tmem_cols = 4 * a.bn <= 64 ? 64 : (4 * a.bn <= 128 ? 128 : (4 * a.bn <= 256 ? 256 : 512));
    const int nmain = iters < 3 ? iters : 3;
    if (iters == 0) return;   // uniform for the whole CTA

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            const int na_boxes = TC_BM / bc, nb_boxes = a.bn / bc;
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const long long q = k_begin + (long long)it * krows;
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, stage_bytes);
                for (int j = 0; j < na_boxes; ++j) {
                    tma_load_2d(sa + j * box_bytes, &tm_dyhi, fb, co0 + bc * j, (int)q);
                    tma_load_2d(sa + a_bytes + j * box_bytes, &tm_dylo, fb, co0 + bc * j, (int)q);
                }
                for (int j = 0; j < nb_boxes; ++j) {
                    tma_load_2d(sa + 2 * a_bytes + j * box_bytes, &tm_xhi, fb, ci0 + bc * j, (int)(q + shift));
                    tma_load_2d(sa + 2 * a_bytes + b_bytes + j * box_bytes, &tm_xlo, fb, ci0 + bc * j, (int)(q + shift));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // both operands MN-major: bits 15 and 16 of the instruction descriptor
            const uint32_t idesc = make_idesc(a.bn, f16) | (1u << 15) | (1u << 16);
            const uint64_t kstep = (uint64_t)((f16 ? 2048 : 1024) >> 4);   // one MMA K step = two row groups
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t d_ahi = make_mnmajor_desc(sa, f16), d_alo = make_mnmajor_desc(sa + a_bytes, f16);
                const uint64_t d_bhi = make_mnmajor_desc(sa + 2 * a_bytes, f16), d_blo = make_mnmajor_desc(sa + 2 * a_bytes + b_bytes, f16);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ko = (uint64_t)k * kstep;
                    umma<0>(f16, tmem_base + 3 * a.bn, d_alo + ko, d_bhi + ko, idesc, (it | k) ? 1u : 0u);
                    umma<1>(f16, tmem_base + 3 * a.bn, d_ahi + ko, d_blo + ko, idesc, 1u);
                    umma<2>(f16, tmem_base + (it % 3) * a.bn, d_ahi + ko, d_bhi + ko, idesc, (it >= 3 || k) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(tfull);
        }
    } else {
        const int quad = warp & 3;
        const int co = co0 + quad * 32 + lane;
        const int taps = a.kh * a.kw;
        float *orow = a.dw + ((size_t)co * taps + tap) * a.cin + ci0;
        float inv = 1.f, corr = 1.f;
        if (f16) {
            inv = (1.f / f16_scale_from_bound(*a.xb)) * (1.f / f16_scale_from_bound(*a.yb));
            corr = 1.f / 2048.f;
        }
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < a.bn; c0 += 32) {
            uint32_t v[32], u[32];
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
            tmem_ld32(trow, v);
            for (int m = 1; m < nmain; ++m) {
                tmem_ld32(trow + m * a.bn, u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
            tmem_ld32(trow + 3 * a.bn, u);
            const int ncol = min(32, a.bn - c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < ncol)
                    atomicAdd(reinterpret_cast<float4 *>(orow + c0 + j),
                              make_float4(fmaf(__uint_as_float(u[j]), corr, __uint_as_float(v[j])) * inv,
                                          fmaf(__uint_as_float(u[j + 1]), corr, __uint_as_float(v[j + 1])) * inv,
                                          fmaf(__uint_as_float(u[j + 2]), corr, __uint_as_float(v[j + 2])) * inv,
                                          fmaf(__uint_as_float(u[j + 3]), corr, __uint_as_float(v[j + 3])) * inv));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
    }
}

int conv_tc_wgrad(const ConvArgs &a, cudaStream_t st) {
    // a.x: padded input (x_hi / x_lo); a.y: dy geometry (w_hi / w_lo carry dy); a.out = dw [cout][kh][kw][cin]
    // f16 mode: x_h2 / w_h2 are the packed planes of x / dy, x_bound / w_bound their bounds
    const bool f16 = a.x_h2 != nullptr;
    if (f16 ? !(a.w_h2 && a.x_bound && a.w_bound) : !(a.x_lo && a.w_lo)) return 0;
    if (a.sh != 1 || a.sw != 1) return 0;
    if (a.kh != 2 * a.ph + 1 || a.kw != 2 * a.pw + 1) return 0;
    const int krows = f16 ? 64 : 32, bc = f16 ? 64 : 32;
    if (a.cout % TC_BM != 0 || a.cin % bc != 0) return 0;
    // x and dy must share one padded grid, with pads covering the kernel reach
    if (a.x.n != a.y.n || a.x.h != a.y.h || a.x.w != a.y.w || a.x.ph != a.y.ph || a.x.pw != a.y.pw) return 0;
    if (a.x.ph < a.ph || a.x.pw < a.pw) return 0;
    int bn = 0;
    for (int c : {128, 64, 32})
        if (a.cin % c == 0 && c % bc == 0) { bn = c; break; }
    if (!bn) return 0;
    const long long rows = (long long)a.x.n * a.x.hp * a.x.wp;
    if (rows >= (1LL << 31) - 4096) return 0;
    if ((((uintptr_t)a.x_hi | (uintptr_t)a.x_lo | (uintptr_t)a.w_hi | (uintptr_t)a.w_lo | (uintptr_t)a.out |
          (uintptr_t)a.x_h2 | (uintptr_t)a.w_h2) & 15) != 0) return 0;

    const int taps = a.kh * a.kw;
    const int tiles = taps * (a.cin / bn) * (a.cout / TC_BM);
    // K splits: one CTA per SM (192 KB of smem), so the grid should fill whole waves of 148 CTAs -- a grid of
    // 450 CTAs (3.04 waves) ran at 76 % of a 444-CTA one.  Pick the split count whose total is closest below a
    // multiple of 148 among 2..4 waves, keeping at least 64 K chunks per CTA.
    const long long max_splits = (rows + 64 * krows - 1) / (64 * krows);
    int splits = 1;
    double best_fill = 0.0;
    for (int waves = 2; waves <= 4; ++waves) {
        int s = waves * 148 / tiles;
        if (s > max_splits) s = (int)max_splits;
        if (s < 1) s = 1;
        const long long ctas = (long long)s * tiles;
        const double fill = (double)ctas / (double)(((ctas + 147) / 148) * 148);
        if (fill > best_fill + 1e-9) { best_fill = fill; splits = s; }
    }
    long long rps = (rows + splits - 1) / splits;
    rps = (rps + krows - 1) / krows * krows;
    if ((rows + rps - 1) / rps < splits) splits = (int)((rows + rps - 1) / rps);

    TcWgradArgs t;
    t.kh = a.kh; t.kw = a.kw; t.ph = a.ph; t.pw = a.pw; t.wp = a.x.wp;
    t.cin = a.cin; t.cout = a.cout; t.bn = bn;
    t.rows = rows; t.rows_per_split = rps; t.dw = a.out;
    t.f16 = f16 ? 1 : 0; t.xb = a.x_bound; t.yb = a.w_bound;
    const int stage_bytes = 2 * TC_BM * TC_BK * 4 + 2 * bn * TC_BK * 4;
    int stages = TC_SMEM_LIMIT / stage_bytes;
    if (stages > 6) stages = 6;
    t.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;

    CUtensorMap mdh, mdl, mxh, mxl;
    int rc;
    const CUtensorMapSwizzle sw = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if ((rc = make_map_pair(&mdh, &mdl, a.w_hi, a.w_lo, a.w_h2, rows, a.cout, krows, sw))) return rc;
    if ((rc = make_map_pair(&mxh, &mxl, a.x_hi, a.x_lo, a.x_h2, rows, a.cin, krows, sw))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        DLIO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
        attr_set = true;
    }
    ProfScope prof(DLIO_PROF_CONV_WGRAD_TC, st);
    DLIO_CUDA(cudaMemsetAsync(a.out, 0, (size_t)a.cout * taps * a.cin * sizeof(float), st));
    dim3 grid((unsigned)splits, (unsigned)(taps * (a.cin / bn)), (unsigned)(a.cout / TC_BM));
    wgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(mdh, mdl, mxh, mxl, t);
    DLIO_LAUNCH_CHECK();
    return 1;
}

}  // namespace dlio
