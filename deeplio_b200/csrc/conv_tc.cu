// tcgen05 implicit-GEMM convolutions for sm_100a: stride-1 "same" convolutions computed with error-compensated
// split operands -- three MMAs per fp32-accurate product (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM), which
// keeps the result within ~1e-6 of an fp32 convolution (the 1e-4 pose gate rules out single-pass TF32 / BF16).
// Two operand formats share the kernels (template F16): packed fp16 hi|lo planes with per-tensor power-of-two scales
// ("3xF16", kind::f16, Cin % 64 == 0: the main path) and fp32 + TF32-residual planes ("3xTF32", kind::tf32,
// Cin % 32 == 0: the first layer through its space-to-depth view).  Strided layers arrive here as stride-1 problems
// through views of the same memory (conv_s2d.cu), with an optional row decimation in the epilogue (hdec).
//
// Formulation ("shifted GEMM over the padded grid").  The input is a padded NHWC tensor whose zero pads are
// real memory, viewed as a matrix X[R = n*hp*wp, Cin].  For EVERY padded position q the kernel computes
//     Y[q, co] = sum_{tap=(dy,dx)} sum_ci X[q + (dy-ph)*wp + (dx-pw), ci] * W[co, tap, ci]
// which is the convolution at interior positions; rows that fall on pads are discarded in the epilogue
// (2-15 % extra MMA work, no im2col, no gather).  Each (tap, 128-byte channel chunk) is one pipeline stage: a plain
// 2-D TMA box [128 rows x 128 bytes] of X at row offset q0 + shift (out-of-range rows are zero-filled by TMA
// and only feed discarded rows) and a [BN x 128 bytes] box of W, both 128B-swizzled, hi and lo planes.
//
// conv_tc_kernel (forward and dgrad): persistent, one CTA per SM, 10 warps: warp 0 TMA producer, warp 1 tcgen05.mma
// issuer (one elected lane of the converged warp) + TMEM owner, warps 2-9 epilogue (TMEM -> registers -> scale /
// bias / ReLU -> global, per-channel BN statistics); the epilogue of tile t overlaps the main loop of tile t + 1.
// wgrad_tc_kernel: one 128 x BN tile of dw per CTA, split-K over pixel ranges, both operands MN-major.
// conv_tc2_kernel / wgrad_tc2_kernel: the same pipelines on CTA PAIRS (tcgen05 cta_group::2, M = 256): each CTA stages
// its own 128 A rows and half of the B tile, the leader CTA issues, commits are multicast to both CTAs' barriers --
// halves the B-side shared-memory traffic that bounds the single-CTA kernels; used whenever the output tile is 128
// channels wide (forward / dgrad) or Cin % 128 == 0 and Cout % 256 == 0 (wgrad).
// "fold" mode (first layer, conv_s2d.cu): one plane of x whose 128-byte rows hold hi AND lo halves, two products.
//
// Accumulation accuracy.  The tensor core adds each MMA into the fp32 accumulator with truncation, so the
// error of ONE accumulator grows linearly with the number of MMA steps (measured: 8e-9 * K relative, 3.7e-5 at
// K = 4608 -- too much for gradient parity).  The forward kernel therefore drains its main accumulator into
// registers (round-to-nearest adds) every 8 K stages and keeps the small lo*hi + hi*lo corrections in a separate
// accumulator (see the kernel); wgrad spreads the hi*hi products round-robin over three accumulators.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv.cuh"

namespace dlio {

constexpr int TC_BM = 128;       // output rows (padded pixels) per CTA
constexpr int TC_BK = 32;        // fp32 elements per K chunk = one 128-byte swizzle row (64 halves in f16 mode)
constexpr int TC_BK16 = 64;
constexpr int TC_THREADS = 192;          // wgrad kernel: producer, MMA, 4 epilogue warps
constexpr int TC_FWD_THREADS = 320;      // forward / dgrad kernel: producer, MMA, 8 epilogue warps
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_SMEM_LIMIT = 221 * 1024;   // pipeline stages + epilogue scratch (227 KB per CTA minus alignment slack)
constexpr int TC_EPI_SMEM = TC_EPI_WARPS * 2 * 64 * 8;   // per-warp fp64 column sums

struct TcArgs {
    Geo x;             // padded input geometry (pads are memory)
    Geo o;             // output geometry
    int kh, kw, ph, pw;
    int cin, cout, bn;  // bn: output channels per CTA (multiple of 16, <= 256)
    int act;
    int stages;
    int seg;           // K stages per main-accumulator segment
    long long mtiles;  // 128-row tiles of the padded grid
    int ntiles;        // cout / bn
    const float *bias;
    float *out;
    double *stats;
    long long rows;    // R = n * hp * wp
    int f16;           // 0: 3xTF32 on fp32 planes; 1: 3xF16 on packed half planes (K chunk = 64 halves)
    int hdec;          // 2: H-stride-2 convolution computed over the whole grid, only even rows are stored / counted
    const float *xb, *wb;   // f16: device bounds of the two operands (-> power-of-two scales)
    int single;        // A/B switch "bwd_single_pass": hi*hi products only (the correction accumulator is initialised
                       // by one lo*hi MMA per tile and otherwise left alone) -- fp16 / TF32 accuracy at 1/3 of the MMAs
    int accum;         // dgrad: add the result to what `out` holds (a second gradient contribution, no temporary)
    int fold;          // "folded split" (first layer, conv_s2d.cu): every 128-byte row of x holds hi AND lo halves of
                       // 32 channels, the weight planes are [w_hi | 0] and [w_lo | w_hi]: no x_lo plane, two products
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
// One lane of a converged warp (elect.sync): the tcgen05.mma / TMA / commit instructions take their operands from
// uniform registers, and ptxas only keeps them there -- without wrapping every instruction in a per-lane
// "elect, move to uniform registers, execute, loop" sequence of ~10 instructions -- when the warp is converged and
// the issuing lane comes from elect.sync.  (Issuing from `if (lane == 0)` cost ~300 instructions per K stage and
// made the single issuing thread, not the tensor pipe, the bound of the main loop: 190 clocks per MMA.)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// MMA kind and A-collector usage are compile-time; F16 selects kind::f16 (fp16 operands, K = 16 per
// instruction) over kind::tf32 (K = 8) -- both consume 32 bytes of every operand row per instruction.
#define DLIO_UMMA_ASM(KIND, COLL)                                                                         \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                      \
                 "tcgen05.mma.cta_group::1.kind::" KIND COLL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), \
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)                                             \
                 : "memory")
// coll: 0 plain, 1 collector::a::fill, 2 collector::a::lastuse
template <bool F16, int COLL>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (F16) {
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

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// v[j] = value of column j in this lane's row; returns the sum over the warp's 32 rows of column `lane`
// (16 + 8 + 4 + 2 + 1 shuffles; v is destroyed)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
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

// Persistent kernel: grid = min(tiles, SMs) CTAs, tile t -> (M tile t / ntiles, N tile t % ntiles), so the CTAs
// that run at the same time share their X rows in L2.  The four TMEM regions of bn columns are two MAIN
// accumulators used in ping-pong and two CORRECTION accumulators (one per tile parity):
//   * the MMA warp sends the hi*hi products of SEG consecutive K stages into one main accumulator, commits it
//     (mfull) and switches to the other one; the lo*hi + hi*lo products of the whole tile go to corr[tile & 1];
//   * the epilogue warps drain each committed main accumulator into REGISTERS (round-to-nearest fp32 adds) while
//     the tensor pipe works on the next segment, so a TMEM accumulator never sees more than 4*SEG truncating MMA
//     steps whatever K is, and at the end of a tile only the correction accumulator is left to read;
//   * bias / ReLU / store / BN statistics of tile t overlap the main loop of tile t + 1.
// Barriers: full/empty per smem stage; mfull/mempty per main accumulator; cfull/cempty per correction accumulator.
template <bool F16>
__global__ void __launch_bounds__(TC_FWD_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
               const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * 8 + 8];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = TC_BM * TC_BK * 4;             // 16 KB per plane
    const uint32_t b_bytes = (uint32_t)a.bn * TC_BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const int S = a.stages;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
    const uint32_t mfull0 = smem_u32(&bars[16]), mempty0 = smem_u32(&bars[18]);
    const uint32_t cfull0 = smem_u32(&bars[20]), cempty0 = smem_u32(&bars[22]);

    constexpr bool f16 = F16;
    constexpr int bk = F16 ? TC_BK16 : TC_BK;               // elements per 128-byte K chunk
    const int cchunks = a.cin / bk;
    const int iters = a.kh * a.kw * cchunks;                // K stages per tile
    const int SEG = a.seg;
    const long long tiles = a.mtiles * a.ntiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(mfull0 + 8 * i, 1);
            mbar_init(mempty0 + 8 * i, TC_EPI_WARPS);
            mbar_init(cfull0 + 8 * i, 1);
            mbar_init(cempty0 + 8 * i, TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    // TMEM columns: main accumulators at 0 and bn, correction accumulators at 2 bn and 3 bn

    if (warp == 0) {
        // ===== TMA producer (whole warp converged, one elected lane issues) =====
        uint32_t s = 0, ph = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long long q0 = (tile / a.ntiles) * TC_BM;
            const int n0 = (int)(tile % a.ntiles) * a.bn;
            int dy = 0, dx = 0, cc = 0, wcol = 0;            // tap (dy, dx), channel chunk, weight column tap*cin + cc*bk
            for (int it = 0; it < iters; ++it) {
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                if (elect_one()) {
                    const int row = (int)(q0 + (long long)(dy - a.ph) * a.x.wp + (dx - a.pw));
                    const uint32_t sa = smem_u32(smem) + s * stage_bytes;
                    const uint32_t fb = full0 + 8 * s;
                    mbar_expect_tx(fb, a.fold ? stage_bytes - a_bytes : stage_bytes);
                    tma_load_2d(sa, &tm_xhi, fb, cc * bk, row);
                    if (!a.fold) tma_load_2d(sa + a_bytes, &tm_xlo, fb, cc * bk, row);
                    tma_load_2d(sa + 2 * a_bytes, &tm_whi, fb, wcol, n0);
                    tma_load_2d(sa + 2 * a_bytes + b_bytes, &tm_wlo, fb, wcol, n0);
                }
                __syncwarp();
                wcol += bk;
                if (++cc == cchunks) {
                    cc = 0;
                    if (++dx == a.kw) { dx = 0; ++dy; }
                }
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp converged, one elected lane issues) =====
        const uint32_t idesc = make_idesc(a.bn, f16);
        const uint32_t sa0 = smem_u32(smem);
        const uint64_t dA_hi = make_kmajor_desc(sa0), dA_lo = make_kmajor_desc(sa0 + a_bytes);
        const uint64_t dB_hi = make_kmajor_desc(sa0 + 2 * a_bytes), dB_lo = make_kmajor_desc(sa0 + 2 * a_bytes + b_bytes);
        const uint32_t stage16 = stage_bytes >> 4;               // descriptor start addresses count 16-byte units
        uint32_t s = 0, ph = 0, seg = 0, tcount = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
            const uint32_t p = tcount & 1u;
            mbar_wait(cempty0 + 8 * p, ((tcount >> 1) & 1u) ^ 1u);     // epilogue has read corr[p] of tile - 2
            const uint32_t t_corr = tmem_base + (2 + p) * a.bn;
            uint32_t t_main = 0;
            int si = 0;
            for (int it = 0; it < iters; ++it) {
                if (si == 0) {
                    mbar_wait(mempty0 + 8 * (seg & 1u), ((seg >> 1) & 1u) ^ 1u);  // epilogue has drained main[seg & 1]
                    t_main = tmem_base + (seg & 1u) * a.bn;
                }
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool seg_end = si == SEG - 1 || it == iters - 1;
                if (elect_one()) {
                    const uint64_t so = (uint64_t)(s * stage16);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                // 128-byte rows / 32 bytes per MMA K step
                        const uint64_t ko = so + (uint64_t)(k * 2);
                        if (!a.fold && (!a.single || !(it | k)))
                            umma<F16, 0>(t_corr, dA_lo + ko, dB_hi + ko, idesc, (it | k) ? 1u : 0u);
                        if (!a.single) umma<F16, 1>(t_corr, dA_hi + ko, dB_lo + ko, idesc, (a.fold && !(it | k)) ? 0u : 1u);
                        umma<F16, 2>(t_main, dA_hi + ko, dB_hi + ko, idesc, (si | k) ? 1u : 0u);
                    }
                    umma_commit(empty0 + 8 * s);                 // frees the smem stage when these MMAs retire
                    if (seg_end) umma_commit(mfull0 + 8 * (seg & 1u));   // this segment's main accumulator is complete
                    if (it == iters - 1) umma_commit(cfull0 + 8 * p);    // and so is the tile's correction accumulator
                }
                __syncwarp();
                if (seg_end) { ++seg; si = 0; } else ++si;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: 8 warps; warp w reads TMEM lane quadrant w % 4 (a hardware rule) and one half of the
        // tile's columns, so every scheduler has two epilogue warps to hide TMEM / shared-memory latency with
        const int quad = warp & 3, ew = warp - 2, half = ew >> 2;
        const int cw = a.bn >= 64 ? a.bn / 2 : a.bn;          // columns per warp
        const bool active = half == 0 || a.bn >= 64;
        const int cb = half * cw;                              // first column of this warp inside the tile
        double *red = reinterpret_cast<double *>(smem + (size_t)S * stage_bytes) + (size_t)ew * 2 * 64;   // [2][64]
        for (int i = lane; i < 2 * 64; i += 32) red[i] = 0.0;
        // f16: undo the two power-of-two operand scales; the correction accumulator carries another 2^-11
        float inv = 1.f, corr = 1.f;
        if (f16) {
            inv = (1.f / f16_scale_from_bound(*a.xb)) * (1.f / f16_scale_from_bound(*a.wb));
            corr = 1.f / 2048.f;
        }
        const int nseg = (iters + SEG - 1) / SEG;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        uint32_t seg = 0, tcount = 0;
        int red_n0 = -1;
        auto flush_stats = [&](int n0) {
            __syncwarp();
            for (int c = lane; c < cw; c += 32) {
                atomicAdd(a.stats + n0 + cb + c, red[c]);
                atomicAdd(a.stats + a.cout + n0 + cb + c, red[64 + c]);
                red[c] = 0.0;
                red[64 + c] = 0.0;
            }
            __syncwarp();
        };
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
            const long long q = (tile / a.ntiles) * TC_BM + quad * 32 + lane;   // row of the tile == TMEM lane
            const int n0 = (int)(tile % a.ntiles) * a.bn;
            bool valid = q < a.rows;
            size_t obase = 0;
            if (valid) {
                int xx = (int)(q % a.x.wp);
                long long t = q / a.x.wp;
                int yy = (int)(t % a.x.hp);
                int n = (int)(t / a.x.hp);
                int h = yy - a.x.ph, w = xx - a.x.pw;
                valid = h >= 0 && h < a.x.h && w >= 0 && w < a.x.w;
                if (a.hdec == 2) {
                    valid = valid && !(h & 1);
                    h >>= 1;
                }
                if (valid) obase = a.o.off(n, h, w) + n0 + cb;
            }
            if (a.stats && active && red_n0 != n0) {
                if (red_n0 >= 0) flush_stats(red_n0);
                red_n0 = n0;
            }
            // --- drain the main-accumulator segments into registers as the MMA warp completes them
            float acc[2][32];
            for (int sg = 0; sg < nseg; ++sg, ++seg) {
                const uint32_t b = seg & 1u;
                mbar_wait(mfull0 + 8 * b, (seg >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (active) {
                    const uint32_t t_main = tmem_base + lane_off + b * a.bn + cb;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c * 32 < cw) {
                            uint32_t v[32];
                            tmem_ld32_nowait(t_main + c * 32, v);
                            tmem_ld_wait();
                            if (sg == 0) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) acc[c][j] = __uint_as_float(v[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) acc[c][j] += __uint_as_float(v[j]);
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(mempty0 + 8 * b);
            }
            // --- correction accumulator, then the tile's epilogue proper (overlaps the next tile's main loop)
            const uint32_t p = tcount & 1u;
            mbar_wait(cfull0 + 8 * p, (tcount >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float rinv = valid ? inv : 0.f;      // rows on pads contribute exact zeros to the statistics
            if (active) {
                const uint32_t t_corr = tmem_base + lane_off + (2 + p) * a.bn + cb;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c * 32 < cw) {
                        uint32_t u[32];
                        tmem_ld32_nowait(t_corr + c * 32, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c][j] = fmaf(__uint_as_float(u[j]), corr, acc[c][j]) * rinv;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(cempty0 + 8 * p);
            if (!active) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int c0 = c * 32;
                if (c0 < cw) {
                    const int ncol = min(32, cw - c0);
                    if (a.bias) {
                        const float bmask = valid ? 1.f : 0.f;
                        const float *bp = a.bias + n0 + cb + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (j < ncol) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bp + j));
                                acc[c][j] = fmaf(b4.x, bmask, acc[c][j]);
                                acc[c][j + 1] = fmaf(b4.y, bmask, acc[c][j + 1]);
                                acc[c][j + 2] = fmaf(b4.z, bmask, acc[c][j + 2]);
                                acc[c][j + 3] = fmaf(b4.w, bmask, acc[c][j + 3]);
                            }
                        }
                    }
                    if (a.act == DLIO_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c][j] = fmaxf(acc[c][j], 0.f);
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (j < ncol)
                            {
                                // accumulate: one fire-and-forget 128-bit reduction per element group (a load of
                                // the old value here would stall the epilogue on an uncoalesced round trip per row)
                                const float4 o4 = make_float4(acc[c][j], acc[c][j + 1], acc[c][j + 2], acc[c][j + 3]);
                                if (a.accum) atomicAdd(reinterpret_cast<float4 *>(a.out + obase + c0 + j), o4);
                                else st4(a.out + obase + c0 + j, o4);
                            }
                    }
                    if (a.stats) {
                        // per-column sums over the warp's 32 rows: butterfly transpose-reduce, lane l ends with column l
                        float sq[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) sq[j] = acc[c][j] * acc[c][j];
                        const float s1 = warp_colsum32(acc[c], lane), s2 = warp_colsum32(sq, lane);
                        if (lane < ncol) {          // this lane owns column c0 + lane of the warp's private sums
                            red[c0 + lane] += (double)s1;
                            red[64 + c0 + lane] += (double)s2;
                        }
                    }
                }
            }
        }
        if (a.stats && active && red_n0 >= 0) flush_stats(red_n0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ==================================================================== CTA-pair variant (cta_group::2)
// The single-CTA kernel above is bound by shared-memory bandwidth, not by the tensor pipe: per 32-byte K step the
// three MMAs read 20 KB of operands (A_lo, B_hi | A_hi, B_lo | B_hi; A_hi reused through the collector) and TMA
// writes 16 KB, 36 KB per 192 tensor-pipe clocks = 187 B/clk against the SM's 128 B/clk -- the measured 67 - 75 %
// tensor-pipe activity IS that ratio.  Here two CTAs on the two SMs of a TPC run ONE tcgen05.mma.cta_group::2 with
// M = 256 (128 rows of A per CTA) and N = bn: each CTA stages only HALF of the B tile (bn / 2 weight rows) and the
// tensor cores of both SMs read both halves, so per CTA and K step the MMAs read 14 KB and TMA writes 12 KB:
// 135 B/clk.  Structure, accumulator scheme and epilogue are those of conv_tc_kernel; what changes is the barrier
// protocol across the pair:
//   * rank 0 (leader) issues every MMA; both CTAs run a TMA producer for their own A rows and B half, and both
//     signal the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2, barrier address mapped to rank 0),
//     on which the leader posts the byte count of both CTAs;
//   * tcgen05.commit.cta_group::2 ... multicast::cluster arrives on the empty / mfull / cfull barriers of BOTH CTAs;
//   * the epilogue warps of both CTAs drain their own TMEM (rows r * 128 ..) and arrive on the leader's mempty /
//     cempty barriers (count 16), remotely for rank 1.
__device__ __forceinline__ uint32_t cluster_ctarank_() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
#define DLIO_UMMA2_ASM(COLL)                                                                                   \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                            \
                 "tcgen05.mma.cta_group::2.kind::f16" COLL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),      \
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)                                                  \
                 : "memory")
template <int COLL>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (COLL == 0) DLIO_UMMA2_ASM("");
    else if (COLL == 1) DLIO_UMMA2_ASM(".collector::a::fill");
    else DLIO_UMMA2_ASM(".collector::a::lastuse");
}
// arrives (when the MMAs issued so far retire) on the barrier at this shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_FWD_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
                const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * 8 + 8];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank_();
    const bool leader = rank == 0;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int bk = TC_BK16;                               // halves per 128-byte K chunk
    const int bnh = a.bn / 2;                                 // weight rows staged by this CTA
    const uint32_t a_bytes = TC_BM * 128;                     // 16 KB per plane
    const uint32_t b_bytes = (uint32_t)bnh * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const int S = a.stages;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
    const uint32_t mfull0 = smem_u32(&bars[16]), mempty0 = smem_u32(&bars[18]);
    const uint32_t cfull0 = smem_u32(&bars[20]), cempty0 = smem_u32(&bars[22]);
    const int cchunks = a.cin / bk;
    const int iters = a.kh * a.kw * cchunks;
    const int SEG = a.seg;
    const long long tiles = a.mtiles * a.ntiles;              // mtiles counts 256-row tiles here
    const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(mfull0 + 8 * i, 1);
            mbar_init(mempty0 + 8 * i, 2 * TC_EPI_WARPS);
            mbar_init(cfull0 + 8 * i, 1);
            mbar_init(cempty0 + 8 * i, 2 * TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();            // the peer's barriers are initialised before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer of this CTA's A rows and B half; completion on the leader's full barrier =====
        uint32_t s = 0, ph = 0;
        for (long long tile = pair; tile < tiles; tile += npairs) {
            const long long q0 = (tile / a.ntiles) * (2 * TC_BM) + (long long)rank * TC_BM;
            const int n0 = (int)(tile % a.ntiles) * a.bn + (int)rank * bnh;
            int dy = 0, dx = 0, cc = 0, wcol = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                if (elect_one()) {
                    const int row = (int)(q0 + (long long)(dy - a.ph) * a.x.wp + (dx - a.pw));
                    const uint32_t sa = smem_u32(smem) + s * stage_bytes;
                    const uint32_t fb = map_to_rank(full0 + 8 * s, 0);
                    if (leader) mbar_expect_tx(full0 + 8 * s, 2 * (a.fold ? stage_bytes - a_bytes : stage_bytes));
                    tma_load_2d_pair(sa, &tm_xhi, fb, cc * bk, row);
                    if (!a.fold) tma_load_2d_pair(sa + a_bytes, &tm_xlo, fb, cc * bk, row);
                    tma_load_2d_pair(sa + 2 * a_bytes, &tm_whi, fb, wcol, n0);
                    tma_load_2d_pair(sa + 2 * a_bytes + b_bytes, &tm_wlo, fb, wcol, n0);
                }
                __syncwarp();
                wcol += bk;
                if (++cc == cchunks) {
                    cc = 0;
                    if (++dx == a.kw) { dx = 0; ++dy; }
                }
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===== MMA issuer of the pair =====
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
            const uint32_t sa0 = smem_u32(smem);
            const uint64_t dA_hi = make_kmajor_desc(sa0), dA_lo = make_kmajor_desc(sa0 + a_bytes);
            const uint64_t dB_hi = make_kmajor_desc(sa0 + 2 * a_bytes), dB_lo = make_kmajor_desc(sa0 + 2 * a_bytes + b_bytes);
            const uint32_t stage16 = stage_bytes >> 4;
            uint32_t s = 0, ph = 0, seg = 0, tcount = 0;
            for (long long tile = pair; tile < tiles; tile += npairs, ++tcount) {
                const uint32_t p = tcount & 1u;
                mbar_wait(cempty0 + 8 * p, ((tcount >> 1) & 1u) ^ 1u);
                const uint32_t t_corr = tmem_base + (2 + p) * a.bn;
                uint32_t t_main = 0;
                int si = 0;
                for (int it = 0; it < iters; ++it) {
                    if (si == 0) {
                        mbar_wait(mempty0 + 8 * (seg & 1u), ((seg >> 1) & 1u) ^ 1u);
                        t_main = tmem_base + (seg & 1u) * a.bn;
                    }
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const bool seg_end = si == SEG - 1 || it == iters - 1;
                    if (elect_one()) {
                        const uint64_t so = (uint64_t)(s * stage16);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ko = so + (uint64_t)(k * 2);
                            if (!a.fold && (!a.single || !(it | k)))
                                umma2<0>(t_corr, dA_lo + ko, dB_hi + ko, idesc, (it | k) ? 1u : 0u);
                            if (!a.single) umma2<1>(t_corr, dA_hi + ko, dB_lo + ko, idesc, (a.fold && !(it | k)) ? 0u : 1u);
                            umma2<2>(t_main, dA_hi + ko, dB_hi + ko, idesc, (si | k) ? 1u : 0u);
                        }
                        umma2_commit_both(empty0 + 8 * s);
                        if (seg_end) umma2_commit_both(mfull0 + 8 * (seg & 1u));
                        if (it == iters - 1) umma2_commit_both(cfull0 + 8 * p);
                    }
                    __syncwarp();
                    if (seg_end) { ++seg; si = 0; } else ++si;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ===== epilogue: as in conv_tc_kernel, on this CTA's 128 rows of the 256-row tile =====
        const int quad = warp & 3, ew = warp - 2, half = ew >> 2;
        const int cw = a.bn >= 64 ? a.bn / 2 : a.bn;
        const bool active = half == 0 || a.bn >= 64;
        const int cb = half * cw;
        double *red = reinterpret_cast<double *>(smem + (size_t)S * stage_bytes) + (size_t)ew * 2 * 64;
        for (int i = lane; i < 2 * 64; i += 32) red[i] = 0.0;
        const float inv = (1.f / f16_scale_from_bound(*a.xb)) * (1.f / f16_scale_from_bound(*a.wb));
        const float corr = 1.f / 2048.f;
        const int nseg = (iters + SEG - 1) / SEG;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const uint32_t mempty_l = map_to_rank(mempty0, 0), cempty_l = map_to_rank(cempty0, 0);
        uint32_t seg = 0, tcount = 0;
        int red_n0 = -1;
        auto flush_stats = [&](int n0) {
            __syncwarp();
            for (int c = lane; c < cw; c += 32) {
                atomicAdd(a.stats + n0 + cb + c, red[c]);
                atomicAdd(a.stats + a.cout + n0 + cb + c, red[64 + c]);
                red[c] = 0.0;
                red[64 + c] = 0.0;
            }
            __syncwarp();
        };
        for (long long tile = pair; tile < tiles; tile += npairs, ++tcount) {
            const long long q = (tile / a.ntiles) * (2 * TC_BM) + (long long)rank * TC_BM + quad * 32 + lane;
            const int n0 = (int)(tile % a.ntiles) * a.bn;
            bool valid = q < a.rows;
            size_t obase = 0;
            if (valid) {
                int xx = (int)(q % a.x.wp);
                long long t = q / a.x.wp;
                int yy = (int)(t % a.x.hp);
                int n = (int)(t / a.x.hp);
                int h = yy - a.x.ph, w = xx - a.x.pw;
                valid = h >= 0 && h < a.x.h && w >= 0 && w < a.x.w;
                if (a.hdec == 2) {
                    valid = valid && !(h & 1);
                    h >>= 1;
                }
                if (valid) obase = a.o.off(n, h, w) + n0 + cb;
            }
            if (a.stats && active && red_n0 != n0) {
                if (red_n0 >= 0) flush_stats(red_n0);
                red_n0 = n0;
            }
            float acc[2][32];
            for (int sg = 0; sg < nseg; ++sg, ++seg) {
                const uint32_t b = seg & 1u;
                mbar_wait(mfull0 + 8 * b, (seg >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (active) {
                    const uint32_t t_main = tmem_base + lane_off + b * a.bn + cb;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c * 32 < cw) {
                            uint32_t v[32];
                            tmem_ld32_nowait(t_main + c * 32, v);
                            tmem_ld_wait();
                            if (sg == 0) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) acc[c][j] = __uint_as_float(v[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) acc[c][j] += __uint_as_float(v[j]);
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mempty_l + 8 * b);
            }
            const uint32_t p = tcount & 1u;
            mbar_wait(cfull0 + 8 * p, (tcount >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float rinv = valid ? inv : 0.f;
            if (active) {
                const uint32_t t_corr = tmem_base + lane_off + (2 + p) * a.bn + cb;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c * 32 < cw) {
                        uint32_t u[32];
                        tmem_ld32_nowait(t_corr + c * 32, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c][j] = fmaf(__uint_as_float(u[j]), corr, acc[c][j]) * rinv;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(cempty_l + 8 * p);
            if (!active) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int c0 = c * 32;
                if (c0 < cw) {
                    const int ncol = min(32, cw - c0);
                    if (a.bias) {
                        const float bmask = valid ? 1.f : 0.f;
                        const float *bp = a.bias + n0 + cb + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (j < ncol) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bp + j));
                                acc[c][j] = fmaf(b4.x, bmask, acc[c][j]);
                                acc[c][j + 1] = fmaf(b4.y, bmask, acc[c][j + 1]);
                                acc[c][j + 2] = fmaf(b4.z, bmask, acc[c][j + 2]);
                                acc[c][j + 3] = fmaf(b4.w, bmask, acc[c][j + 3]);
                            }
                        }
                    }
                    if (a.act == DLIO_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c][j] = fmaxf(acc[c][j], 0.f);
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (j < ncol)
                            {
                                // accumulate: one fire-and-forget 128-bit reduction per element group (a load of
                                // the old value here would stall the epilogue on an uncoalesced round trip per row)
                                const float4 o4 = make_float4(acc[c][j], acc[c][j + 1], acc[c][j + 2], acc[c][j + 3]);
                                if (a.accum) atomicAdd(reinterpret_cast<float4 *>(a.out + obase + c0 + j), o4);
                                else st4(a.out + obase + c0 + j, o4);
                            }
                    }
                    if (a.stats) {
                        float sq[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) sq[j] = acc[c][j] * acc[c][j];
                        const float s1 = warp_colsum32(acc[c], lane), s2 = warp_colsum32(sq, lane);
                        if (lane < ncol) {
                            red[c0 + lane] += (double)s1;
                            red[64 + c0 + lane] += (double)s2;
                        }
                    }
                }
            }
        }
        if (a.stats && active && red_n0 >= 0) flush_stats(red_n0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();            // neither CTA leaves (or frees TMEM) while the other may still signal / read it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

// CTA-pair convolution kernel (conv_tc2_kernel): on unless DLIO_CONV_CG2=0 / dlio_set_option("conv_cg2", 0)
// (measured on the headline step: forward class 4.26 -> 3.94 ms, dgrad 3.61 -> 3.13 ms, outputs bit-identical)
int g_conv_cg2 = -1;
int g_bwd_single = 0;   // option "bwd_single_pass" (A/B measurement only; never the default)
static bool conv_cg2_enabled() {
    if (g_conv_cg2 < 0) {
        const char *e = getenv("DLIO_CONV_CG2");
        g_conv_cg2 = (e && e[0] == '0') ? 0 : 1;
    }
    return g_conv_cg2 != 0;
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
    if (a.hdec != 1 && a.hdec != 2) return 0;
    if (a.cin % (f16 ? TC_BK16 : TC_BK) != 0 || a.cout % 16 != 0) return 0;
    if (a.x.ph < a.ph || a.x.pw < a.pw) return 0;
    if (a.kh != 2 * a.ph + 1 || a.kw != 2 * a.pw + 1) return 0;   // "same" convolution: output extent == input extent
    const int bn = pick_bn(a.cout);
    if (!bn) return 0;
    const long long rows = (long long)a.x.n * a.x.hp * a.x.wp;
    if (rows >= (1LL << 31) - 4096) return 0;
    if ((((uintptr_t)a.x_hi | (uintptr_t)a.x_lo | (uintptr_t)a.w_hi | (uintptr_t)a.w_lo | (uintptr_t)a.out |
          (uintptr_t)a.x_h2 | (uintptr_t)a.w_h2) & 15) != 0) return 0;

    if (a.fold && !(f16 && a.cin == TC_BK16 && a.hdec == 1)) return 0;
    TcArgs t;
    t.f16 = f16 ? 1 : 0; t.xb = a.x_bound; t.wb = a.w_bound;
    t.hdec = a.hdec;
    t.fold = a.fold;
    t.accum = a.accum;
    t.single = (prof_kind == DLIO_PROF_CONV_DGRAD_TC && g_bwd_single) ? 1 : 0;
    t.x = a.x; t.o = a.o;
    t.kh = a.kh; t.kw = a.kw; t.ph = a.ph; t.pw = a.pw;
    t.cin = a.cin; t.cout = a.cout; t.bn = bn; t.act = a.act;
    t.bias = a.bias; t.out = a.out; t.stats = a.stats; t.rows = rows;
    const int stage_bytes = 2 * TC_BM * TC_BK * 4 + 2 * bn * TC_BK * 4;
    int stages = (TC_SMEM_LIMIT - TC_EPI_SMEM) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) return 0;
    t.stages = stages;
    t.seg = 8;   // 32 truncating MMA steps per TMEM accumulator, whatever K is
    const size_t smem = (size_t)stages * stage_bytes + TC_EPI_SMEM + 1024;

    CUtensorMap mxh, mxl, mwh, mwl;
    const long long K = (long long)a.kh * a.kw * a.cin;
    int rc;
    if (a.fold) {       // one plane, rows of 64 halves
        if ((rc = make_map(&mxh, a.x_h2, true, rows, a.cin, a.cin, TC_BM))) return rc;
        mxl = mxh;
    } else if ((rc = make_map_pair(&mxh, &mxl, a.x_hi, a.x_lo, a.x_h2, rows, a.cin, TC_BM))) {
        return rc;
    }
    if ((rc = make_map_pair(&mwh, &mwl, a.w_hi, a.w_lo, a.w_h2, a.cout, K, bn))) return rc;
    t.mtiles = (rows + TC_BM - 1) / TC_BM;
    t.ntiles = a.cout / bn;
    // (Sharing the weight tile of two independent CTAs through TMA multicast measured 6 % slower: the main loop is
    // bound by shared-memory bandwidth -- TMA fills plus the SS operand reads of the MMAs -- not by L2 -> SM traffic.
    // What does help is cta_group::2 below: ONE MMA over a CTA pair, each CTA holding half of the weight tile.)

    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        DLIO_CUDA(cudaGetDevice(&dev));
        DLIO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        DLIO_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
        DLIO_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
    }
    ProfScope prof(prof_kind, st);
    if ((a.o.ph > 0 || a.o.pw > 0) && !a.accum) DLIO_CUDA(cudaMemsetAsync(a.out, 0, a.o.numel() * sizeof(float), st));
    // 64-channel output tiles (ResNet layer1, conv2's dgrad) on CTA pairs too: the B side of their shared-memory
    // traffic halves as well (A dominates there: 4 KB + 1 KB per MMA and CTA instead of 4 + 2); DLIO_CG2_N64=0 keeps
    // them on the single-CTA kernel
    static const bool pairs_n64 = [] { const char *e = getenv("DLIO_CG2_N64"); return !(e && e[0] == '0'); }();
    if (f16 && (bn == 128 || (bn == 64 && pairs_n64)) && conv_cg2_enabled()) {
        // CTA pairs (conv_tc2_kernel): 256-row tiles, each CTA stages half of the weight tile
        CUtensorMap mwh2, mwl2;
        if ((rc = make_map_pair(&mwh2, &mwl2, a.w_hi, a.w_lo, a.w_h2, a.cout, K, bn / 2))) return rc;
        const int stage2 = 2 * TC_BM * 128 + 2 * (bn / 2) * 128;
        int stages2 = (TC_SMEM_LIMIT - TC_EPI_SMEM) / stage2;
        if (stages2 > 6) stages2 = 6;
        t.stages = stages2;
        t.mtiles = (rows + 2 * TC_BM - 1) / (2 * TC_BM);
        static bool attr2 = false;
        if (!attr2) {
            DLIO_CUDA(cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
            attr2 = true;
        }
        const long long tiles2 = t.mtiles * t.ntiles;
        long long pairs = tiles2 < n_sm / 2 ? tiles2 : n_sm / 2;
        if (pairs > t.ntiles) pairs -= pairs % t.ntiles;
        const size_t smem2 = (size_t)stages2 * stage2 + TC_EPI_SMEM + 1024;
        conv_tc2_kernel<<<(unsigned)(2 * pairs), TC_FWD_THREADS, smem2, st>>>(mxh, mxl, mwh2, mwl2, t);
        DLIO_LAUNCH_CHECK();
        return 1;
    }
    const long long tiles = t.mtiles * t.ntiles;
    // one persistent CTA per SM; keep the grid a multiple of ntiles so that a CTA stays on one N tile (its BN
    // statistics then accumulate in shared memory over all its tiles)
    long long grid = tiles < n_sm ? tiles : n_sm;
    if (grid > t.ntiles) grid -= grid % t.ntiles;
    if (f16) conv_tc_kernel<true><<<(unsigned)grid, TC_FWD_THREADS, smem, st>>>(mxh, mxl, mwh, mwl, t);
    else conv_tc_kernel<false><<<(unsigned)grid, TC_FWD_THREADS, smem, st>>>(mxh, mxl, mwh, mwl, t);
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
// grid = (taps * Cin/BN, K splits, Cout/128); partial tiles are combined with fp32 atomic adds into dw.
struct TcWgradArgs {
    int kh, kw, ph, pw, wp;
    int cin, cout, bn, stages;
    int swap;          // operands swapped (Cin == one TMA box): A = x tiles of 128 / Cin taps, B = dy (N = 128 co)
    long long rows, rows_per_split;
    float *dw;
    int f16;
    const float *xb, *yb;   // f16: device bounds of x and dy
    int single;             // see TcArgs::single
    int fold;               // see TcArgs::fold (swap mode only); fold_cs: halves between the dy boxes of consecutive
    int fold_cs;            // 64-channel groups (dy rows are [64 hi | 64 lo] per output pixel: 128)
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

template <bool F16>
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

    // blockIdx.x = (tap, ci block) is the fastest-varying index, blockIdx.y the K split: the CTAs that are resident
    // together work on the SAME pixel range for different taps, so dy (identical for every tap) and x (shifted by a
    // few rows) are served by L2 -- with the split index fastest, every tap streamed both tensors from DRAM again
    // (ncu: 1.7 GB read for 0.42 GB of operands in Simple-1 conv2, 50 % DRAM utilisation)
    //
    // swap mode (Cin == one TMA box: 32 in TF32 mode -- the first layer through its space-to-depth view -- or 64 in
    // fp16 mode -- Simple-1 conv2, FlowNet conv2).  With dy as the A operand, a 128 x Cin tile costs one MMA per 32
    // bytes of K that reads the whole 4 KB dy slice from shared memory for Cin columns of output: shared-memory bound
    // (~60 clocks per MMA against a 16-clock floor at Cin = 32; 50 % tensor-pipe activity at Cin = 64).  Swapped, the
    // M = 128 rows of A are the input channels of 128 / Cin taps (x boxes at different row shifts -- a tap is nothing
    // but a row shift) and dy is the B operand with N = 128 output channels: 1/4 or 1/2 of the MMAs for the same
    // shared-memory bytes each.  Then blockIdx.x = tap group, blockIdx.z = 128-channel block of Cout, the tile is dw^T.
    const int nblk = a.cin / a.bn;
    const bool swap = a.swap != 0;
    const int taps_all = a.kh * a.kw;
    const int tap = swap ? blockIdx.x * (TC_BM / (F16 ? 64 : 32)) : blockIdx.x / nblk;
    const int ci0 = swap ? 0 : (blockIdx.x - tap * nblk) * a.bn;
    const int co0 = blockIdx.z * TC_BM;
    const int dy_ = tap / a.kw, dx_ = tap - dy_ * a.kw;
    const long long shift = (long long)(dy_ - a.ph) * a.wp + (dx_ - a.pw);
    const long long k_begin = (long long)blockIdx.y * a.rows_per_split;
    long long k_end = k_begin + a.rows_per_split;
    if (k_end > a.rows) k_end = a.rows;
    constexpr bool f16 = F16;
    constexpr int krows = F16 ? 64 : 32;                    // pixel rows (K) per stage
    constexpr int bc = F16 ? 64 : 32;                       // channels per TMA box (128 bytes)
    constexpr uint32_t box_bytes = (uint32_t)krows * 128;
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
        // TMA producer: whole warp converged, one elected lane issues (see elect_one)
        constexpr int na_boxes = TC_BM / bc;
        const int nb_boxes = a.bn / bc;
        uint32_t s = 0, ph = 0;
        int q = (int)k_begin;
        const int qs = (int)shift;
        for (int it = 0; it < iters; ++it, q += krows) {
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem) + s * stage_bytes;
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, a.fold ? stage_bytes - a_bytes : stage_bytes);
                if (swap) {
                    // A: x boxes of taps tap .. tap + 3 (the host passes the x maps first); a tap past the last one
                    // reads rows far outside the tensor, which TMA fills with zeros
#pragma unroll
                    for (int j = 0; j < na_boxes; ++j) {
                        const int tj = tap + j, dyj = tj / a.kw, dxj = tj - dyj * a.kw;
                        const int rj = tj < taps_all ? q + (dyj - a.ph) * a.wp + (dxj - a.pw) : -(1 << 30);
                        tma_load_2d(sa + j * box_bytes, &tm_dyhi, fb, 0, rj);
                        if (!a.fold) tma_load_2d(sa + a_bytes + j * box_bytes, &tm_dylo, fb, 0, rj);
                    }
                    for (int j = 0; j < nb_boxes; ++j) {
                        // fold: dy rows are [64 hi | 64 lo] per output pixel, one 64-channel box per pixel
                        const int col = a.fold ? (co0 / bc + j) * a.fold_cs : co0 + bc * j;
                        tma_load_2d(sa + 2 * a_bytes + j * box_bytes, &tm_xhi, fb, col, q);
                        tma_load_2d(sa + 2 * a_bytes + b_bytes + j * box_bytes, &tm_xlo, fb, col, q);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < na_boxes; ++j) {
                        tma_load_2d(sa + j * box_bytes, &tm_dyhi, fb, co0 + bc * j, q);
                        tma_load_2d(sa + a_bytes + j * box_bytes, &tm_dylo, fb, co0 + bc * j, q);
                    }
                    for (int j = 0; j < nb_boxes; ++j) {
                        tma_load_2d(sa + 2 * a_bytes + j * box_bytes, &tm_xhi, fb, ci0 + bc * j, q + qs);
                        tma_load_2d(sa + 2 * a_bytes + b_bytes + j * box_bytes, &tm_xlo, fb, ci0 + bc * j, q + qs);
                    }
                }
            }
            __syncwarp();
            if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // MMA issuer; both operands MN-major: bits 15 and 16 of the instruction descriptor
        const uint32_t idesc = make_idesc(a.bn, f16) | (1u << 15) | (1u << 16);
        constexpr uint64_t kstep = (uint64_t)((F16 ? 2048 : 1024) >> 4);   // one MMA K step = two row groups
        const uint32_t sa0 = smem_u32(smem);
        const uint64_t dA_hi = make_mnmajor_desc(sa0, f16), dA_lo = make_mnmajor_desc(sa0 + a_bytes, f16);
        const uint64_t dB_hi = make_mnmajor_desc(sa0 + 2 * a_bytes, f16), dB_lo = make_mnmajor_desc(sa0 + 2 * a_bytes + b_bytes, f16);
        const uint32_t stage16 = stage_bytes >> 4;
        const uint32_t t_corr = tmem_base + 3 * a.bn;
        uint32_t s = 0, ph = 0, m = 0;                          // m: main accumulator of this stage (round-robin)
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t so = (uint64_t)(s * stage16);
                const uint32_t t_main = tmem_base + m * a.bn;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ko = so + (uint64_t)k * kstep;
                    if (!a.fold && (!a.single || !(it | k)))
                        umma<F16, 0>(t_corr, dA_lo + ko, dB_hi + ko, idesc, (it | k) ? 1u : 0u);
                    if (!a.single || (a.fold && !(it | k)))
                        umma<F16, 1>(t_corr, dA_hi + ko, dB_lo + ko, idesc, (a.fold && !(it | k)) ? 0u : 1u);
                    umma<F16, 2>(t_main, dA_hi + ko, dB_hi + ko, idesc, (it >= 3 || k) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * s);
                if (it == iters - 1) umma_commit(tfull);
            }
            __syncwarp();
            if (++m == 3) m = 0;
            if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
    } else {
        const int quad = warp & 3;
        const int co = co0 + quad * 32 + lane;
        const bool co_ok = co < a.cout;      // Cout % 128 == 64: the upper half of the last tile is TMA zero fill
        const int taps = a.kh * a.kw;
        // swap: TMEM lane m = (tap m / Cin, input channel m % Cin), column = output channel
        const int sm_ = quad * 32 + lane, st_ = sm_ / a.cin;
        float *orow = swap ? a.dw + ((size_t)co0 * taps + (tap + st_)) * a.cin + (sm_ - st_ * a.cin)
                           : a.dw + ((size_t)co * taps + tap) * a.cin + ci0;
        const bool row_ok = !swap || tap + st_ < taps;
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
            if (swap) {
                if (row_ok) {
                    const size_t cstride = (size_t)taps * a.cin;      // next output channel
                    // fold: TMEM lane m = (tap, pixel p, part, channel c); the lanes of the lo halves of x carry
                    // x_lo * dy_hi in the main accumulators (a correction term) and nothing useful in the other
                    const bool lo_lane = a.fold && ((sm_ >> 3) & 1);
                    const float cu = lo_lane ? 0.f : corr, cm = lo_lane ? corr : 1.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        atomicAdd(orow + (size_t)(c0 + j) * cstride,
                                  fmaf(__uint_as_float(u[j]), cu, __uint_as_float(v[j]) * cm) * inv);
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < ncol && co_ok)
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

// CTA-pair variant of wgrad_tc_kernel (fp16 operands, Cin % 128 == 0, Cout % 256 == 0, no operand swap): one
// tcgen05.mma.cta_group::2 computes a 256 x 128 tile of dw for one tap -- each CTA stages its own 128 dy channels (A)
// and only HALF of the x tile (64 of the 128 input channels, B), halving the B traffic through shared memory.  Barrier
// protocol as in conv_tc2_kernel: both producers complete on the leader's full barrier, the leader's commits arrive on
// the empty / tfull barriers of both CTAs; there is one tile per pair, so no accumulator hand-back.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tm_dyhi, const __grid_constant__ CUtensorMap tm_dylo,
                 const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo, TcWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * 8 + 1];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank_();
    const bool leader = rank == 0;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int krows = 64, bc = 64;                      // pixel rows (K) per stage, channels per TMA box
    constexpr uint32_t box_bytes = (uint32_t)krows * 128;
    const uint32_t a_bytes = 2 * box_bytes;                 // dy: 64 rows x 128 channels (two boxes)
    const uint32_t b_bytes = box_bytes;                     // x:  64 rows x 64 channels (this CTA's half of the 128)
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const int S = a.stages;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), tfull = smem_u32(&bars[16]);

    const int nblk = a.cin / 128;
    const int pairx = blockIdx.x >> 1;                      // (tap, 128-channel block of Cin)
    const int tap = pairx / nblk;
    const int ci0 = (pairx - tap * nblk) * 128;
    const int co0 = blockIdx.z * 256 + (int)rank * 128;
    const int dy_ = tap / a.kw, dx_ = tap - dy_ * a.kw;
    const long long shift = (long long)(dy_ - a.ph) * a.wp + (dx_ - a.pw);
    const long long k_begin = (long long)blockIdx.y * a.rows_per_split;
    long long k_end = k_begin + a.rows_per_split;
    if (k_end > a.rows) k_end = a.rows;
    const int iters = k_end > k_begin ? (int)((k_end - k_begin + krows - 1) / krows) : 0;
    const int nmain = iters < 3 ? iters : 3;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (iters > 0) {            // uniform for the pair
        if (warp == 0) {
            uint32_t s = 0, ph = 0;
            int q = (int)k_begin;
            const int qs = (int)shift;
            for (int it = 0; it < iters; ++it, q += krows) {
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem) + s * stage_bytes;
                    const uint32_t fb = map_to_rank(full0 + 8 * s, 0);
                    if (leader) mbar_expect_tx(full0 + 8 * s, 2 * stage_bytes);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        tma_load_2d_pair(sa + j * box_bytes, &tm_dyhi, fb, co0 + bc * j, q);
                        tma_load_2d_pair(sa + a_bytes + j * box_bytes, &tm_dylo, fb, co0 + bc * j, q);
                    }
                    tma_load_2d_pair(sa + 2 * a_bytes, &tm_xhi, fb, ci0 + (int)rank * bc, q + qs);
                    tma_load_2d_pair(sa + 2 * a_bytes + b_bytes, &tm_xlo, fb, ci0 + (int)rank * bc, q + qs);
                }
                __syncwarp();
                if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
            }
        } else if (warp == 1) {
            if (leader) {
                // fp32 accumulate, f16 operands, both MN-major, M = 256, N = 128
                const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24) | (1u << 15) | (1u << 16);
                constexpr uint64_t kstep = (uint64_t)(2048 >> 4);
                const uint32_t sa0 = smem_u32(smem);
                const uint64_t dA_hi = make_mnmajor_desc(sa0, true), dA_lo = make_mnmajor_desc(sa0 + a_bytes, true);
                const uint64_t dB_hi = make_mnmajor_desc(sa0 + 2 * a_bytes, true), dB_lo = make_mnmajor_desc(sa0 + 2 * a_bytes + b_bytes, true);
                const uint32_t stage16 = stage_bytes >> 4;
                const uint32_t t_corr = tmem_base + 3 * 128;
                uint32_t s = 0, ph = 0, m = 0;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint64_t so = (uint64_t)(s * stage16);
                        const uint32_t t_main = tmem_base + m * 128;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ko = so + (uint64_t)k * kstep;
                            if (!a.single || !(it | k)) umma2<0>(t_corr, dA_lo + ko, dB_hi + ko, idesc, (it | k) ? 1u : 0u);
                            if (!a.single) umma2<1>(t_corr, dA_hi + ko, dB_lo + ko, idesc, 1u);
                            umma2<2>(t_main, dA_hi + ko, dB_hi + ko, idesc, (it >= 3 || k) ? 1u : 0u);
                        }
                        umma2_commit_both(empty0 + 8 * s);
                        if (it == iters - 1) umma2_commit_both(tfull);
                    }
                    __syncwarp();
                    if (++m == 3) m = 0;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
                }
            }
        } else {
            const int quad = warp & 3;
            const int co = co0 + quad * 32 + lane;
            const int taps = a.kh * a.kw;
            float *orow = a.dw + ((size_t)co * taps + tap) * a.cin + ci0;
            const float inv = (1.f / f16_scale_from_bound(*a.xb)) * (1.f / f16_scale_from_bound(*a.yb));
            const float corr = 1.f / 2048.f;
            mbar_wait(tfull, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v[32], u[32];
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
                tmem_ld32(trow, v);
                for (int m = 1; m < nmain; ++m) {
                    tmem_ld32(trow + m * 128, u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
                }
                tmem_ld32(trow + 3 * 128, u);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    atomicAdd(reinterpret_cast<float4 *>(orow + c0 + j),
                              make_float4(fmaf(__uint_as_float(u[j]), corr, __uint_as_float(v[j])) * inv,
                                          fmaf(__uint_as_float(u[j + 1]), corr, __uint_as_float(v[j + 1])) * inv,
                                          fmaf(__uint_as_float(u[j + 2]), corr, __uint_as_float(v[j + 2])) * inv,
                                          fmaf(__uint_as_float(u[j + 3]), corr, __uint_as_float(v[j + 3])) * inv));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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
    if (a.cout % bc != 0 || a.cin % bc != 0) return 0;
    // x and dy must share one padded grid, with pads covering the kernel reach
    if (a.x.n != a.y.n || a.x.h != a.y.h || a.x.w != a.y.w || a.x.ph != a.y.ph || a.x.pw != a.y.pw) return 0;
    if (a.x.ph < a.ph || a.x.pw < a.pw) return 0;
    int bn = 0;
    for (int c : {128, 64, 32})
        if (a.cin % c == 0 && c % bc == 0) { bn = c; break; }
    if (!bn) return 0;
    const bool swap = a.cin == bc && a.cout % TC_BM == 0;   // see the kernel: x tiles of 128 / Cin taps as A, dy as B
    if (a.fold && !(f16 && swap)) return 0;
    const int tpg = TC_BM / bc;                              // taps per CTA in swap mode
    if (swap) bn = TC_BM;
    const long long rows = (long long)a.x.n * a.x.hp * a.x.wp;
    if (rows >= (1LL << 31) - 4096) return 0;
    if ((((uintptr_t)a.x_hi | (uintptr_t)a.x_lo | (uintptr_t)a.w_hi | (uintptr_t)a.w_lo | (uintptr_t)a.out |
          (uintptr_t)a.x_h2 | (uintptr_t)a.w_h2) & 15) != 0) return 0;

    const int taps = a.kh * a.kw;
    const int mblocks = (a.cout + TC_BM - 1) / TC_BM;
    const int tiles = swap ? ((taps + tpg - 1) / tpg) * mblocks : taps * (a.cin / bn) * mblocks;
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
    t.cin = a.cin; t.cout = a.cout; t.bn = bn; t.swap = swap ? 1 : 0;
    t.rows = rows; t.rows_per_split = rps; t.dw = a.out;
    t.f16 = f16 ? 1 : 0; t.xb = a.x_bound; t.yb = a.w_bound;
    t.single = g_bwd_single ? 1 : 0;
    t.fold = a.fold; t.fold_cs = 2 * bc;
    const int stage_bytes = 2 * TC_BM * TC_BK * 4 + 2 * bn * TC_BK * 4;
    int stages = TC_SMEM_LIMIT / stage_bytes;
    if (stages > 6) stages = 6;
    t.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;

    CUtensorMap mdh, mdl, mxh, mxl;
    int rc;
    const CUtensorMapSwizzle sw = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if (a.fold) {
        // dy: rows of (cout / 64) output pixels x [64 hi | 64 lo]; x: one plane, rows of 64 halves
        const long long dcols = 2LL * a.cout;
        if ((rc = make_map(&mdh, a.w_h2, true, rows, dcols, dcols, krows, sw))) return rc;
        if ((rc = make_map(&mdl, a.w_h2 + bc, true, rows, dcols - bc, dcols, krows, sw))) return rc;
        if ((rc = make_map(&mxh, a.x_h2, true, rows, a.cin, a.cin, krows, sw))) return rc;
        mxl = mxh;
    } else {
        if ((rc = make_map_pair(&mdh, &mdl, a.w_hi, a.w_lo, a.w_h2, rows, a.cout, krows, sw))) return rc;
        if ((rc = make_map_pair(&mxh, &mxl, a.x_hi, a.x_lo, a.x_h2, rows, a.cin, krows, sw))) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        DLIO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
        DLIO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
        attr_set = true;
    }
    ProfScope prof(DLIO_PROF_CONV_WGRAD_TC, st);
    DLIO_CUDA(cudaMemsetAsync(a.out, 0, (size_t)a.cout * taps * a.cin * sizeof(float), st));
    static const bool wgrad_pairs = [] { const char *e = getenv("DLIO_WGRAD_CG2"); return !(e && e[0] == '0'); }();
    if (f16 && !swap && bn == 128 && a.cout % 256 == 0 && wgrad_pairs && conv_cg2_enabled()) {
        // CTA pairs (wgrad_tc2_kernel): 256 x 128 tiles; K splits sized for whole waves of 74 pairs
        const int tiles2 = taps * (a.cin / 128) * (a.cout / 256);
        int splits2 = 1;
        double best2 = 0.0;
        for (int waves = 2; waves <= 4; ++waves) {
            int s2 = waves * 74 / tiles2;
            if (s2 > max_splits) s2 = (int)max_splits;
            if (s2 < 1) s2 = 1;
            const long long ctas = (long long)s2 * tiles2;
            const double fill = (double)ctas / (double)(((ctas + 73) / 74) * 74);
            if (fill > best2 + 1e-9) { best2 = fill; splits2 = s2; }
        }
        long long rps2 = (rows + splits2 - 1) / splits2;
        rps2 = (rps2 + krows - 1) / krows * krows;
        if ((rows + rps2 - 1) / rps2 < splits2) splits2 = (int)((rows + rps2 - 1) / rps2);
        t.rows_per_split = rps2;
        const int stage2 = 2 * (2 * 64 * 128) + 2 * (64 * 128);       // dy hi|lo (128 ch) + x hi|lo (64 ch)
        int stages2 = TC_SMEM_LIMIT / stage2;
        if (stages2 > 6) stages2 = 6;
        t.stages = stages2;
        static bool attr2 = false;
        if (!attr2) {
            DLIO_CUDA(cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT + 1024));
            attr2 = true;
        }
        dim3 grid2((unsigned)(2 * taps * (a.cin / 128)), (unsigned)splits2, (unsigned)(a.cout / 256));
        wgrad_tc2_kernel<<<grid2, TC_THREADS, (size_t)stages2 * stage2 + 1024, st>>>(mdh, mdl, mxh, mxl, t);
        DLIO_LAUNCH_CHECK();
        return 1;
    }
    dim3 grid((unsigned)(swap ? (taps + tpg - 1) / tpg : taps * (a.cin / bn)), (unsigned)splits, (unsigned)mblocks);
    if (f16 && swap) wgrad_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(mxh, mxl, mdh, mdl, t);   // x as A, dy as B
    else if (f16) wgrad_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(mdh, mdl, mxh, mxl, t);
    else if (swap) wgrad_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(mxh, mxl, mdh, mdl, t);   // x as A, dy as B
    else wgrad_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(mdh, mdl, mxh, mxl, t);
    DLIO_LAUNCH_CHECK();
    return 1;
}

}  // namespace dlio
