// igemm.cu -- stage 2 (sparse_cor, R/TADpole.R:94-100) for integer contact counts: the Gram matrix X^T X = X X^T on
// the 5th-generation tensor cores, exactly.
//
// tcgen05.mma has no FP64 kind, but Hi-C contact matrices are integer counts.  Every count v (|v| < 2^20) is cut into
// three balanced base-128 digits v = d0 + 128 d1 + 128^2 d2, d in [-64, 63], stored as three int8 matrices.  The
// nine digit products are accumulated in INT32 by tcgen05.mma.kind::i8, grouped by scale: D_s = sum_{a+b=s} X_a X_b^T,
// s = 0..4, five accumulators in tensor memory (5 x 64 columns).  |digit product| <= 4096 and at most 3 n products
// meet in one accumulator, so nothing overflows below n = 174 000; the epilogue recombines
// G = sum_s 128^s D_s in FP64 (exact while G < 2^53) and applies the reference's covariance -> correlation -> NaN->0
// arithmetic in the reference's order, so the only rounding left in stage 2 is the reference's own epilogue.
//
// One CTA per 128 x 64 tile of the upper triangle (tiles below the diagonal are mirrored on store).  Warp 0: TMA
// producer (cp.async.bulk.tensor, SWIZZLE_128B, 3-stage mbarrier ring, 72 KB per stage: three digit tiles of A and
// of B).  Warp 1: allocates tensor memory and issues the MMAs (one elected thread, 36 per k-block of 128).
// Warps 2-5: epilogue (tcgen05.ld 32x32b, one output row per thread).  Non-integer input (or counts >= 2^20) is
// detected by the slicing pass and sent to the FP64 DMMA path instead.
#include "common.cuh"
#include "gemm.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <algorithm>

#define IG_BM 128
#define IG_BN 64
#define IG_BK 128                 // int8 elements = bytes: one 128-byte swizzle row
#define IG_STAGES 3
#define IG_THREADS 192
#define IG_A_BYTES (IG_BM * IG_BK)            // one digit tile of A: 16 KB
#define IG_B_BYTES (IG_BN * IG_BK)            // one digit tile of B: 8 KB
#define IG_STAGE_BYTES (3 * IG_A_BYTES + 3 * IG_B_BYTES)
#define IG_TMEM_COLS 512

// ---- digit slicing -------------------------------------------------------------------------------------------------
// S[s][row][k], row pitch Kp bytes, rows / columns beyond n zero; flag |= 1 when a value is not an integer below 2^20
__global__ void __launch_bounds__(256)
ig_slice_kernel(const double *__restrict__ X, int n, int ld, int8_t *__restrict__ S, int rows_pad, int Kp, int *flag) {
    const size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // 16 consecutive k of one row
    const int cpr = Kp / 16;
    if (chunk >= (size_t)rows_pad * cpr) return;
    const int row = (int)(chunk / cpr), k0 = (int)(chunk % cpr) * 16;
    alignas(16) int8_t d[3][16];
    int bad = 0;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const int k = k0 + t;
        double x = (row < n && k < n) ? X[(size_t)row * ld + k] : 0.0;
        const double r = rint(x);
        if (!(r == x) || !(fabs(x) < 1048576.0)) { bad = 1; x = 0.0; }
        int v = (int)r;
        const int d0 = ((v + 64) & 127) - 64; v = (v - d0) >> 7;
        const int d1 = ((v + 64) & 127) - 64; v = (v - d1) >> 7;
        d[0][t] = (int8_t)d0; d[1][t] = (int8_t)d1; d[2][t] = (int8_t)v;
    }
    const size_t plane = (size_t)rows_pad * Kp;
#pragma unroll
    for (int s = 0; s < 3; s++)
        *reinterpret_cast<int4 *>(S + s * plane + (size_t)row * Kp + k0) = *reinterpret_cast<const int4 *>(d[s]);
    if (bad) atomicOr(flag, 1);
}

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ig_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ig_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ig_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ig_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ig_tma_load_3d(unsigned dst, const CUtensorMap *tmap, unsigned bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void ig_mma_i8(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ig_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ig_tmem_ld16(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 bytes apart (SBO), LBO unused
__device__ __forceinline__ unsigned long long ig_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);            // start address, 16-byte units
    d |= (unsigned long long)1 << 16;                                // leading byte offset (ignored for swizzled K-major)
    d |= (unsigned long long)(1024 >> 4) << 32;                      // stride byte offset
    d |= (unsigned long long)1 << 46;                                // descriptor version (sm_100)
    d |= (unsigned long long)2 << 61;                                // SWIZZLE_128B
    return d;
}

struct IgParams {
    int n;                 // valid rows / columns
    int row_begin, row_end;   // rows of the output this launch computes (a rank's row block; all rows on one GPU)
    SymShard ss;           // who computes which block pair (common.cuh); R = 1: plain symmetric launch
    int kblocks;           // Kp / 128
    double *C; long ldc;   // output, row-major
    const double *mean, *sd; double nrows;
    int raw;               // 1: store the Gram matrix itself (test hook)
};

__global__ void __launch_bounds__(IG_THREADS, 1)
ig_gram_kernel(const __grid_constant__ CUtensorMap tmap, IgParams p) {
    // tiles walked in bands of 8 tile rows, column by column: the ~148 CTAs resident together share 8 A row blocks and ~18 B
    // row blocks through L2 instead of one A block and 148 B blocks (ncu at 8000 bins, round 1: 3.3 GB of DRAM reads per
    // launch against 0.19 GB of operand planes)
    unsigned bx = blockIdx.x, by = blockIdx.y;
    {
        const unsigned id = blockIdx.y * gridDim.x + blockIdx.x, per = 8u * gridDim.x;
        const unsigned band = id / per, in = id % per;
        const unsigned h = gridDim.y - band * 8u < 8u ? gridDim.y - band * 8u : 8u;
        by = band * 8u + in % h; bx = in / h;
    }
    const int m0 = p.row_begin + (int)by * IG_BM, n0 = (int)bx * IG_BN;
    if (m0 >= p.row_end || n0 >= p.n) return;                        // padding
    {   // computed by another rank / mirrored from the tile above the diagonal (SymShard, common.cuh)
        const int r_hi = m0 + IG_BM - 1 < p.row_end - 1 ? m0 + IG_BM - 1 : p.row_end - 1;
        const int c_hi = n0 + IG_BN - 1 < p.n - 1 ? n0 + IG_BN - 1 : p.n - 1;
        if (!ss_tile_needed(m0, r_hi, n0, c_hi, p.ss)) return;
    }
    extern __shared__ unsigned char ig_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)ig_raw + 1023) & ~(uintptr_t)1023);    // 1024-byte aligned
    __shared__ __align__(8) unsigned long long s_full[IG_STAGES], s_empty[IG_STAGES], s_done;
    __shared__ unsigned s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < IG_STAGES; s++) { ig_mbar_init(ig_smem(&s_full[s]), 1); ig_mbar_init(ig_smem(&s_empty[s]), 1); }
        ig_mbar_init(ig_smem(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ig_smem(&s_tmem)), "n"(IG_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < p.kblocks; kb++) {
                const int st = kb % IG_STAGES;
                ig_mbar_wait(ig_smem(&s_empty[st]), ((kb / IG_STAGES) & 1) ^ 1);
                const unsigned bar = ig_smem(&s_full[st]);
                ig_mbar_expect_tx(bar, IG_STAGE_BYTES);
                const unsigned base = ig_smem(tiles + (size_t)st * IG_STAGE_BYTES);
                for (int d = 0; d < 3; d++) {
                    ig_tma_load_3d(base + d * IG_A_BYTES, &tmap, bar, kb * IG_BK, m0, d);
                    ig_tma_load_3d(base + d * IG_A_BYTES + IG_A_BYTES / 2, &tmap, bar, kb * IG_BK, m0 + 64, d);
                    ig_tma_load_3d(base + 3 * IG_A_BYTES + d * IG_B_BYTES, &tmap, bar, kb * IG_BK, n0, d);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D_{a+b} += A_a B_b^T =====
        // instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N = 64, M = 128
        const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((IG_BN >> 3) << 17) | ((IG_BM >> 4) << 24);
        const unsigned idesc3 = (2u << 4) | (1u << 7) | (1u << 10) | (((3 * IG_BN) >> 3) << 17) | ((IG_BM >> 4) << 24);    // N = 192
        if (lane == 0) {
            for (int kb = 0; kb < p.kblocks; kb++) {
                const int st = kb % IG_STAGES;
                ig_mbar_wait(ig_smem(&s_full[st]), (kb / IG_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned base = ig_smem(tiles + (size_t)st * IG_STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < IG_BK / 32; kk++) {
                    // The three B digit tiles of a stage are contiguous (64 rows x 128 B each) and so are the accumulators
                    // of scales a, a+1, a+2: one N = 192 MMA per A digit multiplies it with all three B digits, reading the
                    // A tile from shared memory once instead of three times (the N = 64 form is bound by that traffic:
                    // 6 KB per 32 tensor cycles against 128 B / cycle).  The very first k-step has to start accumulators
                    // 3 and 4 at zero while 1 and 2 already accumulate, so it keeps the per-pair form.
                    const unsigned long long bd0 = ig_desc(base + 3 * IG_A_BYTES) + (unsigned long long)(kk * 32 >> 4);
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        const unsigned long long ad = ig_desc(base + a * IG_A_BYTES) + (unsigned long long)(kk * 32 >> 4);
                        if (kb == 0 && kk == 0) {
#pragma unroll
                            for (int b = 0; b < 3; b++) {
                                const unsigned long long bd = ig_desc(base + 3 * IG_A_BYTES + b * IG_B_BYTES) + (unsigned long long)(kk * 32 >> 4);
                                const bool first = a == 0 || b == 2;                         // first product of scale a + b
                                ig_mma_i8(tmem + (unsigned)(a + b) * IG_BN, ad, bd, idesc, first ? 0u : 1u);
                            }
                        } else {
                            ig_mma_i8(tmem + (unsigned)a * IG_BN, ad, bd0, idesc3, 1u);
                        }
                    }
                }
                ig_commit(ig_smem(&s_empty[st]));          // frees the stage once these MMAs have read it
            }
            ig_commit(ig_smem(&s_done));                   // accumulators complete
        }
    } else {
        // ===== epilogue: one output row per thread =====
        ig_mbar_wait(ig_smem(&s_done), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lg = warp & 3;                           // tensor-memory lane group this warp may read
        const int row = m0 + lg * 32 + lane;
        const bool rok = row < p.row_end;
        const int rblk = row / p.ss.rpr;
        const double mi = (rok && !p.raw) ? p.mean[row] : 0.0, si = (rok && !p.raw) ? p.sd[row] : 1.0;
        for (int c0 = 0; c0 < IG_BN; c0 += 16) {
            unsigned r[5][16];
#pragma unroll
            for (int s = 0; s < 5; s++) ig_tmem_ld16(tmem + ((unsigned)(lg * 32) << 16) + (unsigned)(s * IG_BN + c0), r[s]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (rok) {
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const int col = n0 + c0 + c;
                    if (col >= p.n) continue;
                    double g = (double)(int)r[4][c];
                    g = fma(g, 128.0, (double)(int)r[3][c]);
                    g = fma(g, 128.0, (double)(int)r[2][c]);
                    g = fma(g, 128.0, (double)(int)r[1][c]);
                    g = fma(g, 128.0, (double)(int)r[0][c]);
                    double v = g;
                    if (!p.raw) {
                        // (crossprod - nrow * tcrossprod(colMeans)) / (nrow - 1), then / tcrossprod(sd); NaN -> 0
                        v = (g - p.nrows * (mi * p.mean[col])) / (p.nrows - 1.0);
                        v = v / (si * p.sd[col]);
                        v = nan_to_zero(v);
                    }
                    p.C[(size_t)row * p.ldc + col] = v;
                    // mirror inside the owner's diagonal block (its tiles below the diagonal are skipped)
                    if (col != row && col / p.ss.rpr == rblk) p.C[(size_t)col * p.ldc + row] = v;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(IG_TMEM_COLS));
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int ig_encode(CUtensorMap *map, void *base, int rows_pad, int Kp) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        if (!sym || qres != cudaDriverEntryPointSuccess) { tp_set_error("cuTensorMapEncodeTiled not available"); return TP_ERR_CUDA; }
        fn = (PFN_encodeTiled)sym;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows_pad, 3};
    const cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * rows_pad};          // bytes, dims 1 and 2
    const cuuint32_t box[3] = {IG_BK, 64, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tp_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TP_ERR_CUDA; }
    return TP_OK;
}

// C = correlation (or raw Gram when raw != 0) of the n x n symmetric matrix X of integer counts.  *used_out = 0 when
// the input is not integer (nothing written to C): the caller takes the FP64 DMMA path.
int tp_igram(tp_ctx *ctx, const double *X, int n, int ld, double *C, int ldc, const double *mean, const double *sd,
             int raw, int *used_out, int row_begin, int row_end, SymShard ss) {
    cudaStream_t st = ctx->stream;
    const int rows_pad = round_up(n, IG_BM), Kp = round_up(n, IG_BK);
    const size_t plane = (size_t)rows_pad * Kp;
    TP_TRY(ctx->islices.reserve(3 * plane + 64));
    int8_t *S = ctx->islices.as<int8_t>();
    int *flag = (int *)(S + 3 * plane);                       // (3 * plane is a multiple of 128)
    TP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    const size_t chunks = (size_t)rows_pad * (Kp / 16);
    tp_prof_begin(ctx, PC_ISLICE);
    ig_slice_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(X, n, ld, S, rows_pad, Kp, flag);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_TRY(tp_pin_reserve(ctx, 64));
    int *h = (int *)ctx->pin;
    TP_CUDA(cudaMemcpyAsync(h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    TP_CUDA(tp_stream_sync(ctx));
    if (h[0]) { *used_out = 0; return TP_OK; }
    CUtensorMap map;
    TP_TRY(ig_encode(&map, S, rows_pad, Kp));
    IgParams p;
    p.n = n; p.kblocks = Kp / IG_BK; p.C = C; p.ldc = ldc; p.mean = mean; p.sd = sd; p.nrows = (double)n; p.raw = raw;
    p.row_begin = row_begin; p.row_end = row_end; p.ss = ss;
    const size_t smem = (size_t)IG_STAGES * IG_STAGE_BYTES + 1024;
    TP_CUDA(tp_optin_smem(ig_gram_kernel, ctx));
    if (row_end > row_begin) {
        dim3 grid(rows_pad / IG_BN, (row_end - row_begin + IG_BM - 1) / IG_BM);
        tp_prof_begin(ctx, PC_IGEMM);
        if (ctx->prof) {       // executed: 9 digit products on the tiles this launch computes
            double tiles = 0.0;
            for (int m0 = row_begin; m0 < row_end; m0 += IG_BM)
                for (int n0 = 0; n0 < n; n0 += IG_BN)
                    tiles += ss_tile_needed(m0, std::min(m0 + IG_BM, row_end) - 1, n0, std::min(n0 + IG_BN, n) - 1, ss);
            ctx->prof_imma_ops += 2.0 * 9.0 * tiles * IG_BM * IG_BN * (double)Kp;
        }
        ig_gram_kernel<<<grid, IG_THREADS, smem, st>>>(map, p);
        tp_prof_end(ctx);
        ctx->launches += 1;
    }
    TP_CUDA(cudaGetLastError());
    *used_out = 1;
    return TP_OK;
}

// =====================================================================================================================
// Sliced FP64 x FP64 product for the operator applications of stage 3 (Ozaki-style error-free digit products):
//   Yout = alpha * S * Yin + beta * E1 + gamma * E2,   S symmetric n x n (M = Xc Xc^T), Yin n x b.
// Row i of S is scaled by 2^-e_i (e_i = exponent of its largest entry) and cut into NP digits of 7 bits,
// a = 2^e sum_s d_s 2^(-6-7s), |d_s| <= 64; column j of Yin likewise with exponent f_j (planes stored transposed,
// K-major).  The digit products with s + t <= NP - 1 are accumulated exactly in INT32 by tcgen05.mma.kind::i8, one
// accumulator per scale g = s + t, and recombined in FP64 in the epilogue.  The digit pairs left out are below
// 2^(-7 NP) of the row scale x column scale: NP = 5 (3e-11) carries the early filter rounds of the subspace iteration
// while the residual is far above that; NP = 8 (1.4e-17, FP64 level) carries the last rounds and the residual checks
// that decide convergence where nf >= iop_final_min_n (default), the FP64 DMMA operator elsewhere (pca.cu).
// Tiles 128 x 64 x 64 (SWIZZLE_64B rows of 64 bytes), stage = NP x (8 KB + 4 KB), 3 stages.
// =====================================================================================================================
#define IO_BM 128
#define IO_BN 64                  // widest output tile (the planes of the block are padded to multiples of it)
#define IO_BK 64
#define IO_MAXNP 8                // digit planes kept of the operator; a product uses the first NP of them (5 or 8)
#define IO_A_BYTES (IO_BM * IO_BK)
// BN = 64: one CTA per SM (all 512 tensor-memory columns at NP = 8).  BN = 32: half the tensor memory and half the B tile,
// so at NP = 5 two CTAs share an SM (one's epilogue under the other's MMAs) and a 2000 x 256 application is 128 CTAs
// instead of 64 on the 148 SMs; used for the small applications (tiles of 64 columns would not fill the GPU).
template <int NP, int BN> struct IoCfg {
    static constexpr int STAGES = BN == 64 ? (NP <= 5 ? 3 : 2) : 2;
    static constexpr int B_BYTES = BN * IO_BK;
    static constexpr int STAGE_BYTES = NP * (IO_A_BYTES + B_BYTES);
    static constexpr int TMEM_COLS = NP * BN <= 256 ? 256 : 512;
    static constexpr int NGROUP = 256 / BN;      // adjacent B digit planes multiplied by one MMA (N = BN x group <= 256)
    static constexpr int CTAS_PER_SM = (BN == 32 && NP <= 5) ? 2 : 1;
};

// exponent e with max < 2^e (max > 0), else 0; scale arrays hold 2^(e-6)
__device__ __forceinline__ int io_exponent(double mx) {
    if (!(mx > 0.0) || !isfinite(mx)) return 0;
    int e;
    frexp(mx, &e);                 // mx = f 2^e, f in [0.5, 1)
    return e;
}

// one warp per row: largest |entry| of row r of A (n x n, ld) -> exponent and scale
__global__ void io_rowmax_kernel(const double *__restrict__ A, int n, int ld, int *__restrict__ expo,
                                 double *__restrict__ scale) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const double *row = A + (size_t)w * ld;
    double mx = 0.0;
    for (int c = lane; c < n; c += 32) mx = fmax(mx, fabs(row[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) { const int e = io_exponent(mx); expo[w] = e; scale[w] = ldexp(1.0, e - 6); }
}

// largest |entry| of every column of Y (n x b, ld) -> exponent and scale of the column.  At most two CTAs per SM walk the
// 16-row slabs (16 independent coalesced loads in flight per thread), one atomicMax per column and CTA on the IEEE bit
// pattern (non-negative doubles order like unsigned integers; max is exact and order independent); the CTA that finishes
// last turns the maxima into exponents / scales and clears the scratch for the next application (no memset, no second
// launch: five stream operations per operator application were three too many).
#define IO_CM_ROWS 16
__global__ void __launch_bounds__(256)
io_colmax_kernel(const double *__restrict__ Y, int n, int b, int ld, unsigned long long *__restrict__ colmax_bits,
                 unsigned *__restrict__ ticket, int *__restrict__ expo, double *__restrict__ scale) {
    __shared__ int s_last;
    for (int c = threadIdx.x; c < b; c += 256) {
        double mx = 0.0;
        for (int r0 = blockIdx.x * IO_CM_ROWS; r0 < n; r0 += gridDim.x * IO_CM_ROWS) {
            double v[IO_CM_ROWS];
#pragma unroll
            for (int t = 0; t < IO_CM_ROWS; t++) v[t] = (r0 + t < n) ? Y[(size_t)(r0 + t) * ld + c] : 0.0;
#pragma unroll
            for (int t = 0; t < IO_CM_ROWS; t++) mx = fmax(mx, fabs(v[t]));
        }
        if (mx > 0.0 && isfinite(mx)) atomicMax(colmax_bits + c, (unsigned long long)__double_as_longlong(mx));
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < b; c += 256) {
        const int e = io_exponent(__longlong_as_double((long long)__ldcg(colmax_bits + c)));
        expo[c] = e; scale[c] = ldexp(1.0, e - 6);
        colmax_bits[c] = 0ull;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

template <int NP>
__device__ __forceinline__ void io_digits(double a, int e, int8_t (&d)[NP]) {
    double x = ldexp(a, 6 - e);                 // |x| <= 64
    if (!isfinite(x)) x = 0.0;
#pragma unroll
    for (int s = 0; s < NP; s++) {
        const double r = rint(x);
        d[s] = (int8_t)(int)r;
        x = (x - r) * 128.0;                    // |x - r| <= 0.5 -> next digit in [-64, 64]
    }
}

// planes P[s][row][k] (row pitch Kp bytes) of the rows of A: 16 consecutive k per thread
template <int NP>
__global__ void __launch_bounds__(256)
io_slice_rows_kernel(const double *__restrict__ A, int n, int ld, const int *__restrict__ expo, int8_t *__restrict__ P,
                     int rows_pad, int Kp) {
    const size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cpr = Kp / 16;
    if (chunk >= (size_t)rows_pad * cpr) return;
    const int row = (int)(chunk / cpr), k0 = (int)(chunk % cpr) * 16;
    alignas(16) int8_t d[NP][16];
    const int e = row < n ? expo[row] : 0;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const int k = k0 + t;
        int8_t dd[NP];
        io_digits<NP>((row < n && k < n) ? A[(size_t)row * ld + k] : 0.0, e, dd);
#pragma unroll
        for (int s = 0; s < NP; s++) d[s][t] = dd[s];
    }
    const size_t plane = (size_t)rows_pad * Kp;
#pragma unroll
    for (int s = 0; s < NP; s++)
        *reinterpret_cast<int4 *>(P + s * plane + (size_t)row * Kp + k0) = *reinterpret_cast<const int4 *>(d[s]);
}

// planes P[s][j][k] of the COLUMNS of Y (n x b, ld): tiles of 128 k x 32 j transposed through shared memory.  A thread
// converts four consecutive k of one column (loads coalesced along j) and packs their digits into one 32-bit word per
// plane; the planes leave in 16-byte stores, eight threads to a 128-byte run along k.  (32 x 32 tiles with byte stores
// in 32-byte runs ran at a sixth of the HBM rate.)
template <int NP>
__global__ void __launch_bounds__(256)
io_slice_cols_kernel(const double *__restrict__ Y, int n, int b, int ld, const int *__restrict__ expo,
                     int8_t *__restrict__ P, int rows_pad, int Kp) {
    __shared__ unsigned s[NP][32][33];              // [plane][j][k / 4], pitch 33 words: conflict-free both ways
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k0 = blockIdx.x * 128, j0 = blockIdx.y * 32;
    const int j = j0 + tx;
    const int e = j < b ? expo[j] : 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int kq = ty + 8 * i;                  // group of four k
        unsigned w[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) w[p] = 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = k0 + 4 * kq + q;
            int8_t dd[NP];
            io_digits<NP>((k < n && j < b) ? Y[(size_t)k * ld + j] : 0.0, e, dd);
#pragma unroll
            for (int p = 0; p < NP; p++) w[p] |= (unsigned)(unsigned char)dd[p] << (8 * q);
        }
#pragma unroll
        for (int p = 0; p < NP; p++) s[p][tx][kq] = w[p];
    }
    __syncthreads();
    const size_t plane = (size_t)rows_pad * Kp;
    for (int item = threadIdx.x; item < NP * 256; item += 256) {
        const int p = item >> 8, jj = (item >> 3) & 31, k16 = item & 7;
        const int jo = j0 + jj, k = k0 + 16 * k16;
        if (jo < rows_pad && k < Kp) {
            int4 v;
            v.x = (int)s[p][jj][4 * k16]; v.y = (int)s[p][jj][4 * k16 + 1];
            v.z = (int)s[p][jj][4 * k16 + 2]; v.w = (int)s[p][jj][4 * k16 + 3];
            *reinterpret_cast<int4 *>(P + p * plane + (size_t)jo * Kp + k) = v;
        }
    }
}

__device__ __forceinline__ unsigned long long io_desc(unsigned smem_addr) {      // K-major, 64-byte rows, SWIZZLE_64B
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)(512 >> 4) << 32;                       // 8 rows x 64 bytes
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)4 << 61;                                // SWIZZLE_64B
    return d;
}

struct IoParams {
    int n, b;              // rows of S / of the output, columns of Y
    int row_begin, row_end;
    int kblocks;
    double *D; long ldd;
    const double *E1; long lde1; const double *E2; long lde2;
    double alpha, beta, gamma;
    const double *rowscale, *colscale;       // 2^(e_i - 6), 2^(f_j - 6)
    int banded;            // 1: tile order in bands of 8 tile rows (see the kernel)
    int sym;               // 1: symmetric product (B = A, square output): only the tiles SymShard gives this row block are
                           // computed, elements of the owner's diagonal block are stored twice, as in ig_gram_kernel
    SymShard ss;
};

template <int NP, int BN>
__global__ void __launch_bounds__(IG_THREADS, (IoCfg<NP, BN>::CTAS_PER_SM))
io_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, IoParams p) {
    constexpr int IO_STAGES = IoCfg<NP, BN>::STAGES, IO_STAGE_BYTES = IoCfg<NP, BN>::STAGE_BYTES, IO_NP = NP;
    constexpr int IO_B_BYTES = IoCfg<NP, BN>::B_BYTES, IO_NGROUP = IoCfg<NP, BN>::NGROUP, TMEM_COLS = IoCfg<NP, BN>::TMEM_COLS;
    unsigned bx = blockIdx.x, by = blockIdx.y;
    if (p.banded) {
        // wide outputs (Gram): bands of 8 tile rows walked column by column, so that the CTAs resident together share
        // 8 A row blocks and ~18 B row blocks through L2 instead of one A and 148 B
        const unsigned id = blockIdx.y * gridDim.x + blockIdx.x, per = 8u * gridDim.x;
        const unsigned band = id / per, in = id % per;
        const unsigned h = gridDim.y - band * 8u < 8u ? gridDim.y - band * 8u : 8u;
        by = band * 8u + in % h; bx = in / h;
    }
    const int m0 = p.row_begin + (int)by * IO_BM, n0 = (int)bx * BN;
    if (n0 >= p.b || m0 >= p.row_end) return;                        // padding
    if (p.sym) {    // computed by another rank / mirrored from the tile above the diagonal (SymShard, common.cuh)
        const int r_hi = m0 + IO_BM - 1 < p.row_end - 1 ? m0 + IO_BM - 1 : p.row_end - 1;
        const int c_hi = n0 + BN - 1 < p.b - 1 ? n0 + BN - 1 : p.b - 1;
        if (!ss_tile_needed(m0, r_hi, n0, c_hi, p.ss)) return;
    }
    extern __shared__ unsigned char ig_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)ig_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long s_full[IO_STAGES], s_empty[IO_STAGES], s_done;
    __shared__ unsigned s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < IO_STAGES; s++) { ig_mbar_init(ig_smem(&s_full[s]), 1); ig_mbar_init(ig_smem(&s_empty[s]), 1); }
        ig_mbar_init(ig_smem(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ig_smem(&s_tmem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < p.kblocks; kb++) {
                const int st = kb % IO_STAGES;
                ig_mbar_wait(ig_smem(&s_empty[st]), ((kb / IO_STAGES) & 1) ^ 1);
                const unsigned bar = ig_smem(&s_full[st]);
                ig_mbar_expect_tx(bar, IO_STAGE_BYTES);
                const unsigned base = ig_smem(tiles + (size_t)st * IO_STAGE_BYTES);
                for (int d = 0; d < IO_NP; d++) {
                    ig_tma_load_3d(base + d * IO_A_BYTES, &tmapA, bar, kb * IO_BK, m0, d);
                    ig_tma_load_3d(base + d * IO_A_BYTES + IO_A_BYTES / 2, &tmapA, bar, kb * IO_BK, m0 + 64, d);
                    ig_tma_load_3d(base + IO_NP * IO_A_BYTES + d * IO_B_BYTES, &tmapB, bar, kb * IO_BK, n0, d);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < p.kblocks; kb++) {
                const int st = kb % IO_STAGES;
                ig_mbar_wait(ig_smem(&s_full[st]), (kb / IO_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned base = ig_smem(tiles + (size_t)st * IO_STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < IO_BK / 32; kk++) {
#pragma unroll
                    for (int a = 0; a < IO_NP; a++) {
                        const unsigned long long ad = io_desc(base + a * IO_A_BYTES) + (unsigned long long)(kk * 32 >> 4);
                        // digit pairs with a + b <= NP - 1.  The B digit tiles of a stage are contiguous (64 rows x 64 B each)
                        // and so are the accumulators of adjacent scales: one MMA of N = 64 m multiplies A digit a with
                        // m <= IO_NGROUP adjacent B digits, so the A tile is read from shared memory once per group
                        // instead of once per pair (NP = 8: 123 KB instead of 221 KB per 32-deep k-step, below the
                        // 1152 tensor cycles x 128 B / cycle of shared-memory bandwidth)
#pragma unroll
                        for (int b0 = 0; b0 < IO_NP - a; b0 += IO_NGROUP) {
                            const int m = IO_NP - a - b0 < IO_NGROUP ? IO_NP - a - b0 : IO_NGROUP;
                            const unsigned idm = (2u << 4) | (1u << 7) | (1u << 10) | (((unsigned)(m * BN) >> 3) << 17) | ((IO_BM >> 4) << 24);
                            const unsigned long long bd = io_desc(base + IO_NP * IO_A_BYTES + b0 * IO_B_BYTES) + (unsigned long long)(kk * 32 >> 4);
                            const bool first = kb == 0 && kk == 0 && a == 0;      // a = 0 touches every accumulator
                            ig_mma_i8(tmem + (unsigned)(a + b0) * BN, ad, bd, idm, first ? 0u : 1u);
                        }
                    }
                }
                ig_commit(ig_smem(&s_empty[st]));
            }
            ig_commit(ig_smem(&s_done));
        }
    } else {
        ig_mbar_wait(ig_smem(&s_done), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lg = warp & 3;
        const int row = m0 + lg * 32 + lane;
        const bool rok = row < p.row_end && row < p.n;
        const double rs = rok ? p.alpha * p.rowscale[row] : 0.0;
        for (int c0 = 0; c0 < BN; c0 += 16) {
            unsigned r[IO_NP][16];
#pragma unroll
            for (int s = 0; s < IO_NP; s++) ig_tmem_ld16(tmem + ((unsigned)(lg * 32) << 16) + (unsigned)(s * BN + c0), r[s]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (rok) {
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const int col = n0 + c0 + c;
                    if (col >= p.b) continue;
                    double g = (double)(int)r[IO_NP - 1][c];
#pragma unroll
                    for (int s = IO_NP - 2; s >= 0; s--) g = fma(g, 0.0078125, (double)(int)r[s][c]);
                    double v = rs * p.colscale[col] * g;
                    if (p.E1) v += p.beta * p.E1[(size_t)row * p.lde1 + col];
                    if (p.E2) v += p.gamma * p.E2[(size_t)row * p.lde2 + col];
                    p.D[(size_t)row * p.ldd + col] = v;
                    // the digit sums are symmetric in (row, col) and the scales are powers of two: the mirrored
                    // element is the same bits
                    if (p.sym && col != row && col / p.ss.rpr == row / p.ss.rpr) p.D[(size_t)col * p.ldd + row] = v;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

static int io_encode(CUtensorMap *map, void *base, int rows_pad, int Kp, int planes, int box_rows = 64) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        if (!sym || qres != cudaDriverEntryPointSuccess) { tp_set_error("cuTensorMapEncodeTiled not available"); return TP_ERR_CUDA; }
        fn = (PFN_encodeTiled)sym;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows_pad, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * rows_pad};
    const cuuint32_t box[3] = {IO_BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tp_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TP_ERR_CUDA; }
    return TP_OK;
}

// digit planes of the symmetric operator S (n x n, ld); kept in the context until the next call
int tp_iop_prepare(tp_ctx *ctx, const double *S, int n, int ld) {
    cudaStream_t st = ctx->stream;
    const int rows_pad = round_up(n, IO_BM), Kp = round_up(n, 128);
    const size_t plane = (size_t)rows_pad * Kp;
    TP_TRY(ctx->ioA.reserve(IO_MAXNP * plane));
    TP_TRY(ctx->ioscale.reserve((size_t)(n + 1024) * (sizeof(double) + sizeof(int)) * 2 + 1024 * sizeof(unsigned long long) + 64));
    double *rowscale = ctx->ioscale.as<double>();
    int *rowexp = (int *)(rowscale + 2 * (n + 1024));
    tp_prof_begin(ctx, PC_ISLICE);
    io_rowmax_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(S, n, ld, rowexp, rowscale);
    const size_t chunks = (size_t)rows_pad * (Kp / 16);
    io_slice_rows_kernel<IO_MAXNP><<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(S, n, ld, rowexp, ctx->ioA.as<int8_t>(), rows_pad, Kp);
    tp_prof_end(ctx);
    ctx->launches += 2;
    TP_CUDA(cudaGetLastError());
    ctx->io_n = n;
    ctx->io_cmax_clean = false;      // the scratch of the column maxima sits behind the n-dependent scale arrays: clear it once
    return TP_OK;
}

// D[rows] = alpha S[rows, :] Yin + beta E1[rows] + gamma E2[rows], rows = [row_begin, row_end); NP = 5 (digit pairs
// left out below 2^-35 of row scale x column scale) or 8 (2^-56: FP64 level)
template <int NP, int BN>
static int iop_apply_np(tp_ctx *ctx, const double *Yin, int b, int ldy, double *D, int ldd, double alpha, const double *E1,
                        int lde1, double beta, const double *E2, int lde2, double gamma, int row_begin, int row_end) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->io_n;
    const int rows_padA = round_up(n, IO_BM), Kp = round_up(n, 128);
    const int rows_padB = round_up(b, IO_BN);
    const size_t planeB = (size_t)rows_padB * Kp;
    TP_TRY(ctx->ioB.reserve(IO_MAXNP * planeB));
    double *rowscale = ctx->ioscale.as<double>();
    double *colscale = rowscale + (n + 1024);
    int *rowexp = (int *)(rowscale + 2 * (n + 1024));
    int *colexp = rowexp + (n + 1024);
    tp_prof_begin(ctx, PC_ISLICE);
    unsigned long long *cmax = (unsigned long long *)(colexp + (n + 1024));      // 1024 maxima (zero between uses) + the ticket
    unsigned *ticket = (unsigned *)(cmax + 1024);
    if (!ctx->io_cmax_clean) {
        TP_CUDA(cudaMemsetAsync(cmax, 0, 1024 * sizeof(unsigned long long) + 16, st));
        ctx->io_cmax_clean = true;
    }
    const int cm_grid = std::min((n + IO_CM_ROWS - 1) / IO_CM_ROWS, 2 * ctx->sm_count);
    io_colmax_kernel<<<cm_grid, 256, 0, st>>>(Yin, n, b, ldy, cmax, ticket, colexp, colscale);
    dim3 sg(Kp / 128, rows_padB / 32);
    io_slice_cols_kernel<NP><<<sg, 256, 0, st>>>(Yin, n, b, ldy, colexp, ctx->ioB.as<int8_t>(), rows_padB, Kp);
    tp_prof_end(ctx);
    CUtensorMap mapA, mapB;
    TP_TRY(io_encode(&mapA, ctx->ioA.p, rows_padA, Kp, NP));        // the first NP of the IO_MAXNP planes
    TP_TRY(io_encode(&mapB, ctx->ioB.p, rows_padB, Kp, NP, BN));
    IoParams p;
    p.n = n; p.b = b; p.row_begin = row_begin; p.row_end = row_end; p.kblocks = Kp / IO_BK;
    p.D = D; p.ldd = ldd; p.E1 = E1; p.lde1 = lde1; p.E2 = E2; p.lde2 = lde2;
    p.alpha = alpha; p.beta = beta; p.gamma = gamma; p.rowscale = rowscale; p.colscale = colscale; p.sym = 0; p.banded = 0; p.ss = SymShard{1, 1 << 30};
    const size_t smem = (size_t)IoCfg<NP, BN>::STAGES * IoCfg<NP, BN>::STAGE_BYTES + 1024;
    TP_CUDA(tp_optin_smem(io_gemm_kernel<NP, BN>, ctx));
    dim3 grid(rows_padB / BN, (row_end - row_begin + IO_BM - 1) / IO_BM);
    if (ctx->prof) ctx->prof_imma_ops += 2.0 * (NP * (NP + 1) / 2) * (double)grid.x * grid.y * IO_BM * BN * (double)Kp;
    tp_prof_begin(ctx, PC_IGEMM);
    io_gemm_kernel<NP, BN><<<grid, IG_THREADS, smem, st>>>(mapA, mapB, p);
    tp_prof_end(ctx);
    ctx->launches += 3;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}

int tp_iop_apply(tp_ctx *ctx, const double *Yin, int b, int ldy, double *D, int ldd, double alpha, const double *E1,
                 int lde1, double beta, const double *E2, int lde2, double gamma, int row_begin, int row_end, int np) {
    TP_ARG(ctx->io_n > 0, "tp_iop_apply: tp_iop_prepare first");
    TP_ARG(b <= 1024, "tp_iop_apply: block wider than 1024 columns");
    TP_ARG(np == 5 || np == 8, "tp_iop_apply: 5 or 8 digit planes");
    // tiles of 64 columns while they fill the GPU; otherwise 32 (twice the CTAs, two per SM at 5 planes)
    const long tiles64 = (long)((row_end - row_begin + IO_BM - 1) / IO_BM) * ((b + 63) / 64);
    const bool narrow = ctx->io_bn32 >= 0 ? ctx->io_bn32 != 0 : tiles64 < ctx->sm_count;
    if (np == 5) return narrow ? iop_apply_np<5, 32>(ctx, Yin, b, ldy, D, ldd, alpha, E1, lde1, beta, E2, lde2, gamma, row_begin, row_end)
                               : iop_apply_np<5, 64>(ctx, Yin, b, ldy, D, ldd, alpha, E1, lde1, beta, E2, lde2, gamma, row_begin, row_end);
    return narrow ? iop_apply_np<8, 32>(ctx, Yin, b, ldy, D, ldd, alpha, E1, lde1, beta, E2, lde2, gamma, row_begin, row_end)
                  : iop_apply_np<8, 64>(ctx, Yin, b, ldy, D, ldd, alpha, E1, lde1, beta, E2, lde2, gamma, row_begin, row_end);
}

// M[rows] = A[rows, :] A^T for an FP64 matrix A (n x n, ld), rows = [row_begin, row_end): the sliced product above with
// both operands cut from the rows of A (one set of 8 digit planes, per-row exponents), FP64 level (digit pairs left out
// below 2^-56 of row scale x row scale; every product kept is exact, so the sum carries no accumulation rounding at all
// -- the FP64 DMMA Gram carries ~sqrt(n) ulp).  Used for M = Xc Xc^T of stage 3 (pca.cu).  With all rows requested only
// the tiles on and above the diagonal are computed (36 digit products x n^3 / 2 MACs) and mirrored; a row block (rank
// of a sharded call) computes its full width.  The planes share ctx->ioA with tp_iop_prepare, which runs after this.
int tp_igram_sliced(tp_ctx *ctx, const double *A, int n, int ld, double *M, int ldm, int row_begin, int row_end,
                    int sym, SymShard ss) {
    constexpr int NP = IO_MAXNP;
    cudaStream_t st = ctx->stream;
    const int rows_pad = round_up(n, IO_BM), Kp = round_up(n, 128);
    const size_t plane = (size_t)rows_pad * Kp;
    TP_TRY(ctx->ioA.reserve(IO_MAXNP * plane));
    TP_TRY(ctx->ioscale.reserve((size_t)(n + 1024) * (sizeof(double) + sizeof(int)) * 2 + 1024 * sizeof(unsigned long long) + 64));
    double *rowscale = ctx->ioscale.as<double>();
    int *rowexp = (int *)(rowscale + 2 * (n + 1024));
    tp_prof_begin(ctx, PC_ISLICE);
    io_rowmax_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(A, n, ld, rowexp, rowscale);
    const size_t chunks = (size_t)rows_pad * (Kp / 16);
    io_slice_rows_kernel<NP><<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(A, n, ld, rowexp, ctx->ioA.as<int8_t>(), rows_pad, Kp);
    tp_prof_end(ctx);
    ctx->launches += 2;
    if (row_end > row_begin) {
        CUtensorMap map;
        TP_TRY(io_encode(&map, ctx->ioA.p, rows_pad, Kp, NP));
        IoParams p;
        p.n = n; p.b = n; p.row_begin = row_begin; p.row_end = row_end; p.kblocks = Kp / IO_BK;
        p.D = M; p.ldd = ldm; p.E1 = nullptr; p.lde1 = 0; p.E2 = nullptr; p.lde2 = 0;
        p.alpha = 1.0; p.beta = 0.0; p.gamma = 0.0; p.rowscale = rowscale; p.colscale = rowscale;
        p.sym = sym; p.banded = 1; p.ss = ss;
        const size_t smem = (size_t)IoCfg<NP, IO_BN>::STAGES * IoCfg<NP, IO_BN>::STAGE_BYTES + 1024;
        TP_CUDA(tp_optin_smem(io_gemm_kernel<NP, IO_BN>, ctx));
        dim3 grid(rows_pad / IO_BN, (row_end - row_begin + IO_BM - 1) / IO_BM);
        if (ctx->prof) {
            double tiles = 0.0;
            for (int m0 = row_begin; m0 < row_end; m0 += IO_BM)
                for (int n0 = 0; n0 < n; n0 += IO_BN)
                    tiles += !sym || ss_tile_needed(m0, std::min(m0 + IO_BM, row_end) - 1, n0, std::min(n0 + IO_BN, n) - 1, ss);
            ctx->prof_imma_ops += 2.0 * (NP * (NP + 1) / 2) * tiles * IO_BM * IO_BN * (double)Kp;
        }
        tp_prof_begin(ctx, PC_IGEMM);
        io_gemm_kernel<NP, IO_BN><<<grid, IG_THREADS, smem, st>>>(map, map, p);
        tp_prof_end(ctx);
        ctx->launches += 1;
    }
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}

// ---- SymShard: fill the elements their row owner did not compute from the transposed ones (after the all-gather) -----------
__global__ void __launch_bounds__(256)
mirror_fill_kernel(double *__restrict__ D, int n, int ld, SymShard s) {
    __shared__ double t[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    // 32 x 32 tiles lie within one block pair (rpr is a multiple of 64): nothing to do inside diagonal blocks, and a tile
    // none of whose elements needs filling is skipped before the load
    if (i0 / s.rpr == j0 / s.rpr) return;
    const int i1 = i0 + 31 < n - 1 ? i0 + 31 : n - 1, j1 = j0 + 31 < n - 1 ? j0 + 31 : n - 1;
    if (ss_need(i0, j0, s) && ss_need(i0, j1, s) && ss_need(i1, j0, s) && ss_need(i1, j1, s)) return;   // (half-planes: corners decide)
    for (int r = ty; r < 32; r += 8) {          // t[r][c] = D[j0 + r][i0 + c]
        const int j = j0 + r, i = i0 + tx;
        t[r][tx] = (j < n && i < n) ? D[(size_t)j * ld + i] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        if (i < n && j < n && !ss_need(i, j, s)) D[(size_t)i * ld + j] = t[tx][r];
    }
}

int tp_mirror_fill(tp_ctx *ctx, double *D, int n, int ld, SymShard s) {
    if (s.R <= 1) return TP_OK;
    dim3 grid((n + 31) / 32, (n + 31) / 32);
    tp_prof_begin(ctx, PC_ISLICE);
    mirror_fill_kernel<<<grid, 256, 0, ctx->stream>>>(D, n, ld, s);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}
