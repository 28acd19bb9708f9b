// gemm.cu -- FP64 DMMA GEMM (see gemm.cuh).
#include "gemm.cuh"
#include <stdlib.h>

#ifndef GEMM_BK
#define GEMM_BK 16
#endif
#define GEMM_STAGES 3
#define PITCH_K (GEMM_BK + 4)   // operand stored [row][k]: pitch 20 doubles -> conflict-free fragment reads

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// One operand tile: ROWS (m or n) x BK.  KC: global (row, k) at base[row*ld + k], smem [row][PITCH_K].
// !KC: global at base[k*ld + row], smem [k][ROWS+4].
template <int ROWS, bool KC, int NT>
__device__ __forceinline__ void load_tile(double *smem, const double *base, long ld, int row0, int nrows_total,
                                          int k0, int k_end, int tid) {
    if (KC) {
        constexpr int CH = ROWS * (GEMM_BK / 2);
#pragma unroll
        for (int ch = tid; ch < CH; ch += NT) {
            const int r = ch / (GEMM_BK / 2), kc = (ch % (GEMM_BK / 2)) * 2;
            const int gr = row0 + r, gk = k0 + kc;
            int valid = (gr < nrows_total) ? min(max(k_end - gk, 0), 2) : 0;
            const double *src = valid ? base + (size_t)gr * ld + gk : base;
            cp_async16(smem + r * PITCH_K + kc, src, valid * 8);
        }
    } else {
        constexpr int PM = ROWS + 4;
        constexpr int CH = GEMM_BK * (ROWS / 2);
#pragma unroll
        for (int ch = tid; ch < CH; ch += NT) {
            const int k = ch / (ROWS / 2), rc = (ch % (ROWS / 2)) * 2;
            const int gk = k0 + k, gr = row0 + rc;
            int valid = (gk < k_end) ? min(max(nrows_total - gr, 0), 2) : 0;
            const double *src = valid ? base + (size_t)gk * ld + gr : base;
            cp_async16(smem + k * PM + rc, src, valid * 8);
        }
    }
}

template <int ROWS, bool KC>
__device__ __forceinline__ double frag(const double *smem, int row, int k) {
    return KC ? smem[row * PITCH_K + k] : smem[k * (ROWS + 4) + row];
}

template <int ROWS, bool KC>
__host__ __device__ constexpr int tile_doubles() { return KC ? ROWS * PITCH_K : GEMM_BK * (ROWS + 4); }

template <bool A_KC, bool B_KC, int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__(WM *WN * 32)
dgemm_kernel(GemmArgs g, double *__restrict__ partial, int kt_per_split) {
    constexpr int NT = WM * WN * 32;
    constexpr int WTM = BM / WM, WTN = BN / WN;
    constexpr int MI = WTM / 8, NI = WTN / 8;
    constexpr int A_SZ = tile_doubles<BM, A_KC>(), B_SZ = tile_doubles<BN, B_KC>();
    extern __shared__ __align__(16) double gsm[];

    const int tm = blockIdx.y, tn = blockIdx.x;
    const int m0 = tm * BM, n0 = tn * BN;
    if (g.sym && n0 + BN <= m0) return;          // tile entirely below the diagonal: its mirror image is computed
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / WN) * WTM, wn0 = (warp % WN) * WTN;

    const int KT = (g.K + GEMM_BK - 1) / GEMM_BK;
    const int kt0 = blockIdx.z * kt_per_split;
    const int kt1 = min(KT, kt0 + kt_per_split);
    const int nkt = max(kt1 - kt0, 0);

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto stage_a = [&](int s) { return gsm + (size_t)s * (A_SZ + B_SZ); };
    auto stage_b = [&](int s) { return gsm + (size_t)s * (A_SZ + B_SZ) + A_SZ; };

#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; s++) {
        if (s < nkt) {
            load_tile<BM, A_KC, NT>(stage_a(s), g.A, g.lda, m0, g.M, (kt0 + s) * GEMM_BK, g.K, tid);
            load_tile<BN, B_KC, NT>(stage_b(s), g.B, g.ldb, n0, g.N, (kt0 + s) * GEMM_BK, g.K, tid);
        }
        cp_async_commit();
    }
    for (int it = 0; it < nkt; it++) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        const int nx = it + GEMM_STAGES - 1;
        if (nx < nkt) {
            load_tile<BM, A_KC, NT>(stage_a(nx % GEMM_STAGES), g.A, g.lda, m0, g.M, (kt0 + nx) * GEMM_BK, g.K, tid);
            load_tile<BN, B_KC, NT>(stage_b(nx % GEMM_STAGES), g.B, g.ldb, n0, g.N, (kt0 + nx) * GEMM_BK, g.K, tid);
        }
        cp_async_commit();
        const double *As = stage_a(it % GEMM_STAGES), *Bs = stage_b(it % GEMM_STAGES);
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; kk += 4) {
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; i++) af[i] = frag<BM, A_KC>(As, wm0 + i * 8 + (lane >> 2), kk + (lane & 3));
#pragma unroll
            for (int j = 0; j < NI; j++) bf[j] = frag<BN, B_KC>(Bs, wn0 + j * 8 + (lane >> 2), kk + (lane & 3));
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue -----------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < MI; i++) {
        const int m = m0 + wm0 + i * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < NI; j++) {
            const int n = n0 + wn0 + j * 8 + (lane & 3) * 2;
            if (m >= g.M) continue;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int nn = n + e;
                if (nn >= g.N) continue;
                double v = acc[i][j][e];
                if (partial) {
                    partial[((size_t)blockIdx.z * g.M + m) * g.N + nn] = v;
                    continue;
                }
                if (g.epi == EPI_CORR) {
                    // (crossprod - nrow * tcrossprod(colMeans)) / (nrow - 1), then / tcrossprod(sd); NaN -> 0
                    v = (v - g.nrows * (g.mean[m + g.epi_row0] * g.mean[nn])) / (g.nrows - 1.0);
                    v = v / (g.sd[m + g.epi_row0] * g.sd[nn]);
                    v = nan_to_zero(v);
                } else {
                    v *= g.alpha;
                    if (g.E1) v += g.beta * g.E1[(size_t)m * g.lde1 + nn];
                    if (g.E2) v += g.gamma * g.E2[(size_t)m * g.lde2 + nn];
                }
                g.D[(size_t)m * g.ldd + nn] = v;
                if (g.sym && nn != m && nn < g.M && m < g.N) g.D[(size_t)nn * g.ldd + m] = v;
            }
        }
    }
}

__global__ void splitk_reduce_kernel(GemmArgs g, const double *__restrict__ partial, int splits) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)g.M * g.N;
    if (idx >= total) return;
    const int m = (int)(idx / g.N), n = (int)(idx % g.N);
    double v = 0.0;
    for (int z = 0; z < splits; z++) v += partial[(size_t)z * total + idx];
    v *= g.alpha;
    if (g.E1) v += g.beta * g.E1[(size_t)m * g.lde1 + n];
    if (g.E2) v += g.gamma * g.E2[(size_t)m * g.lde2 + n];
    g.D[(size_t)m * g.ldd + n] = v;
}

static inline bool blockIdx_z_first(int) { return true; }

template <bool A_KC, bool B_KC, int BM, int BN, int WM, int WN>
static int launch_cfg(tp_ctx *ctx, const GemmArgs &g, double *partial, int splits, int kt_per_split) {
    constexpr size_t smem = (size_t)GEMM_STAGES * (tile_doubles<BM, A_KC>() + tile_doubles<BN, B_KC>()) * sizeof(double);
    auto kern = dgemm_kernel<A_KC, B_KC, BM, BN, WM, WN>;
    TP_CUDA(tp_optin_smem(kern, ctx));
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splits);
    tp_prof_begin(ctx, PC_GEMM);
    // algorithmic flops: 2 M N K, or M N (K + 1) ~ SYRK count when only one triangle is computed
    if (ctx->prof && blockIdx_z_first(splits)) ctx->prof_gemm_flop += g.sym ? (double)g.M * g.N * g.K : 2.0 * g.M * g.N * g.K;
    kern<<<grid, WM * WN * 32, smem, ctx->stream>>>(g, partial, kt_per_split);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}

template <bool A_KC, bool B_KC>
static int launch_layout(tp_ctx *ctx, const GemmArgs &g, int cfg, double *partial, int splits, int ktp) {
    if (cfg == 2) return launch_cfg<A_KC, B_KC, 128, 128, 2, 4>(ctx, g, partial, splits, ktp);
    if (cfg == 1) return launch_cfg<A_KC, B_KC, 64, 64, 2, 2>(ctx, g, partial, splits, ktp);
    return launch_cfg<A_KC, B_KC, 32, 64, 1, 4>(ctx, g, partial, splits, ktp);
}

int tp_gemm(tp_ctx *ctx, const GemmArgs &g) {
    TP_ARG(g.A && g.B && g.D && g.M > 0 && g.N > 0 && g.K > 0, "tp_gemm: bad arguments");
    TP_ARG((g.lda % 2) == 0 && (g.ldb % 2) == 0, "tp_gemm: leading dimensions must be even");
    // 128x128 tiles once they fill the machine; otherwise 64x64 or 32x64, whichever leaves the 148 SMs
    // better balanced (time ~ ceil(tiles / SMs) * tile area when every SM holds its tiles at once)
    const long tiles128 = (long)((g.M + 127) / 128) * ((g.N + 127) / 128) / (g.sym ? 2 : 1);
    int cfg = 2;
    if (tiles128 < ctx->sm_count) {
        const long z = g.splitk > 1 ? g.splitk : 1;
        const long t64 = (long)((g.M + 63) / 64) * ((g.N + 63) / 64) * z / (g.sym ? 2 : 1);
        const long t32 = (long)((g.M + 31) / 32) * ((g.N + 63) / 64) * z / (g.sym ? 2 : 1);
        const long c64 = ((t64 + ctx->sm_count - 1) / ctx->sm_count) * 2;
        const long c32 = ((t32 + ctx->sm_count - 1) / ctx->sm_count) * 1;
        cfg = (c32 <= c64) ? 0 : 1;      // tie: the smaller tile puts two CTAs (8 warps) on an SM
    }
    if (const char *e = getenv("TADPOLE_GEMM_CFG")) { if (cfg != 2 || atoi(e) >= 10) cfg = atoi(e) % 10; }   // tuning aid
    const int KT = (g.K + GEMM_BK - 1) / GEMM_BK;
    int splits = g.splitk;
    if (splits > KT) splits = KT;
    if (splits < 1) splits = 1;
    double *partial = nullptr;
    int ktp = KT;
    if (splits > 1) {
        TP_ARG(g.epi == EPI_LINEAR && !g.sym, "tp_gemm: split-K supports the linear epilogue only");
        ktp = (KT + splits - 1) / splits;
        splits = (KT + ktp - 1) / ktp;
        TP_TRY(ctx->part.reserve((size_t)splits * g.M * g.N * sizeof(double)));
        partial = ctx->part.as<double>();
    }
    int rc;
    if (g.a_kc) rc = g.b_kc ? launch_layout<true, true>(ctx, g, cfg, partial, splits, ktp)
                            : launch_layout<true, false>(ctx, g, cfg, partial, splits, ktp);
    else        rc = g.b_kc ? launch_layout<false, true>(ctx, g, cfg, partial, splits, ktp)
                            : launch_layout<false, false>(ctx, g, cfg, partial, splits, ktp);
    TP_TRY(rc);
    if (splits > 1) {
        const size_t total = (size_t)g.M * g.N;
        splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(g, partial, splits);
        ctx->launches += 1;
        TP_CUDA(cudaGetLastError());
    }
    return TP_OK;
}
