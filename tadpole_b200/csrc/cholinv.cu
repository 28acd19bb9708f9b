// cholinv.cu -- Cholesky factor and triangular inverse of the b x b Gram matrices of stage 3 (CholQR between
// Chebyshev filter rounds, and the factor the one-sided Jacobi solver works on), b <= 256.
//
// One thread-block cluster of 8 CTAs; CTA j owns block column j (32 columns, rows 32 j .. b) in shared
// memory.  Right-looking: at step p the owner factors its 32 x 32 diagonal block in the registers of one
// warp (lane = row, column broadcast through shared memory), inverts it (lane = column, right-looking
// substitution), turns the rows below into L_ip = A_ip L_pp^-T as a dense product with that inverse, and
// publishes the panel through L2.  One hardware cluster barrier later every CTA j > p applies the rank-32
// update to its own column, so the next owner can start at once; no CTA ever waits for another one's
// update.  After the last panel CTA j computes block column j of L^-1:
// X_jj = L_jj^-1, X_ij = -L_ii^-1 * sum_{k=j}^{i-1} L_ik X_kj, all eight columns in parallel.
// The matrix is padded with the identity up to a multiple of 32 so every block is full.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CI_W 32
#define CI_P 33                  // shared-memory pitch of a 32-wide block (odd: conflict-free by row and by column)
#define CI_THREADS 512
#define CI_CLUSTER 8
#define CI_MAXB 256
#define CI_ROWS CI_MAXB

__device__ __forceinline__ double ci_rsqrt(double x) {        // MUFU seed + three Newton steps: full precision
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * fma(-0.5 * x * r, r, 1.5);
    r = r * fma(-0.5 * x * r, r, 1.5);
    return r * fma(-0.5 * x * r, r, 1.5);
}

// Cholesky of the 32 x 32 diagonal block in col (pitch CI_P), one warp, lane = row, left-looking:
// column c of every row is a[r][c] - sum_{t<c} L[r][t] L[c][t] (own row conflict-free, row c broadcast).
__device__ __forceinline__ void ci_diag_chol(double *col, double *s_rdiag, double clamp, int factor_only, int *info,
                                             int lane) {
    int bad = 0;
    const double *row = col + lane * CI_P;
    for (int c = 0; c < CI_W; c++) {
        const double *rc = col + c * CI_P;
        double v0 = row[c], v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int t = 0;
        for (; t + 4 <= c; t += 4) {
            v0 = fma(-row[t], rc[t], v0); v1 = fma(-row[t + 1], rc[t + 1], v1);
            v2 = fma(-row[t + 2], rc[t + 2], v2); v3 = fma(-row[t + 3], rc[t + 3], v3);
        }
        for (; t < c; t++) v0 = fma(-row[t], rc[t], v0);
        const double v = (v0 + v1) + (v2 + v3);
        double d = __shfl_sync(0xffffffffu, v, c);
        // factor_only: a pivot at round-off level (semi-definite input) keeps a tiny diagonal and a ZERO column,
        // so the null space cannot feed garbage into later columns
        const bool tiny = factor_only && !(d > clamp);
        if (tiny) d = clamp > 0.0 ? clamp : 1e-300;
        const bool ok = d > 0.0;
        bad |= !ok;
        const double r = ci_rsqrt(ok ? d : 1.0);
        __syncwarp();
        if (lane == c) { col[c * CI_P + c] = d * r; s_rdiag[c] = tiny ? 0.0 : r; }
        else if (lane > c) col[lane * CI_P + c] = tiny ? 0.0 : v * r;
        else col[lane * CI_P + c] = 0.0;                 // strictly upper part: explicit zeros
        __syncwarp();
    }
    if (bad && lane == 0) info[0] = 1;
}

// dinv = inverse of the lower-triangular 32 x 32 block Ld (pitch CI_P), one warp, lane = column:
// x[i] = (delta_i,lane - sum_{t<i} L[i][t] x[t]) / L[i][i]; explicit zeros above the diagonal.
__device__ __forceinline__ void ci_diag_inv(const double *Ld, const double *s_rdiag, double *dinv, int lane) {
    for (int i = 0; i < CI_W; i++) {
        const double *li = Ld + i * CI_P;
        double v0 = (i == lane) ? 1.0 : 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int t = 0;
        for (; t + 4 <= i; t += 4) {
            v0 = fma(-li[t], dinv[t * CI_P + lane], v0); v1 = fma(-li[t + 1], dinv[(t + 1) * CI_P + lane], v1);
            v2 = fma(-li[t + 2], dinv[(t + 2) * CI_P + lane], v2); v3 = fma(-li[t + 3], dinv[(t + 3) * CI_P + lane], v3);
        }
        for (; t < i; t++) v0 = fma(-li[t], dinv[t * CI_P + lane], v0);
        dinv[i * CI_P + lane] = (i >= lane) ? ((v0 + v1) + (v2 + v3)) * s_rdiag[i] : 0.0;
    }
}

__global__ void __cluster_dims__(CI_CLUSTER, 1, 1) __launch_bounds__(CI_THREADS, 1)
cholinv8_kernel(double *__restrict__ G, double *__restrict__ Linv, int b, int ld, int *info, int factor_only) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double ci_sm[];
    double *col = ci_sm;                          // own block column: rows x CI_P (rows relative to 32 j)
    double *pan = col + CI_ROWS * CI_P;           // panel rows 32 j .. of the current step / L row panel
    double *dinv = pan + CI_ROWS * CI_P;          // 32 x CI_P: inverse of a diagonal block
    double *sS = dinv + CI_W * CI_P;              // 32 x CI_P scratch (two halves of the k-split sum: 2 x)
    double *sS2 = sS + CI_W * CI_P;
    __shared__ double s_colc[2][CI_W];
    __shared__ double s_rdiag[CI_W];
    __shared__ double s_red[CI_THREADS / 32];
    __shared__ double s_clamp;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nb = (b + CI_W - 1) / CI_W;
    const int j = (int)cluster.block_rank();
    const int rows = j < nb ? (nb - j) * CI_W : 0;       // rows of the own column (padded)
    const int c0 = j * CI_W;

    // ---- load the own column; identity in the padding ------------------------------------------------
    for (int idx = tid; idx < rows * CI_W; idx += CI_THREADS) {
        const int r = idx / CI_W, c = idx % CI_W;
        const int gr = c0 + r, gc = c0 + c;
        col[r * CI_P + c] = (gr < b && gc < b) ? G[(size_t)gr * ld + gc] : (gr == gc ? 1.0 : 0.0);
    }
    {   // pivot clamp (factor_only: semi-definite input allowed): 1e-24 * max |diag|
        double mx = 0.0;
        if (factor_only) for (int i = tid; i < b; i += CI_THREADS) mx = fmax(mx, fabs(G[(size_t)i * ld + i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[wid] = mx;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < CI_THREADS / 32; w++) v = fmax(v, s_red[w]);
            s_clamp = v * 1e-24;
        }
        if (j == 0 && tid == 0) info[0] = 0;
    }
    cluster.sync();                      // nobody overwrites G before every CTA has read what it needs

    for (int p = 0; p < nb; p++) {
        if (j == p) {
            // ---- (1) Cholesky of the diagonal block: warp 0 ---------------------------------------------------
            if (wid == 0) ci_diag_chol(col, s_rdiag, s_clamp, factor_only, info, lane);
            __syncthreads();
            // ---- (2) rows below: L_ip = A_ip L_pp^-T by forward substitution, one thread per row, in place ----
            for (int r = CI_W + tid; r < rows; r += CI_THREADS) {
                double *row = col + r * CI_P;
                for (int c = 0; c < CI_W; c++) {
                    const double *rc = col + c * CI_P;
                    double v0 = row[c], v1 = 0.0, v2 = 0.0, v3 = 0.0;
                    int t = 0;
                    for (; t + 4 <= c; t += 4) {
                        v0 = fma(-row[t], rc[t], v0); v1 = fma(-row[t + 1], rc[t + 1], v1);
                        v2 = fma(-row[t + 2], rc[t + 2], v2); v3 = fma(-row[t + 3], rc[t + 3], v3);
                    }
                    for (; t < c; t++) v0 = fma(-row[t], rc[t], v0);
                    row[c] = ((v0 + v1) + (v2 + v3)) * s_rdiag[c];
                }
            }
            __syncthreads();
            // ---- (4) publish: factor to G (lower triangle), inverse of the diagonal block to Linv ----------
            for (int idx = tid; idx < rows * CI_W; idx += CI_THREADS) {
                const int r = idx / CI_W, c = idx % CI_W;
                const int gr = c0 + r, gc = c0 + c;
                if (gr < b && gc < b && gr >= gc) G[(size_t)gr * ld + gc] = col[r * CI_P + c];
            }
        }
        cluster.sync();                  // panel p is in L2 (release / acquire at cluster scope)
        if (j == p && !factor_only) {
            // off the critical path: inverse of the diagonal block, published for the inverse phase
            if (wid == 0) ci_diag_inv(col, s_rdiag, dinv, lane);
            __syncthreads();
            for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {
                const int r = idx / CI_W, c = idx % CI_W;
                const int gr = c0 + r, gc = c0 + c;
                if (gr < b && gc < b) Linv[(size_t)gr * ld + gc] = dinv[r * CI_P + c];
            }
        }
        if (p == nb - 1) break;
        if (j > p && j < nb) {
            // ---- rank-32 update of the own column with panel rows 32 j .. -----------------------------------
            const int pc0 = p * CI_W;
            for (int idx = tid; idx < rows * CI_W; idx += CI_THREADS) {
                const int r = idx / CI_W, t = idx % CI_W;
                const int gr = c0 + r, gc = pc0 + t;
                pan[r * CI_P + t] = (gr < b) ? __ldcg(&G[(size_t)gr * ld + gc]) : 0.0;
            }
            __syncthreads();
            const int ngroups = rows / CI_W;
            for (int task = wid; task < ngroups * 2; task += CI_THREADS / 32) {
                const int g = task >> 1, ch = (task & 1) * 16;
                const int r = g * CI_W + lane;
                double acc[16];
#pragma unroll
                for (int c = 0; c < 16; c++) acc[c] = col[r * CI_P + ch + c];
#pragma unroll 2
                for (int t = 0; t < CI_W; t++) {
                    const double av = pan[r * CI_P + t];
#pragma unroll
                    for (int c = 0; c < 16; c++) acc[c] = fma(-av, pan[(ch + c) * CI_P + t], acc[c]);
                }
#pragma unroll
                for (int c = 0; c < 16; c++) col[r * CI_P + ch + c] = acc[c];
            }
            __syncthreads();
        }
    }
    if (factor_only) return;
    cluster.sync();                      // every diagonal-block inverse is in L2
    if (j >= nb || __ldcg(&info[0])) return;        // not positive definite: Linv is not used

    // ---- block column j of L^-1 ----------------------------------------------------------------------------
    // col is reused as X (rows relative to 32 j); dinv still holds L_jj^-1 from step j
    for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {
        const int r = idx / CI_W, c = idx % CI_W;
        col[r * CI_P + c] = dinv[r * CI_P + c];
    }
    for (int idx = tid; idx < c0 * CI_W; idx += CI_THREADS) {      // rows above the diagonal block are zero
        const int r = idx / CI_W, c = idx % CI_W;
        if (c0 + c < b) Linv[(size_t)r * ld + c0 + c] = 0.0;
    }
    __syncthreads();
    for (int i = j + 1; i < nb; i++) {
        const int K = (i - j) * CI_W;
        // L row panel, transposed: pan[t * CI_P + r] = L[32 i + r][32 j + t]; and L_ii^-1
        for (int idx = tid; idx < CI_W * K; idx += CI_THREADS) {
            const int r = idx / K, t = idx % K;
            const int gr = i * CI_W + r, gc = c0 + t;
            pan[t * CI_P + r] = (gr < b && gc < b) ? __ldcg(&G[(size_t)gr * ld + gc]) : 0.0;
        }
        for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {
            const int r = idx / CI_W, c = idx % CI_W;
            const int gr = i * CI_W + r, gc = i * CI_W + c;
            dinv[r * CI_P + c] = (gr < b && gc < b) ? __ldcg(&Linv[(size_t)gr * ld + gc]) : (r == c ? 1.0 : 0.0);
        }
        __syncthreads();
        {   // S = L_i,j..i-1 X_j..i-1,j: lane = column; a warp owns 4 rows and one half of the k range
            const int r0 = (wid & 7) * 4, kh = wid >> 3;
            const int t0 = kh * (K / 2), t1 = t0 + K / 2;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
            for (int t = t0; t < t1; t++) {
                const double xv = col[t * CI_P + lane];
                const double *lp = pan + t * CI_P + r0;
                a0 = fma(lp[0], xv, a0); a1 = fma(lp[1], xv, a1); a2 = fma(lp[2], xv, a2); a3 = fma(lp[3], xv, a3);
            }
            double *dst = kh ? sS2 : sS;
            dst[(r0 + 0) * CI_P + lane] = a0; dst[(r0 + 1) * CI_P + lane] = a1;
            dst[(r0 + 2) * CI_P + lane] = a2; dst[(r0 + 3) * CI_P + lane] = a3;
        }
        __syncthreads();
        for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {      // X_ij = -L_ii^-1 S
            const int r = idx / CI_W, c = idx % CI_W;
            double acc = 0.0;
            for (int t = 0; t <= r; t++) acc = fma(dinv[r * CI_P + t], sS[t * CI_P + c] + sS2[t * CI_P + c], acc);
            const double v = -acc;
            col[(K + r) * CI_P + c] = v;
            const int gr = i * CI_W + r, gc = c0 + c;
            if (gr < b && gc < b) Linv[(size_t)gr * ld + gc] = v;
        }
        __syncthreads();
    }
}

int tp_chol_inv_1cta(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *info_dev, int factor_only);

static int launch_cholinv(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *info, int factor_only) {
    if (b > CI_MAXB) return tp_chol_inv_1cta(ctx, G, Linv, b, ld, info, factor_only);
    const size_t smem = (size_t)(2 * CI_ROWS * CI_P + 3 * CI_W * CI_P) * sizeof(double);
    TP_CUDA(cudaFuncSetAttribute(cholinv8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tp_prof_begin(ctx, PC_CHOL);
    cholinv8_kernel<<<CI_CLUSTER, CI_THREADS, smem, ctx->stream>>>(G, Linv, b, ld, info, factor_only);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}

// G (b x b, ld) is overwritten by its Cholesky factor (lower triangle); Linv receives L^-1.  *bad_out = 1 when G
// is not numerically positive definite.
int tp_chol_inv(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *bad_out) {
    TP_TRY(ctx->harm.reserve(64));
    int *info = ctx->harm.as<int>();
    TP_TRY(launch_cholinv(ctx, G, Linv, b, ld, info, 0));
    TP_TRY(tp_pin_reserve(ctx, 64));
    int *h = (int *)ctx->pin;
    TP_CUDA(cudaMemcpyAsync(h, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TP_CUDA(cudaStreamSynchronize(ctx->stream));
    *bad_out = h[0];
    return TP_OK;
}

// G <- Cholesky factor (lower triangle), semi-definite input tolerated (pivots clamped); no read-back
int tp_chol_factor(tp_ctx *ctx, double *G, int b, int ld) {
    TP_TRY(ctx->harm.reserve(64));
    return launch_cholinv(ctx, G, nullptr, b, ld, ctx->harm.as<int>(), 1);
}
