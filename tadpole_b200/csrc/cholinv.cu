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
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CI_W 32
#define CI_P 33                  // shared-memory pitch of a 32-wide block (odd: conflict-free by row and by column)
#define CI_PE 34                 // even pitch ([t][r] panels read as 16-byte broadcasts)
#define CI_THREADS 256
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

// Cholesky of the 32 x 32 diagonal block in col (pitch CI_P): one warp, lane = row, the row in registers,
// right-looking: per column one shuffle, one reciprocal square root and one broadcast of the finished column through
// shared memory; the trailing update of a lane is 31 - c independent FMAs.  Writes L to col (zeros above the diagonal),
// its transpose to dT (pitch 32: row c holds column c of L) and the reciprocal pivots to s_rdiag.
__device__ __forceinline__ void ci_diag_chol(double *col, double *dT, double (*s_colc)[CI_W], double *s_rdiag,
                                             double clamp, int factor_only, int *info, int lane) {
    double a[CI_W];
#pragma unroll
    for (int c = 0; c < CI_W; c++) a[c] = col[lane * CI_P + c];
    int bad = 0;
#pragma unroll
    for (int c = 0; c < CI_W; c++) {
        double d = __shfl_sync(0xffffffffu, a[c], c);
        // factor_only: a pivot at round-off level (semi-definite input) keeps a tiny diagonal and a ZERO column,
        // so the null space cannot feed garbage into later columns
        const bool tiny = factor_only && !(d > clamp);
        if (tiny) d = clamp > 0.0 ? clamp : 1e-300;
        const bool ok = d > 0.0;
        bad |= !ok;
        const double r = ci_rsqrt(ok ? d : 1.0);
        const double lc = (lane == c) ? d * r : ((lane > c && !tiny) ? a[c] * r : 0.0);
        a[c] = lc;
        s_colc[c & 1][lane] = lc;
        if (lane == c) s_rdiag[c] = tiny ? 0.0 : r;
        __syncwarp();
#pragma unroll
        for (int c2 = c + 1; c2 < CI_W; c2++) a[c2] = fma(-lc, s_colc[c & 1][c2], a[c2]);   // rows < c2: unused values
    }
#pragma unroll
    for (int c = 0; c < CI_W; c++) {
        const double v = (c <= lane) ? a[c] : 0.0;
        col[lane * CI_P + c] = v;
        dT[c * CI_W + lane] = v;
    }
    if (bad && lane == 0) info[0] = 1;
}

// dinv = inverse of the lower-triangular diagonal block, one warp, lane = column, right-looking substitution on the
// pending right-hand side in registers; dT[c][i] = L[i][c] (pitch 32).  Explicit zeros above the diagonal.
__device__ __forceinline__ void ci_diag_inv(const double *dT, const double *s_rdiag, double *dinv, int lane) {
    double s[CI_W];
#pragma unroll
    for (int i = 0; i < CI_W; i++) s[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int jx = 0; jx < CI_W; jx++) {
        const double xj = s[jx] * s_rdiag[jx];
        dinv[jx * CI_P + lane] = xj;
#pragma unroll
        for (int i = jx + 1; i < CI_W; i++) s[i] = fma(-dT[jx * CI_W + i], xj, s[i]);
    }
}

__global__ void __cluster_dims__(CI_CLUSTER, 1, 1) __launch_bounds__(CI_THREADS, 1)
cholinv8_kernel(double *__restrict__ G, double *__restrict__ Linv, int b, int ld, int *info, int factor_only,
                long long *trace) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double ci_sm[];
    double *col = ci_sm;                          // own block column: rows x CI_P (rows relative to 32 j); later X
    double *pan = col + CI_ROWS * CI_P;           // panel rows 32 j .. of the current step; later the L row panel [t][r]
    double *dinv = pan + CI_ROWS * CI_PE;         // 32 x CI_P: inverse of a diagonal block
    double *sS = dinv + CI_W * CI_P;              // 32 x CI_P scratch
    double *dT = sS + CI_W * CI_P;                // 32 x 32: transposed diagonal block / transposed panel rows of block j
    __shared__ double s_colc[2][CI_W];
    __shared__ double s_rdiag[CI_W];
    __shared__ double s_red[CI_THREADS / 32];
    __shared__ double s_clamp;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nb = (b + CI_W - 1) / CI_W;
    const int j = (int)cluster.block_rank();
    const int rows = j < nb ? (nb - j) * CI_W : 0;       // rows of the own column (padded)
    const int c0 = j * CI_W;
    // optional phase trace (clock64 per CTA): trace[(j * 16 + step) * 8 + phase]
#define CI_TR(step, ph) do { if (trace && tid == 0) trace[((size_t)j * 16 + (step)) * 8 + (ph)] = clock64(); } while (0)

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
    }
    cluster.sync();                      // nobody overwrites G before every CTA has read what it needs

    for (int p = 0; p < nb; p++) {
        CI_TR(p, 0);
        if (j == p) {
            // ---- (1) Cholesky of the diagonal block: warp 0 ---------------------------------------------------
            if (wid == 0) ci_diag_chol(col, dT, s_colc, s_rdiag, s_clamp, factor_only, info, lane);
            __syncthreads();
            CI_TR(p, 1);
            // ---- (2) rows below: L_ip = A_ip L_pp^-T, one thread per row, the row in registers, right-looking:
            //      x[c] /= L[c][c], then x[c2] -= x[c] L[c2][c] for c2 > c (column c of L broadcast from dT) ----------
            for (int r = CI_W + tid; r < rows; r += CI_THREADS) {
                double x[CI_W];
#pragma unroll
                for (int c = 0; c < CI_W; c++) x[c] = col[r * CI_P + c];
#pragma unroll
                for (int c = 0; c < CI_W; c++) {
                    x[c] *= s_rdiag[c];
#pragma unroll
                    for (int c2 = c + 1; c2 < CI_W; c2++) x[c2] = fma(-x[c], dT[c * CI_W + c2], x[c2]);
                }
#pragma unroll
                for (int c = 0; c < CI_W; c++) col[r * CI_P + c] = x[c];
            }
            __syncthreads();
            CI_TR(p, 2);
            // ---- (3) publish: factor to G (lower triangle); the off-diagonal rows also transposed into the upper
            //      triangle of Linv (scratch for the inverse phase, cleared at the end) ------------------------------
            for (int idx = tid; idx < rows * CI_W; idx += CI_THREADS) {
                const int r = idx / CI_W, c = idx % CI_W;
                const int gr = c0 + r, gc = c0 + c;
                if (gr < b && gc < b && gr >= gc) G[(size_t)gr * ld + gc] = col[r * CI_P + c];
            }
            if (!factor_only) {
                for (int idx = tid; idx < (rows - CI_W) * CI_W; idx += CI_THREADS) {
                    const int c = idx / (rows - CI_W), r = CI_W + idx % (rows - CI_W);
                    const int gr = c0 + r, gc = c0 + c;
                    if (gr < b && gc < b) Linv[(size_t)gc * ld + gr] = col[r * CI_P + c];
                }
            }
        }
        CI_TR(p, 3);
        cluster.sync();                  // panel p is in L2 (release / acquire at cluster scope)
        CI_TR(p, 4);
        if (j == p && !factor_only) {
            // off the critical path: inverse of the diagonal block, published for the inverse phase
            if (wid == 0) ci_diag_inv(dT, s_rdiag, dinv, lane);
            __syncthreads();
            for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {
                const int r = idx / CI_W, c = idx % CI_W;
                const int gr = c0 + r, gc = c0 + c;
                if (gr < b && gc < b) Linv[(size_t)gr * ld + gc] = dinv[r * CI_P + c];
            }
        }
        CI_TR(p, 5);
        if (p == nb - 1) break;
        if (j > p && j < nb) {
            // ---- rank-32 update of the own column with panel rows 32 j .. -----------------------------------
            const int pc0 = p * CI_W;
            for (int idx = tid; idx < rows * CI_W; idx += CI_THREADS) {
                const int r = idx / CI_W, t = idx % CI_W;
                const int gr = c0 + r, gc = pc0 + t;
                const double v = (gr < b) ? __ldcg(&G[(size_t)gr * ld + gc]) : 0.0;
                pan[r * CI_P + t] = v;
                if (r < CI_W) dT[t * CI_W + r] = v;            // transposed rows of block j: broadcast operand
            }
            __syncthreads();
            CI_TR(p, 6);
            const int ngroups = rows / CI_W;
            for (int task = wid; task < ngroups * 2; task += CI_THREADS / 32) {
                const int g = task >> 1, ch = (task & 1) * 16;
                const int r = g * CI_W + lane;
                double acc[16];
#pragma unroll
                for (int c = 0; c < 16; c++) acc[c] = col[r * CI_P + ch + c];
#pragma unroll 4
                for (int t = 0; t < CI_W; t++) {
                    const double av = pan[r * CI_P + t];
                    const double2 *bp = reinterpret_cast<const double2 *>(dT + t * CI_W + ch);
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const double2 bv = bp[c];
                        acc[2 * c] = fma(-av, bv.x, acc[2 * c]);
                        acc[2 * c + 1] = fma(-av, bv.y, acc[2 * c + 1]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 16; c++) col[r * CI_P + ch + c] = acc[c];
            }
            __syncthreads();
            CI_TR(p, 7);
        }
    }
    if (factor_only) return;
    cluster.sync();                      // every diagonal-block inverse is in L2
    if (j < nb && !__ldcg(&info[0])) {   // (not positive definite: Linv is not used)
        // ---- block column j of L^-1: col is reused as X (rows relative to 32 j); dinv still holds L_jj^-1 ----------
        for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {
            const int r = idx / CI_W, c = idx % CI_W;
            col[r * CI_P + c] = dinv[r * CI_P + c];
        }
        __syncthreads();
        CI_TR(8, 0);
        for (int i = j + 1; i < nb; i++) {
            const int K = (i - j) * CI_W;
            CI_TR(8 + i, 0);
            // L row panel from its transposed copy: pan[t * CI_PE + r] = L[32 i + r][32 j + t]; and L_ii^-1
            for (int idx = tid; idx < CI_W * K; idx += CI_THREADS) {
                const int t = idx / CI_W, r = idx % CI_W;
                const int gr = i * CI_W + r, gc = c0 + t;
                pan[t * CI_PE + r] = (gr < b && gc < b) ? __ldcg(&Linv[(size_t)gc * ld + gr]) : 0.0;
            }
            for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {
                const int r = idx / CI_W, c = idx % CI_W;
                const int gr = i * CI_W + r, gc = i * CI_W + c;
                dinv[r * CI_P + c] = (gr < b && gc < b) ? __ldcg(&Linv[(size_t)gr * ld + gc]) : (r == c ? 1.0 : 0.0);
            }
            __syncthreads();
            CI_TR(8 + i, 1);
            {   // S = L_i,j..i-1 X_j..i-1,j: lane = column; a warp owns 4 rows
                const int r0 = wid * 4;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
                for (int t = 0; t < K; t++) {
                    const double xv = col[t * CI_P + lane];
                    const double2 l01 = *reinterpret_cast<const double2 *>(pan + t * CI_PE + r0);
                    const double2 l23 = *reinterpret_cast<const double2 *>(pan + t * CI_PE + r0 + 2);
                    a0 = fma(l01.x, xv, a0); a1 = fma(l01.y, xv, a1); a2 = fma(l23.x, xv, a2); a3 = fma(l23.y, xv, a3);
                }
                sS[(r0 + 0) * CI_P + lane] = a0; sS[(r0 + 1) * CI_P + lane] = a1;
                sS[(r0 + 2) * CI_P + lane] = a2; sS[(r0 + 3) * CI_P + lane] = a3;
            }
            __syncthreads();
            CI_TR(8 + i, 2);
            for (int idx = tid; idx < CI_W * CI_W; idx += CI_THREADS) {      // X_ij = -L_ii^-1 S
                const int r = idx / CI_W, c = idx % CI_W;
                double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
                int t = 0;
                for (; t + 4 <= r + 1; t += 4) {
                    v0 = fma(dinv[r * CI_P + t], sS[t * CI_P + c], v0);
                    v1 = fma(dinv[r * CI_P + t + 1], sS[(t + 1) * CI_P + c], v1);
                    v2 = fma(dinv[r * CI_P + t + 2], sS[(t + 2) * CI_P + c], v2);
                    v3 = fma(dinv[r * CI_P + t + 3], sS[(t + 3) * CI_P + c], v3);
                }
                for (; t <= r; t++) v0 = fma(dinv[r * CI_P + t], sS[t * CI_P + c], v0);
                const double v = -((v0 + v1) + (v2 + v3));
                col[(K + r) * CI_P + c] = v;
                const int gr = i * CI_W + r, gc = c0 + c;
                if (gr < b && gc < b) Linv[(size_t)gr * ld + gc] = v;
            }
            __syncthreads();
            CI_TR(8 + i, 3);
        }
    }
    cluster.sync();                      // every CTA is done reading the transposed panels in the upper triangle
    for (int idx = tid; idx < c0 * CI_W; idx += CI_THREADS) {      // rows above the diagonal block are zero
        const int r = idx / CI_W, c = idx % CI_W;
        if (c0 + c < b) Linv[(size_t)r * ld + c0 + c] = 0.0;
    }
}

int tp_chol_inv_1cta(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *info_dev, int factor_only);

static int launch_cholinv(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *info, int factor_only) {
    if (b > CI_MAXB) return tp_chol_inv_1cta(ctx, G, Linv, b, ld, info, factor_only);
    const size_t smem = (size_t)(CI_ROWS * CI_P + CI_ROWS * CI_PE + 2 * CI_W * CI_P + CI_W * CI_W) * sizeof(double);
    TP_CUDA(tp_optin_smem(cholinv8_kernel, ctx));
    tp_prof_begin(ctx, PC_CHOL);
    long long *trace = nullptr;
    const bool tr = getenv("TADPOLE_CHOL_TRACE") != nullptr;
    DevBuf trbuf;
    if (tr) {
        TP_TRY(trbuf.reserve(CI_CLUSTER * 16 * 8 * sizeof(long long)));
        trace = trbuf.as<long long>();
        TP_CUDA(cudaMemsetAsync(trace, 0, CI_CLUSTER * 16 * 8 * sizeof(long long), ctx->stream));
    }
    cholinv8_kernel<<<CI_CLUSTER, CI_THREADS, smem, ctx->stream>>>(G, Linv, b, ld, info, factor_only, trace);
    tp_prof_end(ctx);
    if (tr) {       // debugging aid: per-CTA phase clocks relative to the earliest one
        static long long h[CI_CLUSTER * 16 * 8];
        TP_CUDA(tp_stream_sync(ctx));
        TP_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
        long long t0 = 0;
        for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        for (int cta = 0; cta < CI_CLUSTER; cta++)
            for (int st = 0; st < 16; st++) {
                bool any = false;
                for (int ph = 0; ph < 8; ph++) any |= h[(cta * 16 + st) * 8 + ph] != 0;
                if (!any) continue;
                fprintf(stderr, "[chol trace] cta %d step %2d:", cta, st);
                for (int ph = 0; ph < 8; ph++) {
                    const long long v = h[(cta * 16 + st) * 8 + ph];
                    if (v) fprintf(stderr, " %7lld", v - t0); else fprintf(stderr, "       -");
                }
                fprintf(stderr, "\n");
            }
        trbuf.release();
    }
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}

// G (b x b, ld) is overwritten by its Cholesky factor (lower triangle); Linv receives L^-1.  A non-positive pivot
// sets the sticky status word ctx->status[0].  bad_out != nullptr: the word is cleared before and read after the
// launch (one stream synchronisation); nullptr: nothing is read back, the caller polls tp_flags_read later.
int tp_chol_inv(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *bad_out) {
    TP_TRY(ctx->status.reserve(64));
    if (bad_out) TP_TRY(tp_flags_reset(ctx));
    TP_TRY(launch_cholinv(ctx, G, Linv, b, ld, ctx->status.as<int>(), 0));
    if (bad_out) {
        int f[4];
        TP_TRY(tp_flags_read(ctx, f));
        *bad_out = f[0];
    }
    return TP_OK;
}

// G <- Cholesky factor (lower triangle), semi-definite input tolerated (pivots clamped); no read-back
int tp_chol_factor(tp_ctx *ctx, double *G, int b, int ld) {
    TP_TRY(ctx->status.reserve(64));
    return launch_cholinv(ctx, G, nullptr, b, ld, ctx->status.as<int>(), 1);
}
