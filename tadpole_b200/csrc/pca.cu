// pca.cu -- stage 2 (sparse_cor, reference R/TADpole.R:94-100 + NaN->0 :363,449) and
//           stage 3 (prcomp(cor, rank.=k)$x, R/TADpole.R:366-367,452-453).
//
// Stage 2: colMeans and diag(cov) come from one row pass over the (symmetric) filtered matrix; the
// Gram matrix X^T X = X X^T runs on the FP64 tensor path (gemm.cu) and the covariance ->
// correlation -> NaN->0 arithmetic is the GEMM epilogue, in the reference's order of operations.
//
// Stage 3: prcomp centres the columns of C and takes the SVD; the scores are x = Xc V = U S.  U and
// S^2 are the eigenpairs of M = Xc Xc^T (n x n, PSD), so the first k score columns are
// u_j sqrt(lambda_j).  Small problems (n <= jacobi_direct_max) are solved directly by the cluster
// Jacobi solver.  Larger ones use Chebyshev-filtered subspace iteration on a block of b > k vectors:
// the filter is a three-term recurrence of GEMMs with M (fused alpha/beta/gamma epilogue),
// orthonormalisation and Rayleigh-Ritz are done together as a generalised b x b problem
// (G = Y^T Y, T = Y^T M Y) solved with two Jacobi eigen-decompositions, and convergence is tested
// on residuals ||M y - theta y|| computed from a fresh product.  Signs of the components are
// arbitrary, as they are in LAPACK; every consumer downstream is sign-invariant.
#include "common.cuh"
#include "gemm.cuh"
#include <stdlib.h>

int tp_jacobi(tp_ctx *ctx, double *A, int b, int ld, double *w, double *Vs, int lds, int ncols_out, int *sweeps_out,
              double tol, double predict = 0.0);
int tp_igram(tp_ctx *ctx, const double *X, int n, int ld, double *C, int ldc, const double *mean, const double *sd,
             int raw, int *used_out, int row_begin, int row_end, SymShard ss);
int tp_chol_inv(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *bad_out);
int tp_iop_prepare(tp_ctx *ctx, const double *S, int n, int ld);
int tp_igram_sliced(tp_ctx *ctx, const double *A, int n, int ld, double *M, int ldm, int row_begin, int row_end,
                    int sym, SymShard ss);
int tp_iop_apply(tp_ctx *ctx, const double *Yin, int b, int ldy, double *D, int ldd, double alpha, const double *E1,
                 int lde1, double beta, const double *E2, int lde2, double gamma, int row_begin, int row_end, int np);

// ---- small kernels -------------------------------------------------------------------------------
// per row: mean and sd of the one-pass formula (columns == rows: the matrix is symmetric)
__global__ void rowstats_kernel(const double *__restrict__ X, int n, int ld, double *__restrict__ mean,
                                double *__restrict__ sd) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const double *row = X + (size_t)w * ld;
    double s = 0.0, q = 0.0;
    for (int c = lane; c < n; c += 32) { const double v = row[c]; s += v; q += v * v; }
    s = warp_sum(s); q = warp_sum(q);
    if (lane == 0) {
        const double dn = (double)n;
        const double m = s / dn;
        mean[w] = m;
        // diag(covmat) = (crossprod_ii - nrow * m_i * m_i) / (nrow - 1)
        sd[w] = sqrt((q - dn * (m * m)) / (dn - 1.0));
    }
}

// column means of an n x n matrix: 32 columns per CTA, 8 row groups, fixed summation order
__global__ void __launch_bounds__(256)
colmean_kernel(const double *__restrict__ C, int n, int ld, double *__restrict__ mu) {
    __shared__ double s[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double a0 = 0.0, a1 = 0.0;
    if (c < n) {
        int r = ty;
        for (; r + 8 < n; r += 16) { a0 += C[(size_t)r * ld + c]; a1 += C[(size_t)(r + 8) * ld + c]; }
        for (; r < n; r += 8) a0 += C[(size_t)r * ld + c];
    }
    s[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty == 0 && c < n) {
        double v = 0.0;
        for (int g = 0; g < 8; g++) v += s[g][tx];
        mu[c] = v / (double)n;
    }
}

__global__ void center_cols_kernel(double *__restrict__ C, int n, int ld, const double *__restrict__ mu) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(idx / ld), c = (int)(idx % ld);
    if (r < n) C[idx] = (c < n) ? C[idx] - mu[c] : 0.0;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
// deterministic start block: uniform(-1, 1)
__global__ void random_block_kernel(double *__restrict__ Y, int n, int b, int ld) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(idx / ld), c = (int)(idx % ld);
    if (r >= n) return;
    double v = 0.0;
    if (c < b) {
        unsigned long long h = splitmix64(((unsigned long long)r << 20) ^ (unsigned long long)c ^ 0x5851f42d4c957f2dULL);
        v = (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    }
    Y[idx] = v;
}

// Gs = D G D, Ts = D (T + T^T)/2 D with D = diag(G)^-1/2 (0 where the diagonal is not positive)
__global__ void scale_gram_kernel(const double *__restrict__ G, const double *__restrict__ T, int b, int ld,
                                  double *__restrict__ Gs, double *__restrict__ Ts, double *__restrict__ dvec) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * b) return;
    const int i = idx / b, j = idx % b;
    const double gi = G[(size_t)i * ld + i], gj = G[(size_t)j * ld + j];
    const double di = gi > 0.0 ? 1.0 / sqrt(gi) : 0.0, dj = gj > 0.0 ? 1.0 / sqrt(gj) : 0.0;
    Gs[(size_t)i * ld + j] = 0.5 * (G[(size_t)i * ld + j] + G[(size_t)j * ld + i]) * di * dj;
    Ts[(size_t)i * ld + j] = 0.5 * (T[(size_t)i * ld + j] + T[(size_t)j * ld + i]) * di * dj;
    if (j == 0) dvec[i] = di;
}

// X = S diag(g^-1/2), directions with g_i <= eps * g_0 dropped (zero column)
__global__ void whiten_kernel(const double *__restrict__ S, const double *__restrict__ g, int b, int ld,
                              double *__restrict__ X) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * b) return;
    const int i = idx / b, j = idx % b;
    const double gj = g[j];
    const double sc = (gj > 1e-14 * g[0] && gj > 0.0) ? 1.0 / sqrt(gj) : 0.0;
    X[(size_t)i * ld + j] = S[(size_t)i * ld + j] * sc;
}

__global__ void symmetrize_kernel(double *__restrict__ A, int b, int ld) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * b) return;
    const int i = idx / b, j = idx % b;
    if (i < j) {
        const double v = 0.5 * (A[(size_t)i * ld + j] + A[(size_t)j * ld + i]);
        A[(size_t)i * ld + j] = v;
        A[(size_t)j * ld + i] = v;
    }
}

__global__ void scale_rows_kernel(double *__restrict__ Q, int b, int ld, const double *__restrict__ d) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * b) return;
    const int i = idx / b, j = idx % b;
    Q[(size_t)i * ld + j] *= d[i];
}

// residual_j = || MY_j - theta_j Y_j ||_2 where MY = (e / sigma1) Y1 + c Y0 (first filter step).  Two stages with a fixed
// summation order (every rank, every run: the same bits): RS_SLABS row slabs x 32-column groups of partial sums of squares,
// then one thread per column adds the slabs.  (One CTA per 32 columns walking all the rows was 7 CTAs and 0.13 ms at
// 2000 bins, four times per call.)
#define RS_SLABS 32
__global__ void __launch_bounds__(256)
residual_partial_kernel(const double *__restrict__ Y0, const double *__restrict__ Y1, int n, int ld, int k,
                        const double *__restrict__ theta, double e_over_sig, double c, double *__restrict__ part) {
    __shared__ double s[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    const int rows = (n + RS_SLABS - 1) / RS_SLABS, r0 = blockIdx.y * rows, r1 = min(n, r0 + rows);
    double a = 0.0;
    if (j < k) {
        const double th = theta[j];
        for (int r = r0 + ty; r < r1; r += 8) {
            const double y0 = Y0[(size_t)r * ld + j];
            const double v = e_over_sig * Y1[(size_t)r * ld + j] + (c - th) * y0;
            a += v * v;
        }
    }
    s[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && j < k) {
        double v = 0.0;
        for (int g = 0; g < 8; g++) v += s[g][tx];
        part[(size_t)blockIdx.y * k + j] = v;
    }
}
__global__ void residual_final_kernel(const double *__restrict__ part, int k, double *__restrict__ res) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    double v = 0.0;
    for (int g = 0; g < RS_SLABS; g++) v += part[(size_t)g * k + j];
    res[j] = sqrt(v);
}

// scores[:, j] = U[:, j] * sqrt(max(w_j, 0)), j < k; padding columns zero
__global__ void scores_kernel(const double *__restrict__ U, int ldu, const double *__restrict__ w, int n, int k,
                              double *__restrict__ S, int lds) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(idx / lds), c = (int)(idx % lds);
    if (r >= n) return;
    S[idx] = (c < k) ? U[(size_t)r * ldu + c] * sqrt(fmax(w[c], 0.0)) : 0.0;
}

// ---- stage 2 -------------------------------------------------------------------------------------
int tp_correlation(tp_ctx *ctx) {
    if (tp_group_dispatch(ctx)) return tp_group_run(ctx, [&](tp_ctx *gc, int) -> int { return tp_correlation(gc); });
    TP_ARG(ctx && ctx->have_X, "tp_correlation: no filtered matrix in the context");
    TP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = ctx->nf, ld = ctx->ldx;
    const bool shard = tp_row_sharded(ctx, n);
    const TpRows rw = tp_rows(ctx, n);
    TP_TRY(ctx->C.reserve((size_t)(shard ? rw.padded : n) * ld * sizeof(double)));
    TP_TRY(ctx->colstat.reserve((size_t)3 * n * sizeof(double)));
    double *mean = ctx->colstat.as<double>(), *sd = mean + n;
    TP_MARK(ctx, EV_CORR0);
    rowstats_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(ctx->X.as<double>(), n, ld, mean, sd);
    ctx->launches += 1;
    GemmArgs g;
    g.A = ctx->X.as<double>(); g.lda = ld; g.a_kc = 1;
    g.B = ctx->X.as<double>(); g.ldb = ld; g.b_kc = 1;     // B = X^T: (k, n) at X[n*ld + k]
    g.D = ctx->C.as<double>(); g.ldd = ld;
    g.M = n; g.N = n; g.K = n;
    g.sym = 1; g.epi = EPI_CORR; g.mean = mean; g.sd = sd; g.nrows = (double)n;
    int igemm_used = 0;
    if (ctx->igemm_min_n > 0 && n >= ctx->igemm_min_n) {
        // integer counts: exact Gram on the tcgen05 int8 path with the same epilogue (igemm.cu).  With several ranks
        // every block pair is computed once (SymShard, common.cuh), the row blocks are all-gathered and the rest is a
        // local transpose; the elements are the same bits as those of the one-GPU launch.
        const bool ssym = shard && ctx->shard_sym;
        if (ssym) {
            const SymShard ss{tp_nranks(ctx), rw.rpr};
            TP_TRY(tp_igram(ctx, ctx->X.as<double>(), n, ld, ctx->C.as<double>(), ld, mean, sd, 0, &igemm_used, rw.r0, rw.r1, ss));
            if (igemm_used) {
                TP_TRY(tp_comm_allgather(ctx, ctx->C.as<double>(), (size_t)rw.rpr * ld));
                TP_TRY(tp_mirror_fill(ctx, ctx->C.as<double>(), n, ld, ss));
            }
        } else {
            TP_TRY(tp_igram(ctx, ctx->X.as<double>(), n, ld, ctx->C.as<double>(), ld, mean, sd, 0, &igemm_used, 0, n,
                            SymShard{1, 1 << 30}));
        }
    }
    if (igemm_used) {
    } else if (shard) {
        // row block [r0, r1) of the correlation matrix on this rank, then NCCL all-gather of the row blocks.  Every
        // element is the same k-ordered sum as in the symmetric single-GPU launch, so the result is bit-identical.
        g.A += (size_t)rw.r0 * ld; g.D += (size_t)rw.r0 * ld; g.M = rw.r1 - rw.r0; g.sym = 0; g.epi_row0 = rw.r0;
        if (g.M > 0) TP_TRY(tp_gemm(ctx, g));
        TP_TRY(tp_comm_allgather(ctx, ctx->C.as<double>(), (size_t)rw.rpr * ld));
    } else {
        TP_TRY(tp_gemm(ctx, g));
    }
    TP_MARK(ctx, EV_CORR1);
    ctx->have_C = true;
    ctx->have_scores = ctx->have_sweep = false;
    return TP_OK;
}

// ---- stage 3 -------------------------------------------------------------------------------------
struct PcaOp {
    tp_ctx *ctx;
    int n, ld;            // Xc is n x n (ld)
    const double *Xc;
    const double *M;      // explicit M (n x ld) or nullptr
    double *Z;            // scratch n x ldb for the two-GEMM form
    int b, ldb;
    bool shard = false;   // row blocks computed by their owner rank and all-gathered (buffers hold rw.padded rows)
    TpRows rw{};
    int np = 0;           // 0: FP64 DMMA operator; 5 / 8: sliced int8 operator on the tcgen05 tensor cores with that many
                          // digit planes (igemm.cu; needs explicit M): 5 for the early rounds, 8 is FP64-level
    long applications = 0, lowprec_applications = 0;
    // D[rows] = alpha * A[rows, :] * B + beta * E1[rows] + gamma * E2[rows]; rows = all, or this rank's block
    int rows_gemm(GemmArgs g, double *D, const double *E1, const double *E2) {
        g.D = D; g.E1 = E1; g.E2 = E2;
        if (!shard) return tp_gemm(ctx, g);
        const size_t r0 = (size_t)rw.r0;
        g.A += g.a_kc ? r0 * g.lda : r0;
        g.D += r0 * g.ldd;
        if (g.E1) g.E1 += r0 * g.lde1;
        if (g.E2) g.E2 += r0 * g.lde2;
        g.M = rw.r1 - rw.r0;
        if (g.M > 0) TP_TRY(tp_gemm(ctx, g));
        return tp_comm_allgather(ctx, D, (size_t)rw.rpr * g.ldd);
    }
    // Yout = alpha * M * Yin + beta * E1 + gamma * E2
    int apply(const double *Yin, double *Yout, double alpha, const double *E1, double beta, const double *E2,
              double gamma) {
        GemmArgs g;
        g.M = n; g.N = b; g.ldd = ldb; g.alpha = alpha;
        g.lde1 = ldb; g.beta = beta; g.lde2 = ldb; g.gamma = gamma;
        applications++;
        if (M && np) {
            lowprec_applications++;
            const int r0 = shard ? rw.r0 : 0, r1 = shard ? rw.r1 : n;
            if (r1 > r0)
                TP_TRY(tp_iop_apply(ctx, Yin, b, ldb, Yout, ldb, alpha, E1, ldb, beta, E2, ldb, gamma, r0, r1, np));
            return shard ? tp_comm_allgather(ctx, Yout, (size_t)rw.rpr * ldb) : (int)TP_OK;
        }
        if (M) {
            g.A = M; g.lda = ld; g.a_kc = 1; g.B = Yin; g.ldb = ldb; g.b_kc = 0; g.K = n;
            return rows_gemm(g, Yout, E1, E2);
        }
        GemmArgs z;   // Z = Xc^T Yin
        z.A = Xc; z.lda = ld; z.a_kc = 0; z.B = Yin; z.ldb = ldb; z.b_kc = 0;
        z.ldd = ldb; z.M = n; z.N = b; z.K = n;
        TP_TRY(rows_gemm(z, Z, nullptr, nullptr));
        g.A = Xc; g.lda = ld; g.a_kc = 1; g.B = Z; g.ldb = ldb; g.b_kc = 0; g.K = n;
        return rows_gemm(g, Yout, E1, E2);
    }
};

static int small_gemm(tp_ctx *ctx, const double *A, int a_kc, const double *B, int b_kc, double *D, int b, int ld) {
    GemmArgs g;
    g.A = A; g.lda = ld; g.a_kc = a_kc; g.B = B; g.ldb = ld; g.b_kc = b_kc;
    g.D = D; g.ldd = ld; g.M = b; g.N = b; g.K = b;
    return tp_gemm(ctx, g);
}

int tp_pca(tp_ctx *ctx, int max_pcs, int *k_out) {
    if (tp_group_dispatch(ctx))
        return tp_group_run(ctx, [&](tp_ctx *gc, int gr) -> int { return tp_pca(gc, max_pcs, gr ? nullptr : k_out); });
    TP_ARG(ctx && ctx->have_C, "tp_pca: no correlation matrix in the context");
    TP_ARG(max_pcs >= 1, "tp_pca: max_pcs must be >= 1");
    TP_CUDA(cudaSetDevice(ctx->device));
    ctx->generation++;
    cudaStream_t st = ctx->stream;
    const int n = ctx->nf, ld = ctx->ldx;
    const int k = max_pcs < n ? max_pcs : n;       // number_pca <- min(max_pcs, nrow(mat))
    const int ldk = round_up(k, 8);
    if (k_out) *k_out = k;
    TP_MARK(ctx, EV_PCA0);
    TP_TRY(ctx->scores.reserve((size_t)n * ldk * sizeof(double)));
    TP_TRY(ctx->colstat.reserve((size_t)3 * n * sizeof(double)));
    double *mu = ctx->colstat.as<double>() + 2 * n;
    double *C = ctx->C.as<double>();
    colmean_kernel<<<(n + 31) / 32, 256, 0, st>>>(C, n, ld, mu);
    center_cols_kernel<<<(unsigned)(((size_t)n * ld + 255) / 256), 256, 0, st>>>(C, n, ld, mu);
    ctx->launches += 2;
    ctx->have_C = false;                               // C now holds Xc
    ctx->timing[7] = ctx->timing[8] = ctx->timing[9] = 0.0;

    // block width: k wanted vectors + a guard band (a quarter of k, at least 32).  The b x b Rayleigh-Ritz, Cholesky and
    // eigen problems are solved on chip (the cluster Jacobi solver keeps its share of the b x b matrix in registers): b <= 704,
    // so above 1024 bins (where the direct solver no longer applies) the guard band narrows for large k and max_pcs stops at
    // 672.  prcomp(rank. = ) has no such limit.
    int b = ctx->pca_block > 0 ? ctx->pca_block : round_up(k + (k / 4 > 32 ? k / 4 : 32), 32);
    if (n > 1024 && n > ctx->jacobi_direct_max) {
        if (b > 704) b = 704;
        if (k + 32 > b) {
            tp_set_error("tp_pca: max_pcs = %d is not supported for matrices above 1024 bins (this one has %d good bins): the "
                         "subspace iteration keeps k + 32 <= 704 vectors on chip, so max_pcs <= 672", max_pcs, n);
            return TP_ERR_ARG;
        }
    }
    const bool direct = n <= ctx->jacobi_direct_max || b >= n;
    const bool use_iop = !direct && ctx->iop_min_n > 0 && n >= ctx->iop_min_n;
    const bool explicitM = direct || n <= 12288 || use_iop;
    double *M = nullptr;
    const bool shard = !direct && tp_row_sharded(ctx, n);
    const TpRows rw = tp_rows(ctx, n);
    const size_t nrows_alloc = shard ? (size_t)rw.padded : (size_t)n;
    if (explicitM) {
        TP_TRY(ctx->M.reserve(nrows_alloc * ld * sizeof(double)));
        M = ctx->M.as<double>();
        GemmArgs g;
        g.A = C; g.lda = ld; g.a_kc = 1; g.B = C; g.ldb = ld; g.b_kc = 1;
        g.D = M; g.ldd = ld; g.M = n; g.N = n; g.K = n; g.sym = 1;
        if (use_iop && ctx->mgram_min_n > 0 && n >= ctx->mgram_min_n) {
            // sliced int8 Gram on the tcgen05 tensor cores (8 digit planes of Xc, FP64 level); row blocks when sharded
            // (several ranks: every block pair once, all-gather, local transpose -- SymShard in common.cuh)
            const int r0 = shard ? rw.r0 : 0, r1 = shard ? rw.r1 : n;
            const bool ssym = shard && ctx->shard_sym;
            const SymShard ss = ssym ? SymShard{tp_nranks(ctx), rw.rpr} : SymShard{1, 1 << 30};
            TP_TRY(tp_igram_sliced(ctx, C, n, ld, M, ld, r0, r1, shard ? (ssym ? 1 : 0) : 1, ss));
            if (shard) TP_TRY(tp_comm_allgather(ctx, M, (size_t)rw.rpr * ld));
            if (ssym) TP_TRY(tp_mirror_fill(ctx, M, n, ld, ss));
        } else if (shard) {      // row blocks of M = Xc Xc^T by their owner, all-gathered
            g.A += (size_t)rw.r0 * ld; g.D += (size_t)rw.r0 * ld; g.M = rw.r1 - rw.r0; g.sym = 0;
            if (g.M > 0) TP_TRY(tp_gemm(ctx, g));
            TP_TRY(tp_comm_allgather(ctx, M, (size_t)rw.rpr * ld));
        } else {
            TP_TRY(tp_gemm(ctx, g));
        }
    }
    if (direct) {
        TP_ARG(n <= 1024, "tp_pca: direct eigensolver limited to 1024 bins (jacobi_direct_max / pca_block set too high)");
        TP_TRY(ctx->W.reserve((size_t)n * ld * sizeof(double)));
        TP_TRY(ctx->Jw.reserve((size_t)n * sizeof(double)));
        int sweeps = 0;
        TP_TRY(tp_jacobi(ctx, M, n, ld, ctx->Jw.as<double>(), ctx->W.as<double>(), ld, n, &sweeps, 1e-14));
        ctx->timing[9] = sweeps;
        scores_kernel<<<(unsigned)(((size_t)n * ldk + 255) / 256), 256, 0, st>>>(ctx->W.as<double>(), ld,
                                                                               ctx->Jw.as<double>(), n, k,
                                                                               ctx->scores.as<double>(), ldk);
        ctx->launches += 1;
    } else {
        const int ldb = round_up(b, 8);
        const size_t blk = nrows_alloc * ldb * sizeof(double);
        const size_t sm = (size_t)b * ldb * sizeof(double);
        TP_TRY(ctx->Y0.reserve(blk)); TP_TRY(ctx->Y1.reserve(blk)); TP_TRY(ctx->Y2.reserve(blk));
        TP_TRY(ctx->W.reserve(blk));
        TP_TRY(ctx->G.reserve(sm)); TP_TRY(ctx->T.reserve(sm)); TP_TRY(ctx->Q.reserve(sm));
        TP_TRY(ctx->small1.reserve(sm)); TP_TRY(ctx->small2.reserve(sm)); TP_TRY(ctx->Jv.reserve(sm));
        TP_TRY(ctx->Jw.reserve((size_t)4 * b * sizeof(double)));
        TP_TRY(ctx->resid.reserve((size_t)(k + b) * sizeof(double)));
        TP_TRY(ctx->part.reserve((size_t)RS_SLABS * k * sizeof(double)));
        TP_TRY(tp_pin_reserve(ctx, (size_t)(k + b + 8) * sizeof(double)));
        double *W = ctx->W.as<double>();
        double *G = ctx->G.as<double>(), *T = ctx->T.as<double>(), *Q = ctx->Q.as<double>();
        double *S1 = ctx->small1.as<double>(), *S2 = ctx->small2.as<double>(), *JV = ctx->Jv.as<double>();
        double *gval = ctx->Jw.as<double>(), *theta = gval + b, *dvec = theta + b;
        double *res = ctx->resid.as<double>();
        PcaOp op{ctx, n, ld, C, M, nullptr, b, ldb};
        op.shard = shard; op.rw = rw;
        if (use_iop) TP_TRY(tp_iop_prepare(ctx, M, n, ld));
        DevBuf zbuf;   // scratch for the two-GEMM operator
        if (!M) { TP_TRY(zbuf.reserve(blk)); op.Z = zbuf.as<double>(); }

        const unsigned gb = (unsigned)((b * b + 255) / 256);
        const int splitk = n >= 4096 ? 32 : (n >= 1024 ? 8 : 1);
        // three rotating n x b buffers: Y = current block, F1 / F2 = scratch
        double *Y = ctx->Y0.as<double>(), *F1 = ctx->Y1.as<double>(), *F2 = ctx->Y2.as<double>();
        int sweeps_total = 0, rc = TP_OK, it = 0;
        bool converged = false;
        double last_res = 1.0;

        // D = Ysrc^T Bsrc  (b x b, long reduction over the bins: deterministic split-K)
        auto gram = [&](const double *Ysrc, const double *Bsrc, double *D) {
            GemmArgs gg;
            gg.A = Ysrc; gg.lda = ldb; gg.a_kc = 0; gg.B = Bsrc; gg.ldb = ldb; gg.b_kc = 0;
            gg.D = D; gg.ldd = ldb; gg.M = b; gg.N = b; gg.K = n; gg.splitk = splitk;
            return tp_gemm(ctx, gg);
        };
        // F1 = Y * R (R: b x b, (k,n) at R[k*ldb+n] or transposed), then F1 becomes the block
        auto rotate = [&](const double *R, int r_kc) {
            GemmArgs r;
            r.A = Y; r.lda = ldb; r.a_kc = 1; r.B = R; r.ldb = ldb; r.b_kc = r_kc;
            r.D = F1; r.ldd = ldb; r.M = n; r.N = b; r.K = b;
            int e = tp_gemm(ctx, r);
            double *t = Y; Y = F1; F1 = t;
            return e;
        };
        // Cholesky QR: Y <- Y L^-T.  One pass leaves ~cond^2 * eps of non-orthogonality, which is enough
        // between two filter rounds; two passes before a Rayleigh-Ritz step.  *bad = 1 when G is not PD.
        // Fast mode (safe == false): nothing is read back here; a non-positive pivot sets a sticky status word that
        // body() polls at its own synchronisation points and answers by running the whole solve again in safe mode.
        bool safe = false;
        auto cholqr = [&](int passes, int *bad) {
            *bad = 0;
            for (int pass = 0; pass < passes; pass++) {
                TP_TRY(gram(Y, Y, G));
                TP_TRY(tp_chol_inv(ctx, G, S1, b, ldb, safe ? bad : nullptr));
                if (*bad) return (int)TP_OK;
                TP_TRY(rotate(S1, 1));          // B = Linv^T: (k, n) at Linv[n*ldb + k]
            }
            return (int)TP_OK;
        };
        // eigensolver call: synchronous status in safe mode only
        auto eig = [&](double *A, double *w, double jtol) {
            int sw = 0;
            // (fast mode: the sweep that would only confirm convergence is skipped on the quadratic-convergence prediction
            //  for relative gaps >= 1e-3; the residual test of the next iteration is the judge of the Ritz pairs anyway)
            TP_TRY(tp_jacobi(ctx, A, b, ldb, w, JV, ldb, b, safe ? &sw : nullptr, jtol, safe ? 0.0 : 1e6));
            sweeps_total += sw;
            return (int)TP_OK;
        };
        // status words of the fast mode, polled where the host waits anyway: 0 fine, 1 rerun in safe mode
        const int RERUN = -1000;
        // (the copy is enqueued before the stream synchronisation the caller does anyway)
        auto poll = [&]() -> int {
            if (safe) return TP_OK;
            const int *f = ctx->pin_flags;
            if (f[0]) return RERUN;
            if (f[1]) { tp_set_error("tp_pca: b x b eigensolver did not converge (b = %d)", b); return TP_ERR_NOCONV; }
            sweeps_total = f[2];
            return TP_OK;
        };
        // Rayleigh-Ritz on an orthonormal block: T = Y^T W, T = Z diag(theta) Z^T, Y <- Y Z
        auto rr_orthonormal = [&](double jtol) {
            TP_TRY(gram(Y, W, T));
            symmetrize_kernel<<<gb, 256, 0, st>>>(T, b, ldb);
            ctx->launches += 1;
            TP_TRY(eig(T, theta, jtol));
            return rotate(JV, 0);
        };
        // orthonormalisation and Rayleigh-Ritz together as the generalised problem (G = Y^T Y,
        // T = Y^T W) through two eigen-decompositions: robust for ill-conditioned blocks
        auto rr_general = [&](double jtol) {
            TP_TRY(gram(Y, Y, G));
            TP_TRY(gram(Y, W, T));
            scale_gram_kernel<<<gb, 256, 0, st>>>(G, T, b, ldb, S1, S2, dvec);       // S1 = Gs, S2 = Ts
            ctx->launches += 1;
            TP_TRY(eig(S1, gval, jtol));                                             // Gs = JV diag(gval) JV^T
            whiten_kernel<<<gb, 256, 0, st>>>(JV, gval, b, ldb, S1);                 // S1 = X
            ctx->launches += 1;
            TP_TRY(small_gemm(ctx, S2, 1, S1, 0, G, b, ldb));                        // G  = Ts X
            TP_TRY(small_gemm(ctx, S1, 0, G, 0, T, b, ldb));                         // T  = X^T Ts X
            symmetrize_kernel<<<gb, 256, 0, st>>>(T, b, ldb);
            ctx->launches += 1;
            TP_TRY(eig(T, theta, jtol));                                             // T = JV diag(theta) JV^T
            TP_TRY(small_gemm(ctx, S1, 1, JV, 0, Q, b, ldb));                        // Q = X Z
            scale_rows_kernel<<<gb, 256, 0, st>>>(Q, b, ldb, dvec);                  // Q = D X Z
            ctx->launches += 1;
            return rotate(Q, 0);
        };
        // Chebyshev filter steps 2..deg on P0 = Y, P1 = F1 (step 1 already done); result becomes Y
        struct Bounds { double top, thk, cut, e, c, sig1; int deg; } bd{};
        auto filter_rest = [&]() {
            double sig = bd.sig1;
            double *P0 = Y, *P1 = F1, *P2 = F2;
            for (int j = 2; j <= bd.deg; j++) {
                const double sig2 = 1.0 / (2.0 / bd.sig1 - sig);
                // P2 = 2 (sig2/e) (M P1 - c P1) - sig sig2 P0
                TP_TRY(op.apply(P1, P2, 2.0 * sig2 / bd.e, P1, -2.0 * sig2 * bd.c / bd.e, P0, -sig * sig2));
                double *t = P0; P0 = P1; P1 = P2; P2 = t;
                sig = sig2;
            }
            Y = P1; F1 = P0; F2 = P2;
            return (int)TP_OK;
        };
        auto filter_step1 = [&]() {   // F1 = (sig1 / e) (M Y - c Y)
            return op.apply(Y, F1, bd.sig1 / bd.e, Y, -bd.sig1 * bd.c / bd.e, nullptr, 0.0);
        };

        auto body = [&]() -> int {
            TP_TRY(tp_flags_reset(ctx));
            sweeps_total = 0; it = 0; converged = false; op.applications = 0; op.lowprec_applications = 0;
            op.np = use_iop ? 5 : 0;
            double prev_res = 1e300;
            Y = ctx->Y0.as<double>(); F1 = ctx->Y1.as<double>(); F2 = ctx->Y2.as<double>();
            random_block_kernel<<<(unsigned)(((size_t)n * ldb + 255) / 256), 256, 0, st>>>(Y, n, b, ldb);
            ctx->launches += 1;
            {   // start: orthonormalise the random block, one Rayleigh-Ritz step at low accuracy
                // (one pass: an n x b block of uniform random numbers has condition number ~(sqrt(n) + sqrt(b)) / (sqrt(n) -
                //  sqrt(b)) < 5 for n > 2 b, so Cholesky QR leaves cond^2 eps ~ 1e-15 of non-orthogonality; closer to square
                //  the block is worse conditioned and gets the second pass)
                int bad = 0;
                TP_TRY(cholqr(n > 2 * b ? 1 : 2, &bad));
                TP_TRY(op.apply(Y, W, 1.0, nullptr, 0.0, nullptr, 0.0));
                if (bad) TP_TRY(rr_general(1e-2)); else TP_TRY(rr_orthonormal(1e-2));      // only the Ritz VALUES of the start step are used (filter bounds)
            }
            double *hbuf = (double *)ctx->pin;
            const int inner = ctx->pca_inner;   // filter + CholQR rounds between two Rayleigh-Ritz steps
            for (it = 1; it <= ctx->pca_maxit * 2; it++) {
                // ---- bounds from the current Ritz values -----------------------------------------
                TP_CUDA(cudaMemcpyAsync(hbuf, theta, (size_t)b * sizeof(double), cudaMemcpyDeviceToHost, st));
                TP_TRY(tp_flags_enqueue(ctx));
                TP_CUDA(tp_stream_sync(ctx));
                TP_TRY(poll());
                hbuf = (double *)ctx->pin;
                bd.top = hbuf[0]; bd.thk = hbuf[k - 1];
                double cut = hbuf[b - 1];
                if (!(cut > 0.0)) { for (int j = b - 1; j >= k; j--) if (hbuf[j] > 0.0) { cut = hbuf[j]; break; } }
                if (!(cut > 0.0) || !(bd.thk > cut)) cut = 0.5 * bd.thk > 0.0 ? 0.5 * bd.thk : 1e-300;
                bd.cut = cut; bd.e = 0.5 * cut; bd.c = 0.5 * cut;
                bd.sig1 = bd.e / (bd.top - bd.c);
                const double xk = (bd.thk - bd.c) / bd.e, x1 = (bd.top - bd.c) / bd.e;
                bd.deg = 1;      // degree bounded by the conditioning of the filtered block
                for (int mdeg = 2; mdeg <= 24; mdeg++) {
                    const double ratio = cosh(mdeg * acosh(x1)) / cosh(mdeg * acosh(xk));
                    if (ratio <= 1e5) bd.deg = mdeg; else break;
                }
                // ---- first filter step doubles as the residual check ---------------------------------
                double rmax = 0.0;
                for (int attempt = 0; attempt < 2; attempt++) {
                    TP_TRY(filter_step1());
                    residual_partial_kernel<<<dim3((k + 31) / 32, RS_SLABS), 256, 0, st>>>(Y, F1, n, ldb, k, theta, bd.e / bd.sig1,
                                                                                         bd.c, ctx->part.as<double>());
                    residual_final_kernel<<<(k + 127) / 128, 128, 0, st>>>(ctx->part.as<double>(), k, res);
                    ctx->launches += 2;
                    TP_CUDA(cudaMemcpyAsync(hbuf, res, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, st));
                    TP_TRY(tp_flags_enqueue(ctx));
                    TP_CUDA(tp_stream_sync(ctx));
                    TP_TRY(poll());
                    rmax = 0.0;
                    for (int j = 0; j < k; j++) rmax = (hbuf[j] > rmax || hbuf[j] != hbuf[j]) ? hbuf[j] : rmax;
                    last_res = rmax / bd.top;
                    if (getenv("TADPOLE_DEBUG"))
                        fprintf(stderr, "[tadpole] pca it=%d res=%.3e top=%.4e thk=%.4e cut=%.4e deg=%d %s eig_sweeps_so_far=%d apps=%d\n", it, last_res,
                                bd.top, bd.thk, bd.cut, bd.deg, op.np == 5 ? "int8x5" : (op.np == 8 ? "int8x8" : "fp64"), sweeps_total, op.applications);
                    // the 5-plane sliced operator (terms below 2^-35 of row scale x column scale left out) carries the
                    // iteration down to iop_switch (or until it stops helping).  From there on, and for the residual that
                    // decides convergence, the operator is the 8-plane sliced one (terms below 2^-56 left out: FP64 level,
                    // every product kept is exact; nf >= iop_final_min_n and iop_final = 8, the defaults) or else the FP64
                    // DMMA GEMM.  The 8-plane residual agrees with the FP64 DMMA one to ~1e-16 of theta_1, four orders
                    // below pca_tol (tests/iop_compare.py; scores vs the pure FP64 solve: 1.3e-11 of the largest score).
                    if (op.np != 5 || !(last_res <= ctx->iop_switch || last_res > 0.1 * prev_res)) break;
                    op.np = n >= ctx->iop_final_min_n ? ctx->iop_final : 0;
                }
                prev_res = last_res;
                if (rmax <= ctx->pca_tol * bd.top) { converged = true; return TP_OK; }
                // ---- filter / orthonormalise rounds, then one Rayleigh-Ritz ------------------------
                bool general = false;
                for (int r = 0; r < inner; r++) {
                    if (r > 0) TP_TRY(filter_step1());
                    TP_TRY(filter_rest());
                    int bad = 0;
                    TP_TRY(cholqr(r + 1 == inner ? 2 : 1, &bad));
                    if (bad) { general = true; break; }
                }
                TP_TRY(op.apply(Y, W, 1.0, nullptr, 0.0, nullptr, 0.0));
                const double jtol = fmin(1e-6, fmax(1e-14, last_res * 1e-9));
                if (general) TP_TRY(rr_general(jtol)); else TP_TRY(rr_orthonormal(jtol));
            }
            return TP_OK;
        };
        rc = body();
        if (rc == RERUN) {           // rank-deficient block met by the fast path: slow and careful this time
            safe = true;
            rc = body();
        }
        zbuf.release();
        TP_TRY(rc);
        ctx->timing[7] = it;
        ctx->timing[8] = (double)op.applications;
        ctx->lowprec_applications = op.lowprec_applications;
        ctx->timing[9] = sweeps_total;
        if (!converged) {
            tp_set_error("tp_pca: subspace iteration stopped at relative residual %.3e after %d iterations", last_res, it);
            return TP_ERR_NOCONV;
        }
        scores_kernel<<<(unsigned)(((size_t)n * ldk + 255) / 256), 256, 0, st>>>(Y, ldb, theta, n, k,
                                                                               ctx->scores.as<double>(), ldk);
        ctx->launches += 1;
    }
    TP_CUDA(cudaGetLastError());
    TP_MARK(ctx, EV_PCA1);
    ctx->k = ctx->k_full = k; ctx->ldk = ldk;
    ctx->have_scores = true;
    ctx->have_sweep = false;
    return TP_OK;
}
