// jacobi.cu -- dense symmetric eigensolver for the small problems of stage 3 (prcomp).
//
// Parallel cyclic Jacobi.  One thread-block cluster of 8 CTAs (8192 threads) owns the whole b x b
// matrix (L2 resident, b <= ~600).  A sweep is b-1 steps of a round-robin tournament; in each step
// the b/2 index pairs are disjoint, so all rotations are computed from the old matrix and applied
// at once: every 2 x 2 block (I, J) of the pair-permuted matrix becomes J_I^T * block * J_J,
// independently of every other block.  Old and new matrices ping-pong between two buffers, so
// the only synchronisation is one hardware cluster barrier per step.
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define JC_CLUSTER 8
#define JC_THREADS 512
#define JC_MAXPAIRS 512

__device__ __forceinline__ void jc_pair(int slot, int step, int m, int &p, int &q) {
    // round-robin tournament over m (even) players: player m-1 is fixed, the others rotate
    int a, b;
    if (slot == 0) { a = m - 1; b = step; }
    else {
        a = (step + slot) % (m - 1);
        b = (step - slot + (m - 1)) % (m - 1);
    }
    p = min(a, b); q = max(a, b);
}

// One step = (A) rotations of this step from the old matrix, (V) eigenvector update with the
// rotations of the PREVIOUS step (so it overlaps the latency of A), (B) all 2x2 blocks of A.
// Everything a thread needs from global memory for (B) is loaded before the rotations are known.
#define JC_TA 4    // A tasks loaded per batch
__global__ void __cluster_dims__(JC_CLUSTER, 1, 1) __launch_bounds__(JC_THREADS, 1)
jacobi_kernel(double *A0, double *A1, double *V0, double *V1,
              int b, int ld, int max_sweeps, double tol, int *info) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_c[2][JC_MAXPAIRS], s_s[2][JC_MAXPAIRS], s_t[JC_MAXPAIRS];
    __shared__ short s_p[2][JC_MAXPAIRS], s_q[2][JC_MAXPAIRS];
    __shared__ double s_red[JC_THREADS / 32];
    __shared__ double s_anorm;
    __shared__ double s_off;               // max |a_pq| seen in the sweep
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gtid = cluster.block_rank() * JC_THREADS + tid;
    const int GT = JC_CLUSTER * JC_THREADS;
    const int m = (b + 1) & ~1;          // even number of players; index b (if any) is a bye
    const int np = m / 2;

    // V = I; anorm = max |diag(A)|
    for (int idx = gtid; idx < b * b; idx += GT) {
        int r = idx / b, c = idx % b;
        V0[(size_t)r * ld + c] = (r == c) ? 1.0 : 0.0;
    }
    {
        double mx = 0.0;
        for (int i = tid; i < b; i += JC_THREADS) mx = fmax(mx, fabs(A0[(size_t)i * ld + i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[wid] = mx;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < JC_THREADS / 32; w++) v = fmax(v, s_red[w]);
            s_anorm = v;
        }
        __syncthreads();
    }
    const double tolabs = tol * s_anorm;
    cluster.sync();

    double *Ao = A0, *An = A1, *Vo = V0, *Vn = V1;
    int sweeps = 0, cur = 0;
    bool have_prev = false;              // rotations of the previous step still to be applied to V
    bool converged = (b < 2);
    const int tasksA = np * np, tasksV = b * np;

    // V <- V * J(previous step): reads Vo, writes Vn with the rotation table `tb`
    auto apply_v = [&](int tb) {
        for (int t0 = gtid; t0 < tasksV; t0 += GT * 4) {
            double v0[4], v1[4];
            int rr[4], r1[4], r2[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int tv = t0 + u * GT;
                rr[u] = -1;
                if (tv < tasksV) {
                    rr[u] = tv / np;
                    const int J = tv % np;
                    r1[u] = s_p[tb][J]; r2[u] = s_q[tb][J];
                    v0[u] = __ldcg(&Vo[(size_t)rr[u] * ld + r1[u]]);
                    v1[u] = (r2[u] < b) ? __ldcg(&Vo[(size_t)rr[u] * ld + r2[u]]) : 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (rr[u] < 0) continue;
                const int J = (t0 + u * GT) % np;
                const double cJ = s_c[tb][J], sJ = s_s[tb][J];
                Vn[(size_t)rr[u] * ld + r1[u]] = cJ * v0[u] - sJ * v1[u];
                if (r2[u] < b) Vn[(size_t)rr[u] * ld + r2[u]] = sJ * v0[u] + cJ * v1[u];
            }
        }
    };

    while (!converged && sweeps < max_sweeps) {
        double offmax = 0.0;               // per thread; reduced once per sweep (no atomics in the step loop)
        for (int step = 0; step < m - 1; step++) {
            // ---- (A) loads for this step's rotations ---------------------------------------------
            int p = 0, q = 0;
            double app = 0.0, aqq = 0.0, apq = 0.0;
            if (tid < np) {
                jc_pair(tid, step, m, p, q);
                if (q < b) {
                    app = __ldcg(&Ao[(size_t)p * ld + p]);
                    aqq = __ldcg(&Ao[(size_t)q * ld + q]);
                    apq = __ldcg(&Ao[(size_t)p * ld + q]);
                }
            }
            // ---- (V) eigenvectors catch up with the previous step while those loads fly ---------
            if (have_prev) {
                apply_v(cur ^ 1);
                double *tV = Vo; Vo = Vn; Vn = tV;
            }
            // ---- (A) rotations -------------------------------------------------------------------
            if (tid < np) {
                double c = 1.0, s = 0.0, t = 0.0;
                if (q < b) {
                    const double aoff = fabs(apq);
                    offmax = fmax(offmax, aoff);
                    if (aoff > 1e-300 && aoff > 1e-18 * sqrt(fabs(app) * fabs(aqq))) {
                        const double tau = (aqq - app) / (2.0 * apq);
                        t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = rsqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                s_c[cur][tid] = c; s_s[cur][tid] = s; s_t[tid] = t; s_p[cur][tid] = (short)p; s_q[cur][tid] = (short)q;
            }
            // ---- (B) 2x2 blocks: loads first (they do not depend on the rotations) -------------
            bool synced = false;
            for (int base = 0; base < tasksA; base += GT * JC_TA) {      // uniform trip count: barrier inside
                const int t0 = base + gtid;
                double b00[JC_TA], b01[JC_TA], b10[JC_TA], b11[JC_TA];
                int pi[JC_TA], qi[JC_TA], ri[JC_TA], si[JC_TA];
                // pair indices come from the schedule, not from shared memory, so no barrier is needed yet
#pragma unroll
                for (int u = 0; u < JC_TA; u++) {
                    const int task = t0 + u * GT;
                    pi[u] = -1;
                    if (task < tasksA) {
                        jc_pair(task / np, step, m, pi[u], qi[u]);
                        jc_pair(task % np, step, m, ri[u], si[u]);
                        const bool qv = qi[u] < b, sv = si[u] < b;
                        b00[u] = __ldcg(&Ao[(size_t)pi[u] * ld + ri[u]]);
                        b01[u] = sv ? __ldcg(&Ao[(size_t)pi[u] * ld + si[u]]) : 0.0;
                        b10[u] = qv ? __ldcg(&Ao[(size_t)qi[u] * ld + ri[u]]) : 0.0;
                        b11[u] = (qv && sv) ? __ldcg(&Ao[(size_t)qi[u] * ld + si[u]]) : 0.0;
                    }
                }
                if (!synced) { __syncthreads(); synced = true; }   // rotations of this step are in shared memory
#pragma unroll
                for (int u = 0; u < JC_TA; u++) {
                    if (pi[u] < 0) continue;
                    const int task = t0 + u * GT;
                    const int I = task / np, J = task % np;
                    const bool qv = qi[u] < b, sv = si[u] < b;
                    const double cI = s_c[cur][I], sI = s_s[cur][I], cJ = s_c[cur][J], sJ = s_s[cur][J];
                    // T = B * J_J ; B' = J_I^T * T ; J = [[c, s], [-s, c]]
                    const double t00 = cJ * b00[u] - sJ * b01[u], t01 = sJ * b00[u] + cJ * b01[u];
                    const double t10 = cJ * b10[u] - sJ * b11[u], t11 = sJ * b10[u] + cJ * b11[u];
                    double n00 = cI * t00 - sI * t10, n01 = cI * t01 - sI * t11;
                    double n10 = sI * t00 + cI * t10, n11 = sI * t01 + cI * t11;
                    if (I == J && qv && s_t[I] != 0.0) {
                        const double tt = s_t[I];
                        n00 = b00[u] - tt * b01[u]; n11 = b11[u] + tt * b01[u]; n01 = 0.0; n10 = 0.0;
                    }
                    An[(size_t)pi[u] * ld + ri[u]] = n00;
                    if (sv) An[(size_t)pi[u] * ld + si[u]] = n01;
                    if (qv) An[(size_t)qi[u] * ld + ri[u]] = n10;
                    if (qv && sv) An[(size_t)qi[u] * ld + si[u]] = n11;
                }
            }
            cluster.sync();     // release/acquire at cluster scope; also invalidates L1
            double *tA = Ao; Ao = An; An = tA;
            have_prev = true;
            cur ^= 1;
        }
        sweeps++;
        // every CTA saw the same rotations, so the maximum is identical in all of them
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) offmax = fmax(offmax, __shfl_xor_sync(0xffffffffu, offmax, o));
        if (lane == 0) s_red[wid] = offmax;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < JC_THREADS / 32; w++) v = fmax(v, s_red[w]);
            s_off = v;
        }
        __syncthreads();
        converged = s_off <= tolabs;
        __syncthreads();
    }
    if (have_prev) {            // the last step's rotations
        apply_v(cur ^ 1);
        double *tV = Vo; Vo = Vn; Vn = tV;
    }
    if (gtid == 0) {
        info[0] = sweeps;
        info[1] = (Ao == A0) ? 0 : 1;     // which buffer holds the diagonalised matrix
        info[2] = converged ? 1 : 0;
        info[3] = (Vo == V0) ? 0 : 1;     // which buffer holds the eigenvectors
    }
}

// eigenvalues = diag(A) sorted descending (ties: lower index first); columns of V permuted to match.
// Vs may alias neither V nor A.
__global__ void sort_eig_kernel(const double *__restrict__ A, const double *__restrict__ V, int b, int ld,
                                double *__restrict__ w, double *__restrict__ Vs, int lds, int ncols_out) {
    extern __shared__ int s_rank[];
    for (int i = threadIdx.x; i < b; i += blockDim.x) {
        const double wi = A[(size_t)i * ld + i];
        int rank = 0;
        for (int j = 0; j < b; j++) {
            const double wj = A[(size_t)j * ld + j];
            rank += (wj > wi) || (wj == wi && j < i);
        }
        s_rank[i] = rank;
        if (blockIdx.x == 0) w[rank] = wi;
    }
    __syncthreads();
    for (int r = blockIdx.x; r < b; r += gridDim.x)
        for (int c = threadIdx.x; c < b; c += blockDim.x) {
            const int dst = s_rank[c];
            if (dst < ncols_out) Vs[(size_t)r * lds + dst] = V[(size_t)r * ld + c];
        }
}

// Eigen-decomposition of the symmetric b x b matrix in A (row-major, ld).  A is destroyed.
// On return w[0..b) holds the eigenvalues in descending order and Vs (b x lds) the matching
// eigenvectors in its first ncols_out columns.
int tp_jacobi(tp_ctx *ctx, double *A, int b, int ld, double *w, double *Vs, int lds, int ncols_out, int *sweeps_out,
              double tol) {
    TP_ARG(b >= 1 && (b + 1) / 2 <= JC_MAXPAIRS, "tp_jacobi: matrix too large for the cluster Jacobi solver");
    cudaStream_t st = ctx->stream;
    const size_t mat = (size_t)b * ld * sizeof(double);
    TP_TRY(ctx->Jt.reserve(3 * mat + 64));
    double *A1 = ctx->Jt.as<double>(), *V0 = A1 + (size_t)b * ld, *V1 = V0 + (size_t)b * ld;
    int *info = (int *)(V1 + (size_t)b * ld);
    tp_prof_begin(ctx, PC_JACOBI);
    jacobi_kernel<<<JC_CLUSTER, JC_THREADS, 0, st>>>(A, A1, V0, V1, b, ld, 40, tol, info);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    TP_TRY(tp_pin_reserve(ctx, 64));
    int *h = (int *)ctx->pin;
    TP_CUDA(cudaMemcpyAsync(h, info, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    TP_CUDA(cudaStreamSynchronize(st));
    if (sweeps_out) *sweeps_out = h[0];
    if (getenv("TADPOLE_DEBUG")) fprintf(stderr, "[tadpole] jacobi b=%d tol=%.1e sweeps=%d converged=%d\n", b, tol, h[0], h[2]);
    if (!h[2]) {
        tp_set_error("tp_jacobi: no convergence after %d sweeps (b = %d)", h[0], b);
        return TP_ERR_NOCONV;
    }
    const double *Af = h[1] ? A1 : A, *Vf = h[3] ? V1 : V0;
    int grid = b < 4 * ctx->sm_count ? b : 4 * ctx->sm_count;
    sort_eig_kernel<<<grid, 256, (size_t)b * sizeof(int), st>>>(Af, Vf, b, ld, w, Vs, lds, ncols_out);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}

// ------------------------------------------------------------------------------------------------
// Cholesky QR support: G = L L^T (b x b, SPD, L2 resident), then Linv = L^-1, both by one CTA.
// Y <- Y Linv^T orthonormalises the block (the GEMM is done by the caller).  info[0] = 1 when a
// pivot is not positive (block numerically rank deficient): the caller falls back to the
// eigen-decomposition based orthonormalisation.
// ------------------------------------------------------------------------------------------------
#define CH_THREADS 1024
__global__ void __launch_bounds__(CH_THREADS, 1)
chol_inv_kernel(double *__restrict__ G, double *__restrict__ Linv, int b, int ld, int *__restrict__ info) {
    extern __shared__ double s_col[];          // current column of L, b doubles
    __shared__ double s_d;
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int j = 0; j < b; j++) {
        if (tid == 0) {
            const double d = G[(size_t)j * ld + j];
            if (!(d > 0.0)) s_bad = 1;
            s_d = sqrt(d > 0.0 ? d : 1.0);
            G[(size_t)j * ld + j] = s_d;
        }
        __syncthreads();
        if (s_bad) break;
        const double inv = 1.0 / s_d;
        for (int i = j + 1 + tid; i < b; i += CH_THREADS) {
            const double v = G[(size_t)i * ld + j] * inv;
            G[(size_t)i * ld + j] = v;
            s_col[i] = v;
        }
        __syncthreads();
        // trailing lower triangle: G[i][k] -= L[i][j] L[k][j], j < k <= i; rows split over warps
        const int rem = b - j - 1;
        const int lane = tid & 31, wid = tid >> 5;
        for (int r = wid; r < rem; r += CH_THREADS / 32) {
            const int i = j + 1 + r;
            const double li = s_col[i];
            for (int k = j + 1 + lane; k <= i; k += 32) G[(size_t)i * ld + k] -= li * s_col[k];
        }
        __syncthreads();
    }
    if (s_bad) { if (tid == 0) info[0] = 1; return; }
    if (tid == 0) info[0] = 0;
    // Linv by forward substitution, one column per group of 4 lanes; loops are uniform across the
    // warp so that the shuffles are always executed by all lanes
    const int grp = tid >> 2, sub = tid & 3;
    for (int cbase = 0; cbase < b; cbase += CH_THREADS / 4) {
        const int c = cbase + grp;
        const bool cv = c < b;
        for (int i = 0; i < b; i++) {
            double s = 0.0;
            if (cv && i > c)
                for (int jj = c + sub; jj < i; jj += 4) s += G[(size_t)i * ld + jj] * Linv[(size_t)jj * ld + c];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (cv && sub == 0) {
                double x = 0.0;
                if (i >= c) x = ((i == c ? 1.0 : 0.0) - s) / G[(size_t)i * ld + i];
                Linv[(size_t)i * ld + c] = x;
            }
            __syncwarp();
        }
    }
}

// G (b x b, ld) is overwritten by its Cholesky factor; Linv receives L^-1.  *bad_out = 1 when G is
// not numerically positive definite.
int tp_chol_inv(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *bad_out) {
    cudaStream_t st = ctx->stream;
    TP_TRY(ctx->Jt.reserve(64));
    int *info = ctx->Jt.as<int>();
    TP_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(b * sizeof(double))));
    tp_prof_begin(ctx, PC_JACOBI);
    chol_inv_kernel<<<1, CH_THREADS, b * sizeof(double), st>>>(G, Linv, b, ld, info);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    TP_TRY(tp_pin_reserve(ctx, 64));
    int *h = (int *)ctx->pin;
    TP_CUDA(cudaMemcpyAsync(h, info, sizeof(int), cudaMemcpyDeviceToHost, st));
    TP_CUDA(cudaStreamSynchronize(st));
    *bad_out = h[0];
    return TP_OK;
}
