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
#define JC_THREADS 1024
#define JC_MAXPAIRS 512

// round-robin tournament over m (even) players: player m-1 is fixed, the others rotate.
// slot 0 pairs (m-1, step); slot J pairs ((step+J) mod (m-1), (step-J) mod (m-1)).
__device__ __forceinline__ void jc_pair(int slot, int step, int m, int &p, int &q) {
    int a, b;
    if (slot == 0) { a = m - 1; b = step; }
    else {
        a = (step + slot) % (m - 1);
        b = (step - slot + (m - 1)) % (m - 1);
    }
    p = min(a, b); q = max(a, b);
}

__device__ __forceinline__ double fast_rcp(double x) {       // ~1e-11 relative: MUFU + one Newton step
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {     // full precision after two Newton steps
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * fma(-0.5 * x * r, r, 1.5);
    r = r * fma(-0.5 * x * r, r, 1.5);
    return r * fma(-0.5 * x * r, r, 1.5);
}

// One step: (A) rotations of this step from the old matrix, (B) every 2 x 2 block (I, J) of the
// pair-permuted matrix becomes J_I^T * block * J_J.  The rotations (c, s) of every step are logged
// to global memory; the eigenvectors are accumulated afterwards by vapply_kernel, which is
// embarrassingly parallel over rows, so the serial loop here only carries the matrix itself.
// TA = 2 x 2 blocks per thread.
template <int TA>
__global__ void __cluster_dims__(JC_CLUSTER, 1, 1) __launch_bounds__(JC_THREADS, 1)
jacobi_kernel(double *A0, double *A1, double2 *__restrict__ rotlog,
              int b, int ld, int max_sweeps, double tol, int *info) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_c[JC_MAXPAIRS], s_s[JC_MAXPAIRS];
    __shared__ short s_p[JC_MAXPAIRS], s_q[JC_MAXPAIRS];
    __shared__ double s_red[JC_THREADS / 32];
    __shared__ double s_anorm;
    __shared__ double s_off;               // max |a_pq| seen in the sweep
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int crank = cluster.block_rank();
    const int gtid = crank * JC_THREADS + tid;
    const int GT = JC_CLUSTER * JC_THREADS;
    const int m = (b + 1) & ~1;          // even number of players; index b (if any) is a bye
    const int np = m / 2;
    // this thread's blocks: task t = gtid + u * GT  <->  (I, J) = (t / np, t % np), fixed for the whole run
    int tI[TA], tJ[TA];
#pragma unroll
    for (int u = 0; u < TA; u++) {
        const int t = gtid + u * GT;
        tI[u] = t / np; tJ[u] = t % np;       // tI >= np marks an empty slot
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

    const double *Ao = A0;
    double *An = A1;
    int sweeps = 0;
    long logpos = 0;                       // step counter across sweeps
    bool converged = (b < 2);
    while (!converged && sweeps < max_sweeps) {
        double offmax = 0.0;               // per thread; reduced once per sweep (no atomics in the step loop)
        // pair of slot `tid`, advanced incrementally from step to step
        int pa = 0, pb = 0;
        if (tid < np) { if (tid == 0) { pa = m - 1; pb = 0; } else { pa = tid % (m - 1); pb = (m - 1 - tid) % (m - 1); } }
        for (int step = 0; step < m - 1; step++) {
            // ---- pair table of this step + loads for its rotations --------------------------------
            int p = 0, q = 0;
            double app = 0.0, aqq = 0.0, apq = 0.0;
            if (tid < np) {
                p = min(pa, pb); q = max(pa, pb);
                s_p[tid] = (short)p; s_q[tid] = (short)q;
                if (q < b) {
                    app = __ldcg(&Ao[p * ld + p]);
                    aqq = __ldcg(&Ao[q * ld + q]);
                    apq = __ldcg(&Ao[p * ld + q]);
                }
                // next step: every rotating player moves on by one
                if (tid == 0) pb = pb + 1;
                else { pa = (pa + 1 == m - 1) ? 0 : pa + 1; pb = (pb + 1 == m - 1) ? 0 : pb + 1; }
            }
            __syncthreads();
            // ---- (B) loads: they do not depend on the rotations -----------------------------------
            double b00[TA], b01[TA], b10[TA], b11[TA];
            int opr[TA], ops[TA], oqr[TA], oqs[TA];     // element offsets, -1 = bye
#pragma unroll
            for (int u = 0; u < TA; u++) {
                opr[u] = ops[u] = oqr[u] = oqs[u] = -1;
                if (tI[u] < np) {
                    const int pi = s_p[tI[u]], qi = s_q[tI[u]], ri = s_p[tJ[u]], si = s_q[tJ[u]];
                    opr[u] = pi * ld + ri;
                    if (si < b) ops[u] = pi * ld + si;
                    if (qi < b) oqr[u] = qi * ld + ri;
                    if (qi < b && si < b) oqs[u] = qi * ld + si;
                    b00[u] = __ldcg(Ao + opr[u]);
                    b01[u] = ops[u] >= 0 ? __ldcg(Ao + ops[u]) : 0.0;
                    b10[u] = oqr[u] >= 0 ? __ldcg(Ao + oqr[u]) : 0.0;
                    b11[u] = oqs[u] >= 0 ? __ldcg(Ao + oqs[u]) : 0.0;
                }
            }
            // ---- (A) rotations, published through shared memory and logged for vapply_kernel ------
            if (tid < np) {
                double c = 1.0, s = 0.0;
                if (q < b) {
                    const double aoff = fabs(apq);
                    offmax = fmax(offmax, aoff);
                    const double dd = aqq - app;
                    // rotate unless a_pq is negligible against both the diagonal gap and the diagonal
                    if (aoff > 1e-300 && aoff > 1e-20 * (fabs(dd) + fabs(app) + fabs(aqq))) {
                        // t = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = dd / (2 a_pq), written without
                        // the division:  t = 2 a_pq sgn / (|dd| + sqrt(dd^2 + 4 a_pq^2))
                        const double h = sqrt(fma(dd, dd, 4.0 * apq * apq));
                        double t = 2.0 * apq * fast_rcp(fabs(dd) + h);
                        if (dd < 0.0) t = -t;
                        // c^2 + s^2 = 1 to working precision whatever the accuracy of t
                        c = fast_rsqrt(fma(t, t, 1.0));
                        s = t * c;
                    }
                }
                s_c[tid] = c; s_s[tid] = s;
                if (crank == 0) rotlog[logpos * np + tid] = make_double2(c, s);
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < TA; u++) {
                if (tI[u] >= np) continue;
                const double cI = s_c[tI[u]], sI = s_s[tI[u]], cJ = s_c[tJ[u]], sJ = s_s[tJ[u]];
                // T = B * J_J ; B' = J_I^T * T ; J = [[c, s], [-s, c]]
                const double t00 = cJ * b00[u] - sJ * b01[u], t01 = sJ * b00[u] + cJ * b01[u];
                const double t10 = cJ * b10[u] - sJ * b11[u], t11 = sJ * b10[u] + cJ * b11[u];
                An[opr[u]] = cI * t00 - sI * t10;
                if (ops[u] >= 0) An[ops[u]] = cI * t01 - sI * t11;
                if (oqr[u] >= 0) An[oqr[u]] = sI * t00 + cI * t10;
                if (oqs[u] >= 0) An[oqs[u]] = sI * t01 + cI * t11;
            }
            cluster.sync();     // release/acquire at cluster scope
            const double *tA = Ao; Ao = An; An = (double *)tA;
            logpos++;
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
    if (gtid == 0) {
        info[0] = sweeps;
        info[1] = (Ao == A0) ? 0 : 1;     // which buffer holds the diagonalised matrix
        info[2] = converged ? 1 : 0;
    }
}

// V = J_1 J_2 ... J_S (all logged steps applied to the identity), two rows of V per CTA in shared
// memory; thread J owns pair J of every step, so a step is two loads, four FMAs and a barrier.
#define VA_ROWS 2
__global__ void __launch_bounds__(JC_MAXPAIRS)
vapply_kernel(const double2 *__restrict__ rotlog, const int *__restrict__ info, int b, int ld, double *__restrict__ V) {
    extern __shared__ double s_row[];        // VA_ROWS x (b + 1)
    const int tid = threadIdx.x;
    const int m = (b + 1) & ~1, np = m / 2;
    const int r0 = blockIdx.x * VA_ROWS;
    const long nsteps = (long)info[0] * (m - 1);
    for (int idx = tid; idx < VA_ROWS * (b + 1); idx += blockDim.x) {
        const int rr = idx / (b + 1), c = idx % (b + 1);
        s_row[idx] = (r0 + rr == c) ? 1.0 : 0.0;
    }
    int pa = 0, pb = 0;
    if (tid < np) { if (tid == 0) { pa = m - 1; pb = 0; } else { pa = tid % (m - 1); pb = (m - 1 - tid) % (m - 1); } }
    double2 cs = (tid < np && nsteps > 0) ? rotlog[tid] : make_double2(1.0, 0.0);
    __syncthreads();
    int step = 0;
    for (long g = 0; g < nsteps; g++) {
        double2 nxt = (tid < np && g + 1 < nsteps) ? rotlog[(g + 1) * np + tid] : make_double2(1.0, 0.0);
        if (tid < np) {
            const int p = min(pa, pb), q = max(pa, pb);      // q == b (bye) lands in the padding column
#pragma unroll
            for (int rr = 0; rr < VA_ROWS; rr++) {
                double *row = s_row + rr * (b + 1);
                const double v0 = row[p], v1 = row[q];
                row[p] = cs.x * v0 - cs.y * v1;
                row[q] = cs.y * v0 + cs.x * v1;
            }
            if (tid == 0) pb = pb + 1;
            else { pa = (pa + 1 == m - 1) ? 0 : pa + 1; pb = (pb + 1 == m - 1) ? 0 : pb + 1; }
        }
        if (++step == m - 1) {      // new sweep: the schedule starts over
            step = 0;
            if (tid < np) { if (tid == 0) { pa = m - 1; pb = 0; } else { pa = tid % (m - 1); pb = (m - 1 - tid) % (m - 1); } }
        }
        cs = nxt;
        __syncthreads();
    }
    for (int idx = tid; idx < VA_ROWS * b; idx += blockDim.x) {
        const int rr = idx / b, c = idx % b;
        if (r0 + rr < b) V[(size_t)(r0 + rr) * ld + c] = s_row[rr * (b + 1) + c];
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

int tp_osj(tp_ctx *ctx, double *T, int b, int ld, double *w, double *Vs, int lds, int ncols_out, int *sweeps_out,
           double tol, double predict);

// Eigen-decomposition of the symmetric b x b matrix in A (row-major, ld).  A is destroyed.
// On return w[0..b) holds the eigenvalues in descending order and Vs (b x lds) the matching
// eigenvectors in its first ncols_out columns.
int tp_jacobi(tp_ctx *ctx, double *A, int b, int ld, double *w, double *Vs, int lds, int ncols_out, int *sweeps_out,
              double tol, double predict) {
    TP_ARG(b >= 1 && (b + 1) / 2 <= JC_MAXPAIRS, "tp_jacobi: matrix too large for the cluster Jacobi solver");
    if (b >= 2 && b <= 384 && !getenv("TADPOLE_TWOSIDED"))
        return tp_osj(ctx, A, b, ld, w, Vs, lds, ncols_out, sweeps_out, tol, predict);
    cudaStream_t st = ctx->stream;
    const int max_sweeps = 40;
    const int m = (b + 1) & ~1, np = m / 2;
    const size_t mat = (size_t)b * ld * sizeof(double);
    const size_t logbytes = (size_t)max_sweeps * (m - 1) * np * sizeof(double2);
    TP_TRY(ctx->Jt.reserve(2 * mat + logbytes + 64));
    double *A1 = ctx->Jt.as<double>(), *V = A1 + (size_t)b * ld;
    double2 *rotlog = (double2 *)(V + (size_t)b * ld);
    int *info = (int *)((char *)rotlog + logbytes);
    const int GT = JC_CLUSTER * JC_THREADS;
    const int ta = (np * np + GT - 1) / GT;
    tp_prof_begin(ctx, PC_JACOBI);
    if (ta <= 1) jacobi_kernel<1><<<JC_CLUSTER, JC_THREADS, 0, st>>>(A, A1, rotlog, b, ld, max_sweeps, tol, info);
    else if (ta <= 2) jacobi_kernel<2><<<JC_CLUSTER, JC_THREADS, 0, st>>>(A, A1, rotlog, b, ld, max_sweeps, tol, info);
    else if (ta <= 4) jacobi_kernel<4><<<JC_CLUSTER, JC_THREADS, 0, st>>>(A, A1, rotlog, b, ld, max_sweeps, tol, info);
    else if (ta <= 8) jacobi_kernel<8><<<JC_CLUSTER, JC_THREADS, 0, st>>>(A, A1, rotlog, b, ld, max_sweeps, tol, info);
    else if (ta <= 16) jacobi_kernel<16><<<JC_CLUSTER, JC_THREADS, 0, st>>>(A, A1, rotlog, b, ld, max_sweeps, tol, info);
    else { tp_set_error("tp_jacobi: b = %d needs more than 16 blocks per thread", b); return TP_ERR_ARG; }
    const size_t vsm = (size_t)VA_ROWS * (b + 1) * sizeof(double);
    vapply_kernel<<<(b + VA_ROWS - 1) / VA_ROWS, JC_MAXPAIRS, vsm, st>>>(rotlog, info, b, ld, V);
    tp_prof_end(ctx);
    ctx->launches += 2;
    TP_CUDA(cudaGetLastError());
    TP_TRY(tp_pin_reserve(ctx, 64));
    int *h = (int *)ctx->pin;
    TP_CUDA(cudaMemcpyAsync(h, info, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    TP_CUDA(tp_stream_sync(ctx));
    if (sweeps_out) *sweeps_out = h[0];
    if (getenv("TADPOLE_DEBUG")) fprintf(stderr, "[tadpole] jacobi b=%d tol=%.1e sweeps=%d converged=%d\n", b, tol, h[0], h[2]);
    if (!h[2]) {
        tp_set_error("tp_jacobi: no convergence after %d sweeps (b = %d)", h[0], b);
        return TP_ERR_NOCONV;
    }
    const double *Af = h[1] ? A1 : A;
    int grid = b < 4 * ctx->sm_count ? b : 4 * ctx->sm_count;
    sort_eig_kernel<<<grid, 256, (size_t)b * sizeof(int), st>>>(Af, V, b, ld, w, Vs, lds, ncols_out);
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
#define CH_PW 32                 // panel width
__global__ void __launch_bounds__(CH_THREADS, 1)
chol_inv_kernel(double *G, double *Linv, int b, int ld, int *info, int factor_only) {
    extern __shared__ double s_pan[];          // panel rows [j0, b) x CH_PW columns, pitch CH_PW + 1
    __shared__ double s_d;
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    constexpr int PP = CH_PW + 1;
    __shared__ double s_clamp;
    __shared__ double s_rd[CH_PW];             // reciprocal pivots of the current panel (0: column dropped)
    if (tid == 0) {
        s_bad = 0;
        double mx = 0.0;
        if (factor_only) for (int i = 0; i < b; i++) mx = fmax(mx, fabs(G[(size_t)i * ld + i]));
        s_clamp = mx * 1e-24;      // factor_only: semi-definite input allowed, tiny pivots are clamped
    }
    __syncthreads();
    // blocked right-looking Cholesky, lower triangle; the panel is factored in shared memory
    for (int j0 = 0; j0 < b && !s_bad; j0 += CH_PW) {
        const int w = min(CH_PW, b - j0), rows = b - j0;
        for (int idx = tid; idx < rows * w; idx += CH_THREADS) {
            const int r = idx / w, c = idx % w;
            s_pan[r * PP + c] = G[(size_t)(j0 + r) * ld + j0 + c];
        }
        __syncthreads();
        // (1) warp 0 factors the w x w diagonal block (lane = row), column by column
        if (tid < 32) {
            for (int c = 0; c < w; c++) {
                double d = s_pan[c * PP + c];
                // semi-definite input: a round-off level pivot keeps a tiny diagonal and a zero column
                const bool tiny = factor_only && !(d > s_clamp);
                if (tiny) d = s_clamp > 0.0 ? s_clamp : 1e-300;
                const bool ok = d > 0.0;
                if (!ok && tid == 0) s_bad = 1;
                const double piv = sqrt(ok ? d : 1.0);
                __syncwarp();
                if (tid == c) { s_pan[c * PP + c] = piv; s_rd[c] = tiny ? 0.0 : 1.0 / piv; }
                if (tid > c && tid < w) s_pan[tid * PP + c] *= tiny ? 0.0 : 1.0 / piv;
                __syncwarp();
                // row `tid` of the remaining block: P[tid][c2] -= P[tid][c] * P[c2][c], c < c2 <= tid
                if (tid > c && tid < w) {
                    const double lrc = s_pan[tid * PP + c];
                    for (int c2 = c + 1; c2 <= tid; c2++) s_pan[tid * PP + c2] -= lrc * s_pan[c2 * PP + c];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // (2) rows below the diagonal block: solve x L11^T = a, one thread per row
        for (int r = w + tid; r < rows; r += CH_THREADS) {
            double *row = s_pan + r * PP;
            for (int c = 0; c < w; c++) {
                double s = row[c];
                for (int t = 0; t < c; t++) s -= row[t] * s_pan[c * PP + t];
                row[c] = s * s_rd[c];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < rows * w; idx += CH_THREADS) {
            const int r = idx / w, c = idx % w;
            if (r >= c) G[(size_t)(j0 + r) * ld + j0 + c] = s_pan[r * PP + c];
        }
        // trailing update: G[i][k] -= sum_c P[i][c] P[k][c] for j0 + w <= k <= i < b
        const int t0 = w, tn = rows - w;       // panel-relative rows [t0, rows)
        for (int idx = tid; idx < tn * tn; idx += CH_THREADS) {
            const int ri = t0 + idx / tn, rk = t0 + idx % tn;
            if (rk > ri) continue;
            double acc = 0.0;
#pragma unroll 8
            for (int c = 0; c < w; c++) acc += s_pan[ri * PP + c] * s_pan[rk * PP + c];
            G[(size_t)(j0 + ri) * ld + j0 + rk] -= acc;
        }
        __syncthreads();
    }
    if (s_bad) { if (tid == 0) info[0] = 1; return; }       // sticky status word: never cleared here
    if (factor_only) return;
    __syncthreads();
    // ---- Linv = L^-1, blocked: 32 x 32 diagonal blocks by forward substitution (one warp each, one
    // column per lane), then every block column j independently: X_ij = -X_ii * sum_{k=j}^{i-1} L_ik X_kj
    const int nb = (b + CH_PW - 1) / CH_PW;
    const int lane = tid & 31, wid = tid >> 5;
    for (int idx = tid; idx < b * b; idx += CH_THREADS) {       // clear (upper triangle must be zero)
        const int r = idx / b, c = idx % b;
        if (c > r || (r / CH_PW) != (c / CH_PW)) Linv[(size_t)r * ld + c] = 0.0;
    }
    for (int d = wid; d < nb; d += CH_THREADS / 32) {
        const int o = d * CH_PW, w = min(CH_PW, b - o);
        double *Ld = s_pan + (size_t)(d % (CH_THREADS / 32)) * CH_PW * PP;     // this warp's copy of L_dd
        for (int r = 0; r < w; r++) Ld[r * PP + lane] = (lane < w) ? G[(size_t)(o + r) * ld + o + lane] : 0.0;
        __syncwarp();
        // lane c: column c of the inverse
        double x[CH_PW];
#pragma unroll
        for (int i = 0; i < CH_PW; i++) {
            double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int jj = 0; jj < CH_PW; jj++) if (jj < i) s -= Ld[i * PP + jj] * x[jj];
            x[i] = (i < w && lane < w && i >= lane) ? s / Ld[i * PP + i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < CH_PW; i++) if (i < w && lane < w) Linv[(size_t)(o + i) * ld + o + lane] = x[i];
        __syncwarp();
    }
    __syncthreads();
    // block columns: group g = tid / tpg owns block column g; all groups walk i together
    const int tpg = max(32, (CH_THREADS / max(nb, 1)) & ~31);      // threads per group (multiple of 32)
    const int ngroups = CH_THREADS / tpg;
    double *S = s_pan;                                            // per group 32 x PP scratch
    for (int jb0 = 0; jb0 < nb; jb0 += ngroups) {
        const int g = tid / tpg, gt = tid % tpg;
        const int jb = jb0 + g;
        double *Sg = S + (size_t)g * CH_PW * PP;
        for (int ib = jb0 + 1; ib < nb; ib++) {
            const bool act = g < ngroups && jb < nb && ib > jb;
            if (act) {
                for (int e = gt; e < CH_PW * CH_PW; e += tpg) {
                    const int r = e / CH_PW, c = e % CH_PW;
                    const int gr = ib * CH_PW + r, gc = jb * CH_PW + c;
                    double acc = 0.0;
                    if (gr < b && gc < b) {
                        const double *Lrow = G + (size_t)gr * ld;
                        for (int t = jb * CH_PW + c; t < ib * CH_PW; t++)     // X_kj is lower triangular: t >= gc
                            acc += Lrow[t] * Linv[(size_t)t * ld + gc];   // same CTA wrote it: visible after the barrier
                    }
                    Sg[r * PP + c] = acc;
                }
            }
            __syncthreads();
            if (act) {
                for (int e = gt; e < CH_PW * CH_PW; e += tpg) {
                    const int r = e / CH_PW, c = e % CH_PW;
                    const int gr = ib * CH_PW + r, gc = jb * CH_PW + c;
                    if (gr < b && gc < b) {
                        double acc = 0.0;
                        for (int t = 0; t <= r; t++)                        // X_ii lower triangular
                            acc += Linv[(size_t)gr * ld + ib * CH_PW + t] * Sg[t * PP + c];
                        Linv[(size_t)gr * ld + gc] = -acc;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// single-CTA launcher (b > 256; the cluster kernel in cholinv.cu handles the common sizes)
int tp_chol_inv_1cta(tp_ctx *ctx, double *G, double *Linv, int b, int ld, int *info, int factor_only) {
    cudaStream_t st = ctx->stream;
    size_t chsm = (size_t)b * (CH_PW + 1) * sizeof(double);
    const int nbw = (b + CH_PW - 1) / CH_PW;
    const size_t chsm2 = (size_t)(nbw < 32 ? nbw : 32) * CH_PW * (CH_PW + 1) * sizeof(double);
    if (!factor_only && chsm2 > chsm) chsm = chsm2;
    TP_ARG(chsm <= (size_t)ctx->max_smem_optin, "tp_chol_inv: block too wide for the shared-memory panel");
    TP_CUDA(tp_optin_smem(chol_inv_kernel, ctx));
    tp_prof_begin(ctx, PC_CHOL);
    chol_inv_kernel<<<1, CH_THREADS, chsm, st>>>(G, Linv, b, ld, info, factor_only);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}
