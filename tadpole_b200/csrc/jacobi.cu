// jacobi.cu -- dense symmetric eigensolver for the small problems of stage 3 (prcomp).
//
// Parallel cyclic Jacobi.  One thread-block cluster of 8 CTAs (8192 threads) owns the whole b x b
// matrix (L2 resident, b <= ~600).  A sweep is b-1 steps of a round-robin tournament; in each step
// the b/2 index pairs are disjoint, so all rotations are computed from the old matrix and applied
// at once: every 2 x 2 block (I, J) of the pair-permuted matrix becomes J_I^T * block * J_J,
// independently of every other block.  Old and new matrices ping-pong between two buffers, so
// the only synchronisation is one hardware cluster barrier per step.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define JC_CLUSTER 8
#define JC_THREADS 1024
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

__global__ void __cluster_dims__(JC_CLUSTER, 1, 1) __launch_bounds__(JC_THREADS, 1)
jacobi_kernel(double *A0, double *A1, double *V0, double *V1,
              int b, int ld, int max_sweeps, double tol, int *info) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_c[JC_MAXPAIRS], s_s[JC_MAXPAIRS], s_t[JC_MAXPAIRS];
    __shared__ int s_p[JC_MAXPAIRS], s_q[JC_MAXPAIRS];
    __shared__ double s_red[JC_THREADS / 32];
    __shared__ double s_anorm;
    __shared__ unsigned long long s_off;   // max |a_pq| of the sweep, as the bit pattern of a non-negative double
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
    int sweeps = 0;
    bool converged = (b < 2);
    while (!converged && sweeps < max_sweeps) {
        if (tid == 0) s_off = 0ULL;
        for (int step = 0; step < m - 1; step++) {
            // ---- phase A: every CTA computes all rotations of this step from the old matrix ----
            if (tid < np) {
                int p, q;
                jc_pair(tid, step, m, p, q);
                double c = 1.0, s = 0.0, t = 0.0;
                if (q < b) {
                    const double app = __ldcg(&Ao[(size_t)p * ld + p]), aqq = __ldcg(&Ao[(size_t)q * ld + q]), apq = __ldcg(&Ao[(size_t)p * ld + q]);
                    const double aoff = fabs(apq);
                    atomicMax(&s_off, (unsigned long long)__double_as_longlong(aoff));
                    if (aoff > 1e-300 && aoff > 1e-18 * sqrt(fabs(app) * fabs(aqq))) {
                        const double tau = (aqq - app) / (2.0 * apq);
                        t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                s_c[tid] = c; s_s[tid] = s; s_t[tid] = t; s_p[tid] = p; s_q[tid] = q;
            }
            __syncthreads();
            // ---- phase B: all 2x2 blocks of A, then all row-pairs of V ---------------------------
            const int tasksA = np * np, tasksV = b * np;
            for (int task = gtid; task < tasksA + tasksV; task += GT) {
                if (task < tasksA) {
                    const int I = task / np, J = task % np;
                    const int p = s_p[I], q = s_q[I], r = s_p[J], s2 = s_q[J];
                    const bool qv = q < b, sv = s2 < b;
                    const double cI = s_c[I], sI = s_s[I], cJ = s_c[J], sJ = s_s[J];
                    const double b00 = __ldcg(&Ao[(size_t)p * ld + r]);
                    const double b01 = sv ? __ldcg(&Ao[(size_t)p * ld + s2]) : 0.0;
                    const double b10 = qv ? __ldcg(&Ao[(size_t)q * ld + r]) : 0.0;
                    const double b11 = (qv && sv) ? __ldcg(&Ao[(size_t)q * ld + s2]) : 0.0;
                    // T = B * J_J ; B' = J_I^T * T ; J = [[c, s], [-s, c]]
                    const double t00 = cJ * b00 - sJ * b01, t01 = sJ * b00 + cJ * b01;
                    const double t10 = cJ * b10 - sJ * b11, t11 = sJ * b10 + cJ * b11;
                    double n00 = cI * t00 - sI * t10, n01 = cI * t01 - sI * t11;
                    double n10 = sI * t00 + cI * t10, n11 = sI * t01 + cI * t11;
                    if (I == J && qv && s_t[I] != 0.0) {
                        const double tt = s_t[I];
                        n00 = b00 - tt * b01; n11 = b11 + tt * b01; n01 = 0.0; n10 = 0.0;
                    }
                    An[(size_t)p * ld + r] = n00;
                    if (sv) An[(size_t)p * ld + s2] = n01;
                    if (qv) An[(size_t)q * ld + r] = n10;
                    if (qv && sv) An[(size_t)q * ld + s2] = n11;
                } else {
                    const int tv = task - tasksA;
                    const int row = tv / np, J = tv % np;
                    const int r = s_p[J], s2 = s_q[J];
                    const double cJ = s_c[J], sJ = s_s[J];
                    const double v0 = __ldcg(&Vo[(size_t)row * ld + r]);
                    if (s2 < b) {
                        const double v1 = __ldcg(&Vo[(size_t)row * ld + s2]);
                        Vn[(size_t)row * ld + r] = cJ * v0 - sJ * v1;
                        Vn[(size_t)row * ld + s2] = sJ * v0 + cJ * v1;
                    } else {
                        Vn[(size_t)row * ld + r] = v0;
                    }
                }
            }
            cluster.sync();     // release/acquire at cluster scope; also invalidates L1
            double *tA = Ao; Ao = An; An = tA;
            double *tV = Vo; Vo = Vn; Vn = tV;
        }
        sweeps++;
        // every CTA saw the same rotations, so s_off is identical in all of them
        __syncthreads();
        converged = __longlong_as_double((long long)s_off) <= tolabs;
        __syncthreads();
    }
    if (gtid == 0) {
        info[0] = sweeps;
        info[1] = (Ao == A0) ? 0 : 1;     // which buffer pair holds the result
        info[2] = converged ? 1 : 0;
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
int tp_jacobi(tp_ctx *ctx, double *A, int b, int ld, double *w, double *Vs, int lds, int ncols_out, int *sweeps_out) {
    TP_ARG(b >= 1 && (b + 1) / 2 <= JC_MAXPAIRS, "tp_jacobi: matrix too large for the cluster Jacobi solver");
    cudaStream_t st = ctx->stream;
    const size_t mat = (size_t)b * ld * sizeof(double);
    TP_TRY(ctx->Jt.reserve(3 * mat + 64));
    double *A1 = ctx->Jt.as<double>(), *V0 = A1 + (size_t)b * ld, *V1 = V0 + (size_t)b * ld;
    int *info = (int *)(V1 + (size_t)b * ld);
    tp_prof_begin(ctx, PC_JACOBI);
    jacobi_kernel<<<JC_CLUSTER, JC_THREADS, 0, st>>>(A, A1, V0, V1, b, ld, 40, 1e-14, info);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    TP_TRY(tp_pin_reserve(ctx, 64));
    int *h = (int *)ctx->pin;
    TP_CUDA(cudaMemcpyAsync(h, info, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    TP_CUDA(cudaStreamSynchronize(st));
    if (sweeps_out) *sweeps_out = h[0];
    if (!h[2]) {
        tp_set_error("tp_jacobi: no convergence after %d sweeps (b = %d)", h[0], b);
        return TP_ERR_NOCONV;
    }
    const double *Af = h[1] ? A1 : A, *Vf = h[1] ? V1 : V0;
    int grid = b < 4 * ctx->sm_count ? b : 4 * ctx->sm_count;
    sort_eig_kernel<<<grid, 256, (size_t)b * sizeof(int), st>>>(Af, Vf, b, ld, w, Vs, lds, ncols_out);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    return TP_OK;
}
