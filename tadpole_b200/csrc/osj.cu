// osj.cu -- one-sided (Hestenes) block Jacobi for the b x b Rayleigh-Ritz problems of stage 3, b <= 256.
//
// T = L L^T (Cholesky).  Right rotations W that make the columns of L orthogonal give L W = U S, and
// then T = U S^2 U^T: the normalised columns of the rotated factor ARE the eigenvectors of T and their
// squared norms its eigenvalues (Veselic-Hari), so no rotation product has to be accumulated.  One-sided Jacobi only ever mixes two COLUMNS, so a CTA that owns
// a set of columns can work on them in shared memory without talking to anyone.  The 16 H columns are cut
// into 16 half-panels of H columns; a cluster of 8 CTAs runs a block round-robin: in each of 15 block
// steps CTA r loads two half-panels of the factor (from L2) into shared memory, orthogonalises every
// column pair across them in H sub-steps of H disjoint pairs (one warp per pair, shuffle reductions, one
// CTA barrier per sub-step), writes them back, and the cluster barrier ends the block step.  A sweep is
// therefore 15 cluster barriers instead of b-1, and the rotation steps themselves run at shared-memory
// latency.  Working on the Cholesky factor instead of T keeps the accuracy of small eigenvalues.
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define OSJ_CLUSTER 8
#define OSJ_THREADS 256             // 16 half-warps: one column pair per half-warp and sub-step
#define OSJ_NHP 16              // half-panels
#define OSJ_MAXH 24             // columns per half-panel
#define OSJ_MAX_B 384            // 2 * 24 columns x 384 rows x 8 B = 147 KB of shared memory

__device__ __forceinline__ void osj_pair(int slot, int step, int m, int &p, int &q) {
    int a, b;
    if (slot == 0) { a = m - 1; b = step; }
    else { a = (step + slot) % (m - 1); b = (step - slot + (m - 1)) % (m - 1); }
    p = min(a, b); q = max(a, b);
}

__device__ __forceinline__ double osj_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double osj_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * fma(-0.5 * x * r, r, 1.5);
    r = r * fma(-0.5 * x * r, r, 1.5);
    return r * fma(-0.5 * x * r, r, 1.5);
}

// column-major copy of the Cholesky factor: Ac[c * BP + r] = L[r][c] for c <= r < b, zero elsewhere
__global__ void osj_init_kernel(const double *__restrict__ L, int b, int ld, int BP, int NC,
                                double *__restrict__ Ac) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)NC * BP) return;
    const int c = (int)(idx / BP), r = (int)(idx % BP);
    Ac[idx] = (c < b && r < b && r >= c) ? L[(size_t)r * ld + c] : 0.0;
}

// shared memory through 32-bit shared-window addresses (no generic-address arithmetic in the chain)
__device__ __forceinline__ double osj_lds(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void osj_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// (c, s) of the rotation that makes two columns with squared norms al, be and inner product g orthogonal; false = leave the
// pair alone (a zero column, or already orthogonal to working precision).  rmax2 collects the squared scaled cosines.
__device__ __forceinline__ bool osj_cs(double g, double al, double be, double skip2, double &rmax2, double &c, double &sn) {
    const double ab = al * be, gg = g * g;
    if (!(ab > 0.0)) return false;
    // squared scaled cosine g^2 / (al be): the convergence measure (off the critical path; a crude reciprocal is enough)
    double rab;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rab) : "d"(ab));
    rmax2 = fmax(rmax2, gg * rab);
    if (!(gg > skip2 * ab)) return false;
    // Tangent t = sgn(dd) 2 g / (|dd| + sqrt(dd^2 + 4 g^2)), dd = be - al.  t only steers convergence (an error of 2^-20 in
    // t leaves 2^-20 of the pair's cosine behind, far below what the other rotations of the sweep put back), so it is
    // computed in FP32 on operands scaled into range by a power of two; (c, s) is then normalised to c^2 + s^2 = 1 in
    // full precision: two Newton steps take the FP32 reciprocal square root to 2^-80.
    const double dd = be - al;
    const int hm = max(__double2hiint(dd) & 0x7ff00000, __double2hiint(g) & 0x7ff00000);
    const double scl = __hiloint2double(0x7fd00000 - hm, 0);            // 2^(1021 - E): the larger operand lands in [1/4, 1/2)
    const float df = __double2float_rn(dd * scl), gf = __double2float_rn(g * scl);
    const float hh = fmaf(df, df, 4.0f * gf * gf);                      // >= 1/16: the larger operand is >= 1/4
    const float hf = hh * rsqrtf(hh);
    float tf = __fdividef(2.0f * gf, fabsf(df) + hf);
    if (df < 0.0f) tf = -tf;
    const float cf = rsqrtf(fmaf(tf, tf, 1.0f));
    const double t = (double)tf;
    const double htt = 0.5 * fma(t, t, 1.0);
    c = (double)cf;
    c = c * fma(-htt * c, c, 1.5);
    c = c * fma(-htt * c, c, 1.5);
    sn = t * c;
    return true;
}

// Rotate columns ix and iy of the shared-memory panel: one HALF-warp (16 lanes, RL rows per lane), so that a warp works on
// two pairs at once: the scalar chain between the dot product and the rotation (the latency of a sub-step) is issued once
// for both, the reduction tree is four levels, and 8 warps per CTA (2 per scheduler) contend less for the FP64 pipe than
// 16 did.  hmask = the lanes of this half; the two halves of a warp may diverge.
template <int RL>
__device__ __forceinline__ void osj_rotate(unsigned sA, unsigned sN, int BP, int ix, int iy,
                                           int hl, unsigned hmask, double skip2, double &rmax2) {
    const unsigned ax = sA + 8u * (ix * BP + hl), ay = sA + 8u * (iy * BP + hl);
    double xa[RL], ya[RL];
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
    for (int u = 0; u < RL; u++) { xa[u] = osj_lds(ax + 128u * u); ya[u] = osj_lds(ay + 128u * u); }
    const double al = osj_lds(sN + 8u * ix), be = osj_lds(sN + 8u * iy);
#pragma unroll
    for (int u = 0; u < RL; u += 4) {
        g0 = fma(xa[u], ya[u], g0);
        if (u + 1 < RL) g1 = fma(xa[u + 1], ya[u + 1], g1);
        if (u + 2 < RL) g2 = fma(xa[u + 2], ya[u + 2], g2);
        if (u + 3 < RL) g3 = fma(xa[u + 3], ya[u + 3], g3);
    }
    double g = (g0 + g1) + (g2 + g3);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) g += __shfl_xor_sync(hmask, g, o);
    double c, sn;
    if (!osj_cs(g, al, be, skip2, rmax2, c, sn)) return;
#pragma unroll
    for (int u = 0; u < RL; u++) {
        osj_sts(ax + 128u * u, c * xa[u] - sn * ya[u]);
        osj_sts(ay + 128u * u, sn * xa[u] + c * ya[u]);
    }
    // new squared norms (the tangent is only approximately Jacobi's, so from c and s): c^2 al - 2 c s g + s^2 be and mirror
    if (hl == 0) {
        const double cc = c * c, ss = sn * sn, cs2 = 2.0 * c * sn * g;
        osj_sts(sN + 8u * ix, fma(cc, al, fma(ss, be, -cs2)));
        osj_sts(sN + 8u * iy, fma(ss, al, fma(cc, be, cs2)));
    }
}

// All pairs across the two half-panels, H sub-steps of H disjoint pairs (x_p, y_(p + st) mod H), H <= 16: half-warp p keeps its
// x column in REGISTERS over the H sub-steps and only the y columns travel through shared memory.  (With both columns
// round-tripping through shared memory every sub-step moved the whole 64 KB panel in and out: 128 KB at 128 B / clock =
// 1024 of the ~1850 cycles a sub-step took -- profiles/r02_osj_trace.md.)
template <int RL>
__device__ __forceinline__ void osj_cross_xreg(unsigned sA, unsigned sN, int BP, int H, int hwid, int hl, unsigned hmask,
                                               double skip2, double &rmax2) {
    const bool act = hwid < H;
    const unsigned ax = sA + 8u * (hwid * BP + hl);
    double xa[RL], al = 0.0;
    if (act) {
#pragma unroll
        for (int u = 0; u < RL; u++) xa[u] = osj_lds(ax + 128u * u);
        al = osj_lds(sN + 8u * hwid);
    }
    // (Splitting the ring into four groups of four pairs that synchronise among themselves only -- named barriers, so that one
    // group's shared-memory bursts overlap another's arithmetic -- changed nothing: a sub-step is one dependent chain, not a
    // contended pipe.)
    for (int st = 0; st < H; st++) {
        if (act) {
            int jj = hwid + st; if (jj >= H) jj -= H;
            const int iy = H + jj;
            const unsigned ay = sA + 8u * (iy * BP + hl);
            double ya[RL];
            double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
            for (int u = 0; u < RL; u++) ya[u] = osj_lds(ay + 128u * u);
            const double be = osj_lds(sN + 8u * iy);
#pragma unroll
            for (int u = 0; u < RL; u += 4) {
                g0 = fma(xa[u], ya[u], g0);
                if (u + 1 < RL) g1 = fma(xa[u + 1], ya[u + 1], g1);
                if (u + 2 < RL) g2 = fma(xa[u + 2], ya[u + 2], g2);
                if (u + 3 < RL) g3 = fma(xa[u + 3], ya[u + 3], g3);
            }
            double g = (g0 + g1) + (g2 + g3);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) g += __shfl_xor_sync(hmask, g, o);
            double c, sn;
            if (osj_cs(g, al, be, skip2, rmax2, c, sn)) {
#pragma unroll
                for (int u = 0; u < RL; u++) {
                    const double x = xa[u], y = ya[u];
                    xa[u] = c * x - sn * y;
                    osj_sts(ay + 128u * u, sn * x + c * y);
                }
                const double cc = c * c, ss = sn * sn, cs2 = 2.0 * c * sn * g;
                if (hl == 0) osj_sts(sN + 8u * iy, fma(ss, al, fma(cc, be, cs2)));
                al = fma(cc, al, fma(ss, be, -cs2));
            }
        }
        __syncthreads();
    }
    if (act) {
#pragma unroll
        for (int u = 0; u < RL; u++) osj_sts(ax + 128u * u, xa[u]);
        if (hl == 0) osj_sts(sN + 8u * hwid, al);
    }
    __syncthreads();
}

template <int RL>
__global__ void __cluster_dims__(OSJ_CLUSTER, 1, 1) __launch_bounds__(OSJ_THREADS, 1)
osj_kernel(double *Ac, int b, int BP, int H, int max_sweeps, double tol, double predict,
           double *cmax, double *w_out, int *info, long long *trace) {
    // optional phase trace (TADPOLE_OSJ_TRACE): cycles of thread 0 of every CTA summed per phase over the whole solve
    long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tr_t = 0;
#define OSJ_TR0() do { if (trace && threadIdx.x == 0) tr_t = clock64(); } while (0)
#define OSJ_TR(ph) do { if (trace && threadIdx.x == 0) { const long long t_ = clock64(); tr_acc[ph] += t_ - tr_t; tr_t = t_; } } while (0)
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double s_osj[];
    double *sA = s_osj;                       // 2H columns x BP
    __shared__ double s_norm[2 * OSJ_MAXH];
    const double tol2 = tol * tol;
    __shared__ double s_red[OSJ_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int hwid = tid >> 4, hl = tid & 15;                       // half-warp index (one column pair each) and lane in it
    const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
    const int crank = cluster.block_rank();
    const unsigned uA = (unsigned)__cvta_generic_to_shared(sA), uN = (unsigned)__cvta_generic_to_shared(s_norm);
    const int Hm = (H + 1) & ~1;
    const double skip2 = 1e-34;
    int sweeps = 0;
    bool converged = false;
    while (!converged && sweeps < max_sweeps) {
        double rmax = 0.0;
        for (int bs = 0; bs < OSJ_NHP - 1; bs++) {
            int hI, hJ;
            osj_pair(crank, bs, OSJ_NHP, hI, hJ);
            OSJ_TR0();
            // ---- load the two half-panels (columns contiguous in global: coalesced) -----------
            const int colsz = H * BP;
            for (int i0 = 0; i0 < colsz; i0 += OSJ_THREADS * 8) {       // all loads of a batch in flight together
                double ti[8], tj[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int idx = i0 + u * OSJ_THREADS + tid;
                    if (idx < colsz) { ti[u] = __ldcg(&Ac[(size_t)hI * colsz + idx]); tj[u] = __ldcg(&Ac[(size_t)hJ * colsz + idx]); }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int idx = i0 + u * OSJ_THREADS + tid;
                    if (idx < colsz) { sA[idx] = ti[u]; sA[colsz + idx] = tj[u]; }
                }
            }
            __syncthreads();
            OSJ_TR(0);
            // squared column norms, recomputed every block step (no drift)
            for (int c = wid; c < 2 * H; c += OSJ_THREADS / 32) {
                double a = 0.0;
                for (int r = lane; r < BP; r += 32) { const double v = sA[c * BP + r]; a = fma(v, v, a); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) s_norm[c] = a;
            }
            __syncthreads();
            OSJ_TR(1);
            // ---- pairs inside each half-panel, once per sweep ---------------------------------
            if (bs == 0) {
                for (int st = 0; st < Hm - 1; st++) {
                    for (int pw = hwid; pw < Hm; pw += OSJ_THREADS / 16) {     // Hm pair slots over the half-warps
                        const int half = pw / (Hm / 2), slot = pw % (Hm / 2);
                        int p, q;
                        osj_pair(slot, st, Hm, p, q);
                        if (q < H) osj_rotate<RL>(uA, uN, BP, half * H + p, half * H + q, hl, hmask, skip2, rmax);
                    }
                    __syncthreads();
                }
            }
            OSJ_TR(2);
            // ---- all pairs across the two half-panels: H sub-steps of H disjoint pairs ----------
            if (H <= OSJ_THREADS / 16) {
                osj_cross_xreg<RL>(uA, uN, BP, H, hwid, hl, hmask, skip2, rmax);
            } else {
                for (int st = 0; st < H; st++) {
                    for (int pw = hwid; pw < H; pw += OSJ_THREADS / 16) {
                        int jj = pw + st; if (jj >= H) jj -= H;
                        osj_rotate<RL>(uA, uN, BP, pw, H + jj, hl, hmask, skip2, rmax);
                    }
                    __syncthreads();
                }
            }
            OSJ_TR(3);
            // ---- write back ---------------------------------------------------------------------------
            for (int idx = tid; idx < colsz; idx += OSJ_THREADS) {
                Ac[(size_t)hI * colsz + idx] = sA[idx];
                Ac[(size_t)hJ * colsz + idx] = sA[colsz + idx];
            }
            if (bs == OSJ_NHP - 2 && w_out) {           // eigenvalue estimates of this sweep
                for (int c = tid; c < 2 * H; c += OSJ_THREADS) {
                    const int gcol = (c < H ? hI : hJ) * H + (c < H ? c : c - H);
                    w_out[gcol] = s_norm[c];
                }
            }
            OSJ_TR(4);
            cluster.sync();
            OSJ_TR(5);
        }
        sweeps++;
        OSJ_TR0();
        // largest scaled cosine seen by this CTA -> cluster-wide maximum through global memory
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, 16));   // only lane values of the warp matter
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        if (lane == 0) s_red[wid] = rmax;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < OSJ_THREADS / 32; w++) v = fmax(v, s_red[w]);
            cmax[(sweeps & 1) * OSJ_CLUSTER + crank] = v;
        }
        cluster.sync();
        double v = 0.0;
        for (int r = 0; r < OSJ_CLUSTER; r++) v = fmax(v, __ldcg(&cmax[(sweeps & 1) * OSJ_CLUSTER + r]));
        // squared scaled cosines, as they were BEFORE this sweep's rotations.  Jacobi converges quadratically once the
        // cosines are below the relative gaps of the eigenvalues: with predict > 0 (callers that check the result
        // themselves: the Rayleigh-Ritz steps of the subspace iteration, whose residual test follows) the sweep that would
        // only confirm convergence is not run when predict * v^2 -- what this sweep leaves behind for relative gaps down to
        // predict^(-1/2) -- is already below the tolerance.
        converged = v <= tol2 || (predict > 0.0 && predict * v * v <= tol2);
        OSJ_TR(6);
    }
    if (trace && threadIdx.x == 0) {
        for (int p = 0; p < 7; p++) trace[crank * 8 + p] = tr_acc[p];
        trace[crank * 8 + 7] = sweeps;
    }
    if (crank == 0 && tid == 0) { if (!converged) info[1] = 1; info[2] += sweeps; }     // sticky status words
}

// =====================================================================================================================
// The same solver in FP32, for solves whose tolerance is loose (the start Rayleigh-Ritz step of the subspace iteration, of
// which only the Ritz VALUES are used: 1e-2).  A sub-step is one dependent chain (profiles/r02_osj_trace.md); in FP32 its
// links cost ~4 cycles instead of ~20 and the y columns are half the shared-memory bytes.  The factor stays FP64 in global
// memory: a half-panel is converted on load (scaled by a power of two so that the largest diagonal entry of L is in [1, 2):
// squared norms stay far from the FP32 range limits) and converted back on write-back.
// =====================================================================================================================
__device__ __forceinline__ float osj_ldsf(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void osj_stsf(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

__device__ __forceinline__ bool osj_csf(float g, float al, float be, float skip2, float &rmax2, float &c, float &sn) {
    const float ab = al * be, gg = g * g;
    if (!(ab > 0.0f)) return false;
    rmax2 = fmaxf(rmax2, __fdividef(gg, ab));
    if (!(gg > skip2 * ab)) return false;
    const float dd = be - al;
    const float hh = fmaf(dd, dd, 4.0f * gg);
    const float hf = hh * rsqrtf(hh);
    float t = __fdividef(2.0f * g, fabsf(dd) + hf);
    if (dd < 0.0f) t = -t;
    const float tt = fmaf(t, t, 1.0f);
    c = rsqrtf(tt);
    c = c * fmaf(-0.5f * tt * c, c, 1.5f);          // c^2 + s^2 = 1 to FP32 rounding
    sn = t * c;
    return true;
}

template <int RL>
__device__ __forceinline__ void osj_rotatef(unsigned sA, unsigned sN, int BP, int ix, int iy, int hl, unsigned hmask,
                                            float skip2, float &rmax2) {
    const unsigned ax = sA + 4u * (ix * BP + hl), ay = sA + 4u * (iy * BP + hl);
    float xa[RL], ya[RL];
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
    for (int u = 0; u < RL; u++) { xa[u] = osj_ldsf(ax + 64u * u); ya[u] = osj_ldsf(ay + 64u * u); }
    const float al = osj_ldsf(sN + 4u * ix), be = osj_ldsf(sN + 4u * iy);
#pragma unroll
    for (int u = 0; u < RL; u += 4) {
        g0 = fmaf(xa[u], ya[u], g0);
        if (u + 1 < RL) g1 = fmaf(xa[u + 1], ya[u + 1], g1);
        if (u + 2 < RL) g2 = fmaf(xa[u + 2], ya[u + 2], g2);
        if (u + 3 < RL) g3 = fmaf(xa[u + 3], ya[u + 3], g3);
    }
    float g = (g0 + g1) + (g2 + g3);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) g += __shfl_xor_sync(hmask, g, o);
    float c, sn;
    if (!osj_csf(g, al, be, skip2, rmax2, c, sn)) return;
#pragma unroll
    for (int u = 0; u < RL; u++) {
        osj_stsf(ax + 64u * u, c * xa[u] - sn * ya[u]);
        osj_stsf(ay + 64u * u, sn * xa[u] + c * ya[u]);
    }
    if (hl == 0) {
        const float cc = c * c, ss = sn * sn, cs2 = 2.0f * c * sn * g;
        osj_stsf(sN + 4u * ix, fmaf(cc, al, fmaf(ss, be, -cs2)));
        osj_stsf(sN + 4u * iy, fmaf(ss, al, fmaf(cc, be, cs2)));
    }
}

template <int RL>
__global__ void __cluster_dims__(OSJ_CLUSTER, 1, 1) __launch_bounds__(OSJ_THREADS, 1)
osj32_kernel(double *Ac, int b, int BP, int H, int max_sweeps, double tol, const double *scale_p,
             double *cmax, double *w_out, int *info) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float s_osjf[];
    float *sA = s_osjf;                        // 2H columns x BP
    __shared__ float s_norm[2 * OSJ_MAXH];
    __shared__ float s_red[OSJ_THREADS / 32];
    const float tol2 = (float)(tol * tol);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int hwid = tid >> 4, hl = tid & 15;
    const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
    const int crank = cluster.block_rank();
    const unsigned uA = (unsigned)__cvta_generic_to_shared(sA), uN = (unsigned)__cvta_generic_to_shared(s_norm);
    const int Hm = (H + 1) & ~1;
    const float skip2 = 1e-13f;
    const double scl = __ldg(scale_p), iscl = 1.0 / scl;          // powers of two
    int sweeps = 0;
    bool converged = false;
    while (!converged && sweeps < max_sweeps) {
        float rmax = 0.f;
        for (int bs = 0; bs < OSJ_NHP - 1; bs++) {
            int hI, hJ;
            osj_pair(crank, bs, OSJ_NHP, hI, hJ);
            const int colsz = H * BP;
            for (int i0 = 0; i0 < colsz; i0 += OSJ_THREADS * 8) {
                double ti[8], tj[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int idx = i0 + u * OSJ_THREADS + tid;
                    if (idx < colsz) { ti[u] = __ldcg(&Ac[(size_t)hI * colsz + idx]); tj[u] = __ldcg(&Ac[(size_t)hJ * colsz + idx]); }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int idx = i0 + u * OSJ_THREADS + tid;
                    if (idx < colsz) { sA[idx] = (float)(ti[u] * scl); sA[colsz + idx] = (float)(tj[u] * scl); }
                }
            }
            __syncthreads();
            for (int c = wid; c < 2 * H; c += OSJ_THREADS / 32) {
                float a = 0.f;
                for (int r = lane; r < BP; r += 32) { const float v = sA[c * BP + r]; a = fmaf(v, v, a); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) s_norm[c] = a;
            }
            __syncthreads();
            if (bs == 0) {
                for (int st = 0; st < Hm - 1; st++) {
                    for (int pw = hwid; pw < Hm; pw += OSJ_THREADS / 16) {
                        const int half = pw / (Hm / 2), slot = pw % (Hm / 2);
                        int p, q;
                        osj_pair(slot, st, Hm, p, q);
                        if (q < H) osj_rotatef<RL>(uA, uN, BP, half * H + p, half * H + q, hl, hmask, skip2, rmax);
                    }
                    __syncthreads();
                }
            }
            if (H <= OSJ_THREADS / 16) {       // x column of every pair in registers over the H sub-steps
                const bool act = hwid < H;
                const unsigned ax = uA + 4u * (hwid * BP + hl);
                float xa[RL], al = 0.f;
                if (act) {
#pragma unroll
                    for (int u = 0; u < RL; u++) xa[u] = osj_ldsf(ax + 64u * u);
                    al = osj_ldsf(uN + 4u * hwid);
                }
                for (int st = 0; st < H; st++) {
                    if (act) {
                        int jj = hwid + st; if (jj >= H) jj -= H;
                        const int iy = H + jj;
                        const unsigned ay = uA + 4u * (iy * BP + hl);
                        float ya[RL];
                        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
                        for (int u = 0; u < RL; u++) ya[u] = osj_ldsf(ay + 64u * u);
                        const float be = osj_ldsf(uN + 4u * iy);
#pragma unroll
                        for (int u = 0; u < RL; u += 4) {
                            g0 = fmaf(xa[u], ya[u], g0);
                            if (u + 1 < RL) g1 = fmaf(xa[u + 1], ya[u + 1], g1);
                            if (u + 2 < RL) g2 = fmaf(xa[u + 2], ya[u + 2], g2);
                            if (u + 3 < RL) g3 = fmaf(xa[u + 3], ya[u + 3], g3);
                        }
                        float g = (g0 + g1) + (g2 + g3);
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) g += __shfl_xor_sync(hmask, g, o);
                        float c, sn;
                        if (osj_csf(g, al, be, skip2, rmax, c, sn)) {
#pragma unroll
                            for (int u = 0; u < RL; u++) {
                                const float x = xa[u], y = ya[u];
                                xa[u] = c * x - sn * y;
                                osj_stsf(ay + 64u * u, sn * x + c * y);
                            }
                            const float cc = c * c, ss = sn * sn, cs2 = 2.0f * c * sn * g;
                            if (hl == 0) osj_stsf(uN + 4u * iy, fmaf(ss, al, fmaf(cc, be, cs2)));
                            al = fmaf(cc, al, fmaf(ss, be, -cs2));
                        }
                    }
                    __syncthreads();
                }
                if (act) {
#pragma unroll
                    for (int u = 0; u < RL; u++) osj_stsf(ax + 64u * u, xa[u]);
                    if (hl == 0) osj_stsf(uN + 4u * hwid, al);
                }
                __syncthreads();
            } else {
                for (int st = 0; st < H; st++) {
                    for (int pw = hwid; pw < H; pw += OSJ_THREADS / 16) {
                        int jj = pw + st; if (jj >= H) jj -= H;
                        osj_rotatef<RL>(uA, uN, BP, pw, H + jj, hl, hmask, skip2, rmax);
                    }
                    __syncthreads();
                }
            }
            for (int idx = tid; idx < colsz; idx += OSJ_THREADS) {
                Ac[(size_t)hI * colsz + idx] = (double)sA[idx] * iscl;
                Ac[(size_t)hJ * colsz + idx] = (double)sA[colsz + idx] * iscl;
            }
            if (bs == OSJ_NHP - 2 && w_out) {
                for (int c = tid; c < 2 * H; c += OSJ_THREADS) {
                    const int gcol = (c < H ? hI : hJ) * H + (c < H ? c : c - H);
                    w_out[gcol] = (double)s_norm[c] * iscl * iscl;
                }
            }
            cluster.sync();
        }
        sweeps++;
        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 16));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        if (lane == 0) s_red[wid] = rmax;
        __syncthreads();
        if (tid == 0) {
            float v = 0.f;
            for (int w = 0; w < OSJ_THREADS / 32; w++) v = fmaxf(v, s_red[w]);
            cmax[(sweeps & 1) * OSJ_CLUSTER + crank] = (double)v;
        }
        cluster.sync();
        float v = 0.f;
        for (int r = 0; r < OSJ_CLUSTER; r++) v = fmaxf(v, (float)__ldcg(&cmax[(sweeps & 1) * OSJ_CLUSTER + r]));
        converged = v <= tol2;
    }
    if (crank == 0 && tid == 0) { if (!converged) info[1] = 1; info[2] += sweeps; }
}

// scale = 2^-e with max_i L_ii in [2^e, 2^(e+1)) (1 when the factor is zero)
__global__ void osj_scale_kernel(const double *__restrict__ L, int b, int ld, double *__restrict__ scale) {
    __shared__ double s[32];
    double mx = 0.0;
    for (int i = threadIdx.x; i < b; i += blockDim.x) mx = fmax(mx, fabs(L[(size_t)i * ld + i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned w = 1; w < (blockDim.x >> 5); w++) mx = fmax(mx, s[w]);
        int e = 0;
        if (mx > 0.0 && isfinite(mx)) frexp(mx, &e);            // mx = f 2^e, f in [0.5, 1)
        *scale = ldexp(1.0, 1 - e);
    }
}

// eigenvalues w (unsorted, one per column) -> descending order; eigenvector = normalised column of the
// rotated factor: Vs[r][rank(c)] = Ac[c * BP + r] / sqrt(w_c)
__global__ void osj_sort_kernel(const double *__restrict__ w_in, const double *__restrict__ Vc, int b, int BP,
                                double *__restrict__ w, double *__restrict__ Vs, int lds, int ncols_out) {
    extern __shared__ int s_rank[];
    for (int i = threadIdx.x; i < b; i += blockDim.x) {
        const double wi = w_in[i];
        int rank = 0;
        for (int j = 0; j < b; j++) {
            const double wj = w_in[j];
            rank += (wj > wi) || (wj == wi && j < i);
        }
        s_rank[i] = rank;
        if (blockIdx.x == 0) w[rank] = wi;
    }
    __syncthreads();
    for (int c = blockIdx.x; c < b; c += gridDim.x) {
        const int dst = s_rank[c];
        if (dst >= ncols_out) continue;
        const double wc = w_in[c];
        const double sc = wc > 0.0 ? 1.0 / sqrt(wc) : 0.0;
        for (int r = threadIdx.x; r < b; r += blockDim.x) Vs[(size_t)r * lds + dst] = Vc[(size_t)c * BP + r] * sc;
    }
}

int tp_chol_factor(tp_ctx *ctx, double *G, int b, int ld);

// Eigen-decomposition of the symmetric positive semi-definite b x b matrix T (row-major, ld), b <= 384.
// T is destroyed.  w[0..b) descending, Vs (b x lds) eigenvectors in its first ncols_out columns.
int tp_osj(tp_ctx *ctx, double *T, int b, int ld, double *w, double *Vs, int lds, int ncols_out, int *sweeps_out,
           double tol, double predict) {
    TP_ARG(b >= 2 && b <= OSJ_MAX_B, "tp_osj: b out of range");
    cudaStream_t st = ctx->stream;
    const int H = (b + OSJ_NHP - 1) / OSJ_NHP, NC = H * OSJ_NHP, BP = round_up(b, 32);
    const size_t panel = (size_t)NC * BP * sizeof(double);
    TP_TRY(ctx->Jt.reserve(panel + (size_t)(NC + 64) * sizeof(double) + 128));
    double *Ac = ctx->Jt.as<double>();
    double *wtmp = Ac + (size_t)NC * BP, *cmax = wtmp + NC;
    TP_TRY(ctx->status.reserve(64));
    int *info = ctx->status.as<int>();
    if (sweeps_out) TP_TRY(tp_flags_reset(ctx));
    TP_TRY(tp_chol_factor(ctx, T, b, ld));
    tp_prof_begin(ctx, PC_JACOBI);
    osj_init_kernel<<<(unsigned)(((size_t)NC * BP + 255) / 256), 256, 0, st>>>(T, b, ld, BP, NC, Ac);
    // loose tolerances (>= 1e-4: the start Rayleigh-Ritz step) are solved in FP32
    const bool f32 = tol >= 1e-4 && !getenv("TADPOLE_OSJ_NOF32");
    if (f32) {
        double *scale = cmax + 2 * OSJ_CLUSTER;
        osj_scale_kernel<<<1, 256, 0, st>>>(T, b, ld, scale);
        const size_t smem32 = (size_t)2 * H * BP * sizeof(float);
#define OSJ32_LAUNCH(R)                                                                                         \
    case R:                                                                                                     \
        TP_CUDA(tp_optin_smem(osj32_kernel<2 * R>, ctx)); \
        osj32_kernel<2 * R><<<OSJ_CLUSTER, OSJ_THREADS, smem32, st>>>(Ac, b, BP, H, 30, tol, scale, cmax, wtmp, info); \
        break;
        switch (BP / 32) {
            OSJ32_LAUNCH(1) OSJ32_LAUNCH(2) OSJ32_LAUNCH(3) OSJ32_LAUNCH(4) OSJ32_LAUNCH(5) OSJ32_LAUNCH(6)
            OSJ32_LAUNCH(7) OSJ32_LAUNCH(8) OSJ32_LAUNCH(9) OSJ32_LAUNCH(10) OSJ32_LAUNCH(11) OSJ32_LAUNCH(12)
            default: tp_set_error("tp_osj: unsupported panel height"); return TP_ERR_ARG;
        }
#undef OSJ32_LAUNCH
        ctx->launches += 1;
    }
    const size_t smem = (size_t)2 * H * BP * sizeof(double);
    const double otol = tol > 2e-15 ? tol : 2e-15;
    long long *trace = nullptr;
    DevBuf trbuf;
    if (getenv("TADPOLE_OSJ_TRACE")) { TP_TRY(trbuf.reserve(OSJ_CLUSTER * 8 * sizeof(long long))); trace = trbuf.as<long long>(); }
#define OSJ_LAUNCH(R)                                                                                         \
    case R:                                                                                                   \
        TP_CUDA(tp_optin_smem(osj_kernel<2 * R>, ctx)); \
        osj_kernel<2 * R><<<OSJ_CLUSTER, OSJ_THREADS, smem, st>>>(Ac, b, BP, H, 30, otol, predict, cmax, wtmp, info, trace); \
        break;
    if (!f32) switch (BP / 32) {
        OSJ_LAUNCH(1) OSJ_LAUNCH(2) OSJ_LAUNCH(3) OSJ_LAUNCH(4) OSJ_LAUNCH(5) OSJ_LAUNCH(6)
        OSJ_LAUNCH(7) OSJ_LAUNCH(8) OSJ_LAUNCH(9) OSJ_LAUNCH(10) OSJ_LAUNCH(11) OSJ_LAUNCH(12)
        default: tp_set_error("tp_osj: unsupported panel height"); return TP_ERR_ARG;
    }
#undef OSJ_LAUNCH
    const int grid = b < 2 * ctx->sm_count ? b : 2 * ctx->sm_count;
    osj_sort_kernel<<<grid, 256, (size_t)b * sizeof(int), st>>>(wtmp, Ac, b, BP, w, Vs, lds, ncols_out);
    tp_prof_end(ctx);
    ctx->launches += 3;
    TP_CUDA(cudaGetLastError());
    if (trace) {
        long long h[OSJ_CLUSTER * 8];
        TP_CUDA(cudaStreamSynchronize(st));
        TP_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
        static const char *ph[7] = {"load", "norms", "intra", "cross", "writeback", "cluster_sync", "converge"};
        for (int c = 0; c < OSJ_CLUSTER; c += 7) {
            fprintf(stderr, "[osj trace] b=%d tol=%.1e cta %d sweeps %lld:", b, otol, c, h[c * 8 + 7]);
            for (int p = 0; p < 7; p++) fprintf(stderr, " %s %lld", ph[p], h[c * 8 + p]);
            fprintf(stderr, "\n");
        }
        trbuf.release();
    }
    if (sweeps_out) {        // synchronous use: status read back now; otherwise the caller polls tp_flags_read
        int f[4];
        TP_TRY(tp_flags_read(ctx, f));
        *sweeps_out = f[2];
        if (getenv("TADPOLE_DEBUG")) fprintf(stderr, "[tadpole] osj b=%d tol=%.1e sweeps=%d converged=%d\n", b, otol, f[2], !f[1]);
        if (f[1]) {
            tp_set_error("tp_osj: no convergence after %d sweeps (b = %d)", f[2], b);
            return TP_ERR_NOCONV;
        }
    }
    return TP_OK;
}
