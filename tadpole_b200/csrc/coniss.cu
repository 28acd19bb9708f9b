// coniss.cu -- stages 4 and 5: the find_params sweep (reference R/TADpole.R:104-123).
//
// Per candidate number of PCs i the reference builds dist(pcs[,1:i]) (O(n^2 i)) and runs
// rioja::chclust(method="coniss") on it.  CONISS on Euclidean data is Ward clustering restricted
// to ADJACENT clusters; every cluster is a contiguous interval of bins, so a cluster sum is a
// difference of two rows of a column-wise prefix sum P of the score matrix, and the same P
// (n+1 x k, L2 resident) serves every candidate: candidate i reads the first i columns.
//
//   dSS(A,B) = nA nB / (nA + nB) * sum_c (sumA_c / nA - sumB_c / nB)^2
//
// One warp owns one candidate.  The per-boundary dSS array lives in shared memory under a 32-ary
// min tree (block minima of 32 entries, two levels), so one merge step costs: descend the tree
// (ballots), two dot products over i columns (L2 loads + shuffle reduce), and re-reduce at most
// three leaf blocks.  Ties resolve to the lowest boundary index at every level, which is the
// reference's strict '<' scan.  The loop is serial in n by nature.
#include "common.cuh"

#define INF_D (__longlong_as_double(0x7ff0000000000000LL))

// boundary links live in shared memory when they fit, otherwise in global memory read through L2
template <bool SMEM, typename T> __device__ __forceinline__ int ld_link(const T *p) { return SMEM ? (int)*p : (int)__ldcg(p); }
template <bool SMEM, typename T> __device__ __forceinline__ void st_link(T *p, int v) { if (SMEM) *p = (T)v; else __stcg(p, (T)v); }

// ------------------------------------------------------------------------------------------
// prefix sums: P[0] = 0, P[r+1] = P[r] + scores[r]  (per column);  Qp[r+1] = Qp[r] + |scores[r]|^2
// ------------------------------------------------------------------------------------------
__global__ void rownorm2_kernel(const double *__restrict__ s, int n, int k, int ldk, double *__restrict__ out) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    double a = 0.0;
    for (int c = lane; c < k; c += 32) { double v = s[(size_t)w * ldk + c]; a += v * v; }
    a = warp_sum(a);
    if (lane == 0) out[w] = a;
}

__global__ void prefix_kernel(const double *__restrict__ s, const double *__restrict__ rn2, int n, int k, int ldk,
                              double *__restrict__ P, double *__restrict__ Qp) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < k) {
        double acc = 0.0;
        P[c] = 0.0;
#pragma unroll 8
        for (int r = 0; r < n; r++) {
            acc += s[(size_t)r * ldk + c];
            P[(size_t)(r + 1) * ldk + c] = acc;
        }
    } else if (c < ldk) {
        for (int r = 0; r <= n; r++) P[(size_t)r * ldk + c] = 0.0;   // padding columns
    } else if (c == ldk) {
        double acc = 0.0;
        Qp[0] = 0.0;
#pragma unroll 8
        for (int r = 0; r < n; r++) { acc += rn2[r]; Qp[r + 1] = acc; }
    }
}

// d0[c][j] = 1/2 sum_{c' <= c} (x_j[c'] - x_{j+1}[c'])^2 : the initial adjacent-pair dSS of every
// candidate is a running sum over columns, so all k candidates are initialised in O(n k).
__global__ void d0_kernel(const double *__restrict__ s, int n, int k, int ldk, double *__restrict__ d0, int ldd) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n - 1) return;
    const double *a = s + (size_t)j * ldk, *b = a + ldk;
    double acc = 0.0;
    for (int c = 0; c < k; c++) {
        double t = a[c] - b[c];
        acc += 0.5 * t * t;
        d0[(size_t)c * ldd + j] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// the merge loop
// ------------------------------------------------------------------------------------------
template <typename LinkT, bool LINKS_SMEM>
__global__ void __launch_bounds__(32)
coniss_sweep_kernel(const double *__restrict__ P, int ldk, int n,
                    const double *__restrict__ d0, int ldd,
                    const int *__restrict__ cand_list,
                    double *__restrict__ seqdist, int4 *__restrict__ merges,
                    LinkT *__restrict__ glinks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int cand = cand_list[blockIdx.x];
    const int ncol = cand + 1;
    const int n1 = n - 1;
    const int n1p = (n1 + 31) & ~31;
    const int B1 = n1p >> 5;
    const int B1p = (B1 + 31) & ~31;
    const int B2 = B1p >> 5;
    const int B2p = (B2 + 31) & ~31;

    double *d = (double *)smem_raw;
    double *m1 = d + n1p;
    double *m2 = m1 + B1p;
    LinkT *prv, *nxt;
    if (LINKS_SMEM) {
        prv = (LinkT *)(m2 + B2p);
        nxt = prv + n1;
    } else {
        prv = glinks + (size_t)blockIdx.x * 2 * n1;
        nxt = prv + n1;
    }
    const double *d0row = d0 + (size_t)cand * ldd;
    double *seq = seqdist + (size_t)cand * ldd;
    int4 *mrg = merges + (size_t)cand * ldd;

    for (int j = lane; j < n1p; j += 32) d[j] = (j < n1) ? d0row[j] : INF_D;
    for (int j = lane; j < n1; j += 32) { st_link<LINKS_SMEM>(prv + j, j); st_link<LINKS_SMEM>(nxt + j, j + 1); }   // prv holds index+1, 0 = none
    for (int b = lane; b < B1p; b += 32) m1[b] = INF_D;
    for (int b = lane; b < B2p; b += 32) m2[b] = INF_D;
    __syncwarp();
    for (int b = 0; b < B1; b++) {
        double v = warp_min_nonneg(d[(b << 5) + lane]);
        if (lane == 0) m1[b] = v;
    }
    __syncwarp();
    for (int b = 0; b < B2; b++) {
        double v = warp_min_nonneg(m1[(b << 5) + lane]);
        if (lane == 0) m2[b] = v;
    }
    __syncwarp();

    double total = 0.0;
    for (int t = 0; t < n1; t++) {
        // ---- find the lowest-index minimum ------------------------------------------------
        double v = INF_D;
        for (int q = lane; q < B2p; q += 32) v = fmin(v, m2[q]);
        const double mn = warp_min_nonneg(v);
        int b2 = 0;
        for (int q0 = 0; q0 < B2p; q0 += 32) {
            unsigned bal = __ballot_sync(0xffffffffu, m2[q0 + lane] == mn);
            if (bal) { b2 = q0 + __ffs(bal) - 1; break; }
        }
        unsigned bal1 = __ballot_sync(0xffffffffu, m1[(b2 << 5) + lane] == mn);
        const int b1 = (b2 << 5) + __ffs(bal1) - 1;
        unsigned bal0 = __ballot_sync(0xffffffffu, d[(b1 << 5) + lane] == mn);
        const int j = (b1 << 5) + __ffs(bal0) - 1;

        // ---- neighbours ---------------------------------------------------------------------
        const int pj = ld_link<LINKS_SMEM>(prv + j) - 1;         // previous live boundary or -1
        const int nj = ld_link<LINKS_SMEM>(nxt + j);            // next live boundary or n1
        const bool hasL = pj >= 0, hasR = nj < n1;
        const int ppj = hasL ? ld_link<LINKS_SMEM>(prv + pj) - 1 : -1;
        const int nnj = hasR ? ld_link<LINKS_SMEM>(nxt + nj) : n1;
        total += mn;
        if (lane == 0) {
            seq[j] = total;
            mrg[t] = make_int4(j, pj, nj, 0);
        }
        // P rows: LL = [a, b), C = [b, c), RR = [c, e)
        const int a = ppj + 1, b = pj + 1, c = nj + 1, e = nnj + 1;
        const double nC = (double)(c - b), nLL = (double)(b - a), nRR = (double)(e - c);
        const double iC = 1.0 / nC, iLL = 1.0 / nLL, iRR = 1.0 / nRR;
        const double *Pa = P + (size_t)a * ldk, *Pb = P + (size_t)b * ldk;
        const double *Pc = P + (size_t)c * ldk, *Pe = P + (size_t)e * ldk;
        double accL = 0.0, accR = 0.0;
        for (int col = lane; col < ncol; col += 32) {
            const double pb = __ldg(Pb + col), pc = __ldg(Pc + col);
            const double mC = (pc - pb) * iC;
            if (hasL) { double tL = (pb - __ldg(Pa + col)) * iLL - mC; accL += tL * tL; }
            if (hasR) { double tR = mC - (__ldg(Pe + col) - pc) * iRR; accR += tR * tR; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            accL += __shfl_xor_sync(0xffffffffu, accL, o);
            accR += __shfl_xor_sync(0xffffffffu, accR, o);
        }
        accL *= nLL * nC / (nLL + nC);
        accR *= nC * nRR / (nC + nRR);

        // ---- write back: boundary j dies, its neighbours get new increases ---------------
        if (lane == 0) {
            d[j] = INF_D;
            if (hasL) { d[pj] = accL; st_link<LINKS_SMEM>(nxt + pj, nj); }
            if (hasR) { d[nj] = accR; st_link<LINKS_SMEM>(prv + nj, pj + 1); }
        }
        __syncwarp();
        const int k0 = j >> 5;
        const int k1 = hasL ? (pj >> 5) : k0;
        const int k2 = hasR ? (nj >> 5) : k0;
        {
            double x = warp_min_nonneg(d[(k0 << 5) + lane]);
            if (lane == 0) m1[k0] = x;
        }
        if (k1 != k0) {
            double x = warp_min_nonneg(d[(k1 << 5) + lane]);
            if (lane == 0) m1[k1] = x;
        }
        if (k2 != k0) {
            double x = warp_min_nonneg(d[(k2 << 5) + lane]);
            if (lane == 0) m1[k2] = x;
        }
        __syncwarp();
        const int g0 = k0 >> 5, g1 = k1 >> 5, g2 = k2 >> 5;
        {
            double x = warp_min_nonneg(m1[(g0 << 5) + lane]);
            if (lane == 0) m2[g0] = x;
        }
        if (g1 != g0) {
            double x = warp_min_nonneg(m1[(g1 << 5) + lane]);
            if (lane == 0) m2[g1] = x;
        }
        if (g2 != g0 && g2 != g1) {
            double x = warp_min_nonneg(m1[(g2 << 5) + lane]);
            if (lane == 0) m2[g2] = x;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// stage 5: rioja::bstick + first-TRUE-run rule + fpc::calinhara per level (R/TADpole.R:111-120)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
ch_kernel(const double *__restrict__ P, const double *__restrict__ Qp, int ldk, int n, int k,
          const double *__restrict__ seqdist, const int4 *__restrict__ merges, int ldd,
          const int *__restrict__ cand_list, int min_clusters,
          double *__restrict__ bs_scratch, int *__restrict__ ncl_out,
          double *__restrict__ chs, int ld_chs) {
    const int lane = threadIdx.x;
    const int cand = cand_list[blockIdx.x];
    const int n1 = n - 1;                 // nobj = number of merges = length(height)
    const double *seq = seqdist + (size_t)cand * ldd;
    const int4 *mrg = merges + (size_t)cand * ldd;
    double *bs = bs_scratch + (size_t)blockIdx.x * ldd;
    double *out = chs + (size_t)cand * ld_chs;

    for (int l = lane; l < ld_chs; l += 32) out[l] = __longlong_as_double(0x7ff8000000000000LL);

    // height[t] = cumulative dispersion after merge t; tot = height[n1-1]
    const double tot = seq[mrg[n1 - 1].x];
    // vegan::bstick.default(nobj, tot) = rev(cumsum(tot / nobj:1) / nobj): the cumulative sum runs
    // from m = nobj down to 1 in this order, one thread, to round exactly as R does.
    if (lane == 0) {
        double c = 0.0;
        const double dn = (double)n1;
        for (int m = n1; m >= 1; m--) {
            c += tot / (double)m;
            bs[m - 1] = c / dn;
        }
    }
    __syncwarp();
    // dispersion_j = |disp[j+1] - disp[j]|, disp = rev(height), j = 1..n1-1;  flag_j = dispersion_j > bs_j
    int first = -1, runlen = 0;
    bool done = false;
    for (int j0 = 1; j0 <= n1 - 1 && !done; j0 += 32) {
        const int j = j0 + lane;
        bool f = false;
        if (j <= n1 - 1) {
            const double hi = seq[mrg[n1 - j].x];        // disp[j]   = height[n1 - j]
            const double lo = seq[mrg[n1 - j - 1].x];    // disp[j+1] = height[n1 - j - 1]
            f = fabs(lo - hi) > bs[j - 1];
        }
        unsigned bal = __ballot_sync(0xffffffffu, f);
        unsigned valid = (j0 + 31 <= n1 - 1) ? 0xffffffffu : ((1u << (n1 - j0)) - 1u);
        if (first < 0) {
            if (bal) {
                int s = __ffs(bal) - 1;
                first = j0 + s;
                unsigned rest = (~bal & valid) >> s;      // first FALSE at or after s
                if (rest) { runlen = __ffs(rest) - 1; done = true; }
                else if (valid != 0xffffffffu) { runlen = (n1 - 1) - first + 1; done = true; }
                else runlen = 32 - s;
            }
        } else {
            unsigned nb = ~bal & valid;
            if (nb) { runlen += __ffs(nb) - 1; done = true; }
            else if (valid != 0xffffffffu) { runlen += n1 - j0; done = true; }
            else runlen += 32;
        }
    }
    const int ncl = (first < 0) ? -1 : runlen;
    if (lane == 0) ncl_out[cand] = ncl;
    if (ncl < 1) return;

    // Calinski-Harabasz on all k columns; level n is reached by undoing the last n-1 merges.
    double ssq = 0.0;
    for (int c = lane; c < k; c += 32) { double v = P[(size_t)n * ldk + c]; ssq += v * v; }
    ssq = warp_sum(ssq);
    const double trS = Qp[n] - ssq / (double)n;
    double W = trS;
    const int mc = min(min_clusters, ncl);
    const double dN = (double)n;
    if (mc <= 1 && lane == 0 && ld_chs > 0) out[0] = (dN - 1.0) * (trS - W) / (0.0 * W);
    for (int lev = 2; lev <= ncl; lev++) {
        const int4 mg = mrg[n1 - (lev - 1)];
        const int a = mg.y + 1, b = mg.x + 1, c = mg.z + 1;   // A = [a, b), B = [b, c) in P rows
        const double nA = (double)(b - a), nB = (double)(c - b);
        const double iA = 1.0 / nA, iB = 1.0 / nB;
        const double *Pa = P + (size_t)a * ldk, *Pb = P + (size_t)b * ldk, *Pc = P + (size_t)c * ldk;
        double acc = 0.0;
        for (int col = lane; col < k; col += 32) {
            const double pb = Pb[col];
            const double tt = (pb - Pa[col]) * iA - (Pc[col] - pb) * iB;
            acc += tt * tt;
        }
        acc = warp_sum(acc);
        W -= acc * (nA * nB / (nA + nB));
        if (lev >= mc && lev <= ld_chs && lane == 0)
            out[lev - 1] = (dN - (double)lev) * (trS - W) / ((double)(lev - 1) * W);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int build_cand_list(int k, int begin, int stride, std::vector<int> &out) {
    out.clear();
    for (int c = begin; c < k; c += stride) out.push_back(c);
    // heaviest (most columns) first
    for (size_t i = 0, j = out.size(); i + 1 < j; i++, j--) std::swap(out[i], out[j - 1]);
    return (int)out.size();
}

int tp_sweep_device(tp_ctx *ctx, int min_clusters, int cand_begin, int cand_stride, int *ncand_out) {
    TP_ARG(ctx->have_scores, "tp_sweep: no PC scores in the context (run tp_pca or tp_set_scores)");
    TP_ARG(cand_stride >= 1 && cand_begin >= 0, "tp_sweep: bad candidate range");
    TP_ARG(min_clusters >= 1, "tp_sweep: min_clusters must be >= 1");
    const int n = ctx->nf, k = ctx->k, ldk = ctx->ldk;
    TP_ARG(n >= 3, "tp_sweep: need at least 3 bins");
    const int n1 = n - 1;
    const int ldd = round_up(n1, 8);
    std::vector<int> cands;
    const int ncand = build_cand_list(k, cand_begin, cand_stride, cands);
    *ncand_out = ncand;
    if (ncand == 0) return TP_OK;
    cudaStream_t st = ctx->stream;

    TP_TRY(ctx->P.reserve((size_t)(n + 1) * ldk * sizeof(double)));
    TP_TRY(ctx->Qp.reserve((size_t)(n + 2) * sizeof(double) * 2));
    TP_TRY(ctx->d0.reserve((size_t)k * ldd * sizeof(double)));
    TP_TRY(ctx->seqdist.reserve((size_t)k * ldd * sizeof(double)));
    TP_TRY(ctx->order.reserve((size_t)k * ldd * sizeof(int4)));
    TP_TRY(ctx->ncl.reserve((size_t)(2 * k + 8) * sizeof(int)));
    TP_TRY(ctx->bsbuf.reserve((size_t)ncand * ldd * sizeof(double)));
    int *d_cands = ctx->ncl.as<int>() + k;
    TP_CUDA(cudaMemcpyAsync(d_cands, cands.data(), ncand * sizeof(int), cudaMemcpyHostToDevice, st));
    TP_CUDA(cudaMemsetAsync(ctx->ncl.p, 0, k * sizeof(int), st));

    double *rn2 = ctx->Qp.as<double>() + (n + 2);
    TP_MARK(ctx, EV_SWEEP0);
    rownorm2_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(ctx->scores.as<double>(), n, k, ldk, rn2);
    prefix_kernel<<<(ldk + 1 + 63) / 64, 64, 0, st>>>(ctx->scores.as<double>(), rn2, n, k, ldk,
                                                     ctx->P.as<double>(), ctx->Qp.as<double>());
    d0_kernel<<<(n1 + 127) / 128, 128, 0, st>>>(ctx->scores.as<double>(), n, k, ldk, ctx->d0.as<double>(), ldd);
    ctx->launches += 3;

    // shared memory plan
    const int n1p = round_up(n1, 32), B1p = round_up(n1p / 32, 32), B2p = round_up(B1p / 32, 32);
    const size_t base = (size_t)(n1p + B1p + B2p) * sizeof(double);
    const bool small_links = n <= 65535;
    const size_t link_bytes = (size_t)2 * n1 * (small_links ? 2 : 4);
    const size_t limit = (size_t)ctx->max_smem_optin;
    TP_ARG(base <= limit, "tp_sweep: matrix too large for the shared-memory dSS array (n > ~28k bins); split by centromere");
    const bool links_smem = base + link_bytes <= limit;
    const size_t smem = links_smem ? base + link_bytes : base;
    void *glinks = nullptr;
    if (!links_smem) {
        TP_TRY(ctx->links.reserve((size_t)ncand * link_bytes));
        glinks = ctx->links.p;
    }
#define LAUNCH_SWEEP(LT, LS)                                                                              \
    do {                                                                                                  \
        tp_prof_begin(ctx, PC_SWEEP);                                                                     \
        TP_CUDA(cudaFuncSetAttribute(coniss_sweep_kernel<LT, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        coniss_sweep_kernel<LT, LS><<<ncand, 32, smem, st>>>(ctx->P.as<double>(), ldk, n, ctx->d0.as<double>(), ldd, \
                                                             d_cands, ctx->seqdist.as<double>(),           \
                                                             ctx->order.as<int4>(), (LT *)glinks);         \
        tp_prof_end(ctx);                                                                                 \
    } while (0)
    if (small_links) { if (links_smem) LAUNCH_SWEEP(unsigned short, true); else LAUNCH_SWEEP(unsigned short, false); }
    else             { if (links_smem) LAUNCH_SWEEP(int, true); else LAUNCH_SWEEP(int, false); }
#undef LAUNCH_SWEEP
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    TP_MARK(ctx, EV_SWEEP1);
    ctx->have_sweep = true;
    (void)min_clusters;
    return TP_OK;
}

int tp_ch_device(tp_ctx *ctx, int min_clusters, int ncand, int ld_chs) {
    const int n = ctx->nf, k = ctx->k, ldk = ctx->ldk;
    const int ldd = round_up(n - 1, 8);
    cudaStream_t st = ctx->stream;
    TP_TRY(ctx->chs.reserve((size_t)k * ld_chs * sizeof(double)));
    ctx->ld_chs = ld_chs;
    int *d_cands = ctx->ncl.as<int>() + k;
    // rows of candidates that are not run must read as NaN too
    TP_CUDA(cudaMemsetAsync(ctx->chs.p, 0xff, (size_t)k * ld_chs * sizeof(double), st));
    tp_prof_begin(ctx, PC_CH);
    ch_kernel<<<ncand, 32, 0, st>>>(ctx->P.as<double>(), ctx->Qp.as<double>(), ldk, n, k,
                                    ctx->seqdist.as<double>(), ctx->order.as<int4>(), ldd, d_cands, min_clusters,
                                    ctx->bsbuf.as<double>(), ctx->ncl.as<int>(), ctx->chs.as<double>(), ld_chs);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    TP_MARK(ctx, EV_CH1);
    return TP_OK;
}
