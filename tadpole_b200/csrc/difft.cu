// difft.cu -- stage 6: diffT (reference R/DiffT.R:41-49) on padded per-bin TAD labels.
//
// The reference, for every bin b, builds two length-L logical vectors
//     x = tad_x[b] != tad_x | tad_x[b] == 0 ,   y = likewise for tad_y
// and scores sum(xor(x, y)): O(L^2) per pair.  With SX(b) = {j : tad_x[j] == tad_x[b] != 0} (empty
// for an uncovered bin) and SY(b) likewise, xor(x, y) is TRUE exactly on the symmetric
// difference of SX and SY, so
//     score[b] = |SX| + |SY| - 2 |SX & SY| = cnt_x[lx] + cnt_y[ly] - 2 joint[lx, ly]
// which needs only three label histograms: O(L) per pair and HBM-bound (16 L bytes per pair).
// One CTA per pair; the histograms live in one shared-memory hash table (keys tagged x-only /
// y-only / joint), filled with warp-aggregated atomics because consecutive bins share labels.
// All arithmetic is integer until the final division, so results are bit-identical to R.
#include "common.cuh"

#define DT_THREADS 256
#define DT_EMPTY 0xffffffffffffffffULL

__device__ __forceinline__ unsigned dt_hash(unsigned long long key) {
    key ^= key >> 33; key *= 0xff51afd7ed558ccdULL; key ^= key >> 33;
    return (unsigned)key;
}

__device__ __forceinline__ bool dt_add(unsigned long long *keys, int *vals, unsigned mask,
                                       unsigned long long key, int cnt) {
    unsigned s = dt_hash(key) & mask;
    for (unsigned probe = 0; probe <= mask; probe++) {
        unsigned long long old = keys[s];
        if (old != key) {
            if (old != DT_EMPTY) { s = (s + 1) & mask; continue; }
            old = atomicCAS(&keys[s], DT_EMPTY, key);
            if (old != DT_EMPTY && old != key) { s = (s + 1) & mask; continue; }
        }
        atomicAdd(&vals[s], cnt);
        return true;
    }
    return false;
}

__device__ __forceinline__ int dt_get(const unsigned long long *keys, const int *vals, unsigned mask,
                                      unsigned long long key) {
    unsigned s = dt_hash(key) & mask;
    for (unsigned probe = 0; probe <= mask; probe++) {
        unsigned long long old = keys[s];
        if (old == key) return vals[s];
        if (old == DT_EMPTY) return 0;
        s = (s + 1) & mask;
    }
    return 0;
}

// add `key` once per distinct value in the warp, with the number of lanes holding it
__device__ __forceinline__ bool dt_add_aggregated(unsigned long long *keys, int *vals, unsigned mask,
                                                  unsigned long long key, bool active) {
    unsigned amask = __ballot_sync(0xffffffffu, active);
    bool ok = true;
    if (active) {
        unsigned peers = __match_any_sync(amask, key);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) ok = dt_add(keys, vals, mask, key, __popc(peers));
    }
    return ok;
}

__device__ __forceinline__ long long dt_score(const unsigned long long *keys, const int *vals, unsigned mask,
                                              int lx, int ly) {
    long long s = 0;
    if (lx != 0) s += dt_get(keys, vals, mask, ((unsigned long long)(unsigned)(lx + 1) << 32));
    if (ly != 0) s += dt_get(keys, vals, mask, (unsigned long long)(unsigned)(ly + 1));
    if (lx != 0 && ly != 0)
        s -= 2LL * dt_get(keys, vals, mask, ((unsigned long long)(unsigned)(lx + 1) << 32) | (unsigned)(ly + 1));
    return s;
}

// GLOBAL_TABLE = false: table in shared memory; a pair whose labels do not fit is appended to
// `overflow` (count in overflow[0]) and left for a second launch with GLOBAL_TABLE = true, where
// `pair_list` names the pairs and each CTA owns `slots` entries of a global table.
template <bool GLOBAL_TABLE>
__global__ void __launch_bounds__(DT_THREADS)
difft_kernel(const int *__restrict__ labx, const int *__restrict__ laby, int L, int npairs,
             double *__restrict__ out, unsigned slots,
             unsigned long long *__restrict__ gkeys, int *__restrict__ gvals,
             const int *__restrict__ pair_list, int *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = GLOBAL_TABLE ? gkeys + (size_t)blockIdx.x * slots : (unsigned long long *)smem_raw;
    int *vals = GLOBAL_TABLE ? gvals + (size_t)blockIdx.x * slots : (int *)(keys + slots);
    const unsigned mask = slots - 1;
    __shared__ long long s_warp[DT_THREADS / 32];
    __shared__ long long s_carry;
    __shared__ long long s_total;
    __shared__ int s_overflow;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int pi = blockIdx.x; pi < npairs; pi += gridDim.x) {
        const int pair = pair_list ? pair_list[pi] : pi;
        const int *lx = labx + (size_t)pair * L, *ly = laby + (size_t)pair * L;
        double *o = out + (size_t)pair * L;
        {
            for (unsigned s = tid; s <= mask; s += DT_THREADS) { keys[s] = DT_EMPTY; vals[s] = 0; }
            if (tid == 0) { s_overflow = 0; s_total = 0; s_carry = 0; }
            __syncthreads();
            // pass 1: histograms
            bool ok = true;
            for (int b0 = 0; b0 < L; b0 += DT_THREADS) {
                const int b = b0 + tid;
                const bool act = b < L;
                const int x = act ? lx[b] : 0, y = act ? ly[b] : 0;
                ok &= dt_add_aggregated(keys, vals, mask, ((unsigned long long)(unsigned)(x + 1) << 32), act && x != 0);
                ok &= dt_add_aggregated(keys, vals, mask, (unsigned long long)(unsigned)(y + 1), act && y != 0);
                ok &= dt_add_aggregated(keys, vals, mask,
                                        ((unsigned long long)(unsigned)(x + 1) << 32) | (unsigned)(y + 1),
                                        act && x != 0 && y != 0);
            }
            if (!ok) s_overflow = 1;
            __syncthreads();
            if (s_overflow) {           // table too small for this pair: leave it to the second launch
                if (tid == 0 && overflow) overflow[1 + atomicAdd(&overflow[0], 1)] = pair;
                __syncthreads();
                continue;
            }
        }
        // total = sum over bins of score = sum over joint keys of count * score(key), plus the bins
        // with a zero label on one side; simplest exact form: one more pass over the bins.
        long long part = 0;
        for (int b = tid; b < L; b += DT_THREADS) part += dt_score(keys, vals, mask, lx[b], ly[b]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane == 0) atomicAdd((unsigned long long *)&s_total, (unsigned long long)part);
        __syncthreads();
        const long long total = s_total;
        const double dtotal = (double)total;
        // pass 2: per-bin score, inclusive scan, normalise
        for (int b0 = 0; b0 < L; b0 += DT_THREADS) {
            const int b = b0 + tid;
            long long v = (b < L) ? dt_score(keys, vals, mask, lx[b], ly[b]) : 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                long long u = __shfl_up_sync(0xffffffffu, v, off);
                if (lane >= off) v += u;
            }
            if (lane == 31) s_warp[wid] = v;
            __syncthreads();
            long long base = s_carry;
            for (int w = 0; w < wid; w++) base += s_warp[w];
            v += base;
            if (b < L) o[b] = (total != 0) ? (double)v / dtotal : (double)v;
            __syncthreads();
            if (tid == DT_THREADS - 1) s_carry = v;
            __syncthreads();
        }
        __syncthreads();
    }
}

int tp_difft_batch(tp_ctx *ctx, const int32_t *labels_x, const int32_t *labels_y, int L, int npairs,
                   int on_device, double *out) {
    TP_ARG(ctx && labels_x && labels_y && out, "tp_difft_batch: null argument");
    TP_ARG(L >= 1 && npairs >= 0, "tp_difft_batch: bad sizes");
    if (npairs == 0) return TP_OK;
    TP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nl = (size_t)npairs * L;
    const int *dx = labels_x, *dy = labels_y;
    double *dout = out;
    if (!on_device) {
        TP_TRY(ctx->lx.reserve(nl * sizeof(int)));
        TP_TRY(ctx->ly.reserve(nl * sizeof(int)));
        TP_TRY(ctx->dout.reserve(nl * sizeof(double)));
        TP_CUDA(cudaMemcpyAsync(ctx->lx.p, labels_x, nl * sizeof(int), cudaMemcpyHostToDevice, st));
        TP_CUDA(cudaMemcpyAsync(ctx->ly.p, labels_y, nl * sizeof(int), cudaMemcpyHostToDevice, st));
        dx = ctx->lx.as<int>(); dy = ctx->ly.as<int>(); dout = ctx->dout.as<double>();
    }
    const unsigned smem_slots = 4096;                       // 48 KB: 4 CTAs per SM
    const size_t smem = (size_t)smem_slots * 12;
    int grid = npairs < ctx->sm_count * 4 ? npairs : ctx->sm_count * 4;
    TP_TRY(ctx->dhash.reserve((size_t)(npairs + 1) * sizeof(int)));
    int *d_over = ctx->dhash.as<int>();
    TP_CUDA(cudaMemsetAsync(d_over, 0, sizeof(int), st));
    TP_CUDA(cudaFuncSetAttribute(difft_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TP_MARK(ctx, EV_DIFFT0);
    tp_prof_begin(ctx, PC_DIFFT);
    difft_kernel<false><<<grid, DT_THREADS, smem, st>>>(dx, dy, L, npairs, dout, smem_slots, nullptr, nullptr,
                                                       nullptr, d_over);
    tp_prof_end(ctx);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    int nover = 0;
    TP_CUDA(cudaMemcpyAsync(&nover, d_over, sizeof(int), cudaMemcpyDeviceToHost, st));
    TP_CUDA(cudaStreamSynchronize(st));
    if (nover > 0) {
        // pairs with more distinct labels than the shared table holds: global tables, <= 3L keys
        unsigned gslots = 1024;
        while (gslots < 6u * (unsigned)L) gslots <<= 1;
        int grid2 = nover < 64 ? nover : 64;
        TP_TRY(ctx->lhash.reserve((size_t)grid2 * gslots * 12));
        unsigned long long *gkeys = ctx->lhash.as<unsigned long long>();
        int *gvals = (int *)(gkeys + (size_t)grid2 * gslots);
        difft_kernel<true><<<grid2, DT_THREADS, 0, st>>>(dx, dy, L, nover, dout, gslots, gkeys, gvals,
                                                        d_over + 1, nullptr);
        ctx->launches += 1;
        TP_CUDA(cudaGetLastError());
    }
    TP_MARK(ctx, EV_DIFFT1);
    if (!on_device) TP_CUDA(cudaMemcpyAsync(out, dout, nl * sizeof(double), cudaMemcpyDeviceToHost, st));
    TP_CUDA(cudaStreamSynchronize(st));
    return TP_OK;
}
