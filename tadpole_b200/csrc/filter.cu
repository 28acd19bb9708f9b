// filter.cu -- stage 1: the numeric core of load_mat (reference R/TADpole.R:19-22,35-37,88).
//
//   mat[is.na(mat)] <- 0 ; mat <- forceSymmetric(mat, uplo='U')          (:19-20)
//   r <- rowMeans(mat) ; bad <- diag(mat) == 0 | r < quantile(r, bad_frac)  (:35-37)
//   mat <- mat[!bad, !bad]                                                  (:88)
//
// The symmetrised matrix is never materialised: only the upper triangle U(i,j), i <= j, of the
// caller's buffer is read.  With A the row-major view of the buffer, U(i,j) = A[i][j] for a C /
// numpy buffer (upper triangle of A) and U(i,j) = A[j][i] for an R column-major buffer (lower
// triangle of A).  Row i of the symmetric matrix is then one contiguous row segment of A plus one
// column segment of A; a CTA owning R consecutive rows streams an R-row horizontal strip and an
// R-column vertical strip, both in full 32-byte sectors, so the pass reads N^2 doubles in total
// (every off-diagonal element of the triangle twice, once per incident row) with a fixed
// summation order -- no atomics, bit-reproducible.  HBM-bound: 8 N^2 bytes.
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <thread>

#define FT_THREADS 256

// upper == true : valid triangle is c >= r of A ; row i = A[i][i..n) + A[0..i)[i]
// upper == false: valid triangle is c <= r of A ; row i = A[i][0..i] + A(i..n)[i]
template <int R>
__global__ void __launch_bounds__(FT_THREADS)
rowmean_kernel(const double *__restrict__ A, int n, int upper, double *__restrict__ rowmean,
               unsigned char *__restrict__ diag0) {
    __shared__ double s_h[R];
    __shared__ double s_v[FT_THREADS / R][R];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int i0 = blockIdx.x * R;
    // horizontal segments: one warp per row, lanes along the row
    for (int rr = wid; rr < R; rr += FT_THREADS / 32) {
        const int i = i0 + rr;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        if (i < n) {
            const double *row = A + (size_t)i * n;
            const int c0 = upper ? i : 0, c1 = upper ? n : i + 1;
            int c = c0 + lane;
            for (; c + 96 < c1; c += 128) {
                acc0 += nan_to_zero(row[c]);
                acc1 += nan_to_zero(row[c + 32]);
                acc2 += nan_to_zero(row[c + 64]);
                acc3 += nan_to_zero(row[c + 96]);
            }
            for (; c < c1; c += 32) acc0 += nan_to_zero(row[c]);
        }
        double acc = warp_sum((acc0 + acc1) + (acc2 + acc3));
        if (lane == 0) s_h[rr] = acc;
    }
    // vertical strip: thread (g, cc) walks rows g, g+G, ... of column i0+cc
    {
        const int G = FT_THREADS / R;
        const int cc = tid % R, g = tid / R;
        const int i = i0 + cc;
        double acc0 = 0.0, acc1 = 0.0;
        if (i < n) {
            const int r0 = upper ? 0 : i + 1, r1 = upper ? i : n;
            // start every column of the strip at the same row so that warps stay coalesced
            const int rs = upper ? 0 : i0 + 1;
            int r = rs + g;
            for (; r + G < r1; r += 2 * G) {
                double a = (r >= r0) ? nan_to_zero(A[(size_t)r * n + i]) : 0.0;
                double b = (r + G >= r0) ? nan_to_zero(A[(size_t)(r + G) * n + i]) : 0.0;
                acc0 += a; acc1 += b;
            }
            for (; r < r1; r += G) if (r >= r0) acc0 += nan_to_zero(A[(size_t)r * n + i]);
        }
        s_v[g][cc] = acc0 + acc1;
    }
    __syncthreads();
    if (tid < R) {
        const int i = i0 + tid;
        if (i < n) {
            double v = 0.0;
            for (int g = 0; g < FT_THREADS / R; g++) v += s_v[g][tid];
            rowmean[i] = (s_h[tid] + v) / (double)n;
            diag0[i] = nan_to_zero(A[(size_t)i * n + i]) == 0.0;
        }
    }
}

// rank by counting: rank[i] = #{j : v_j < v_i} + #{j < i : v_j == v_i}; the two order statistics the
// type-7 quantile needs are the values whose rank is lo-1 and hi-1.
__global__ void __launch_bounds__(256)
order_stat_kernel(const double *__restrict__ v, int n, int rank_lo, int rank_hi, double *__restrict__ out2) {
    __shared__ double tile[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double vi = (i < n) ? v[i] : 0.0;
    int rank = 0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = (j < n) ? v[j] : 0.0;
        __syncthreads();
        const int m = min(256, n - j0);
        for (int t = 0; t < m; t++) {
            const double vj = tile[t];
            rank += (vj < vi) || (vj == vi && (j0 + t) < i);
        }
        __syncthreads();
    }
    if (i < n) {
        if (rank == rank_lo) out2[0] = vi;
        if (rank == rank_hi) out2[1] = vi;
    }
}

// stats::quantile type 7 on the two order statistics, then the flags (R/TADpole.R:36-37)
__global__ void flags_kernel(const double *__restrict__ rowmean, const unsigned char *__restrict__ diag0, int n,
                             int use_q, double index, int lo, const double *__restrict__ xs,
                             unsigned char *__restrict__ bad, double *__restrict__ thr_out) {
    double q = __longlong_as_double(0x7ff8000000000000LL);
    if (use_q) {
        q = xs[0];
        const double xhi = xs[1];
        if (index > (double)lo && xhi != q) {
            const double h = index - (double)lo;
            q = (1.0 - h) * q + h * xhi;
        }
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *thr_out = q;
    if (i < n) bad[i] = diag0[i] | (use_q ? (rowmean[i] < q) : 0);
}

// mat[keep, keep] with NA -> 0 and the lower triangle mirrored from the upper one.
// 32 x 32 tiles of the output's upper triangle; each tile is read once from the valid triangle of
// A and written twice (as is, and transposed through shared memory).
__global__ void __launch_bounds__(256)
compact_kernel(const double *__restrict__ A, int n, int upper, const int *__restrict__ keep, int nf,
               double *__restrict__ X, int ldx) {
    __shared__ double tile[32][33];
    const int ta = blockIdx.y, tb = blockIdx.x;
    if (tb < ta) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int a0 = ta * 32, b0 = tb * 32;
    // tile[y][x] = S(keep[a0+y], keep[b0+x]); choose the thread mapping so that the fast thread index
    // runs along the contiguous direction of A
    for (int s = ty; s < 32; s += 8) {
        int ya, xb;
        if (upper) { ya = s; xb = tx; } else { ya = tx; xb = s; }
        const int a = a0 + ya, b = b0 + xb;
        double v = 0.0;
        if (a < nf && b < nf) {
            int ia = keep[a], ib = keep[b];
            int lo = min(ia, ib), hi = max(ia, ib);
            // U(lo, hi): upper -> A[lo][hi], lower -> A[hi][lo]
            v = upper ? A[(size_t)lo * n + hi] : A[(size_t)hi * n + lo];
            v = nan_to_zero(v);
        }
        tile[ya][xb] = v;
    }
    __syncthreads();
    for (int s = ty; s < 32; s += 8) {
        const int a = a0 + s, b = b0 + tx;
        if (a < nf && b < nf) X[(size_t)a * ldx + b] = tile[s][tx];
        if (ta != tb) {
            const int bb = b0 + s, aa = a0 + tx;
            if (bb < nf && aa < nf) X[(size_t)bb * ldx + aa] = tile[tx][s];
        }
    }
}

// zero the padding columns [nf, ldx) of an nf x ldx matrix
__global__ void zero_pad_kernel(double *__restrict__ X, int nf, int ldx) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nf) for (int c = nf; c < ldx; c++) X[(size_t)r * ldx + c] = 0.0;
}

template <int R>
static void launch_rowmean(tp_ctx *ctx, const double *A, int n, int upper, double *rm, unsigned char *d0) {
    rowmean_kernel<R><<<(n + R - 1) / R, FT_THREADS, 0, ctx->stream>>>(A, n, upper, rm, d0);
}

// rows (columns, for R's layout) per upload band: ~32 bands
static int upload_bands(int n) {
    static const int forced = getenv("TADPOLE_UPLOAD_BANDS") ? atoi(getenv("TADPOLE_UPLOAD_BANDS")) : 0;     // experiments
    const int nb = forced > 0 ? forced : 32;
    return n / nb > 64 ? n / nb : 64;
}

#define LANE_BYTES ((size_t)8 << 20)

// A pageable host matrix goes through the driver's single staging buffer at the speed of ONE memcpy (~10 GB/s measured:
// 250 ms for the 2.6 GB upper triangle of a 25 000-bin matrix).  Above 64 MB the pieces of the bands are instead copied by a few
// helper threads into their own pinned double buffers and sent from there on their own streams; `st` then waits for the lanes.
struct UploadPiece { size_t dst_off, src_off, width; int height; };       // offsets and width in bytes

static int upload_lanes_run(tp_ctx *ctx, void *dst_, const void *src_, size_t pitch, const std::vector<UploadPiece> &pieces,
                            cudaStream_t st) {
    const int L = ctx->upload_lanes;
    if ((int)ctx->lanes.size() < L) {
        const size_t have = ctx->lanes.size();
        ctx->lanes.resize((size_t)L);
        for (size_t l = have; l < (size_t)L; l++) {
            tp_ctx::UploadLane &ln = ctx->lanes[l];
            TP_CUDA(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking));
            TP_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
            for (int b = 0; b < 2; b++) {
                TP_CUDA(cudaMallocHost(&ln.pin[b], LANE_BYTES));
                TP_CUDA(cudaEventCreateWithFlags(&ln.ev[b], cudaEventDisableTiming));
            }
        }
    }
    char *dst = (char *)dst_;
    const char *host = (const char *)src_;
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    auto lane = [&](int l) {
        tp_ctx::UploadLane &ln = ctx->lanes[(size_t)l];
        if (cudaSetDevice(ctx->device) != cudaSuccess) { failed = 1; return; }
        int used = 0;
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= pieces.size() || failed) break;
            const UploadPiece &pc = pieces[i];
            const int b = used & 1;
            // the buffer may still be in flight from this run or from an earlier one (a never-recorded event is complete)
            if (cudaEventSynchronize(ln.ev[b]) != cudaSuccess) { failed = 1; break; }
            char *pin = (char *)ln.pin[b];
            const char *src = host + pc.src_off;
            for (int r = 0; r < pc.height; r++) memcpy(pin + (size_t)r * pc.width, src + (size_t)r * pitch, pc.width);
            const cudaError_t ce = pc.height == 1
                ? cudaMemcpyAsync(dst + pc.dst_off, pin, pc.width, cudaMemcpyHostToDevice, ln.st)
                : cudaMemcpy2DAsync(dst + pc.dst_off, pitch, pin, pc.width, pc.width, (size_t)pc.height, cudaMemcpyHostToDevice, ln.st);
            if (ce != cudaSuccess || cudaEventRecord(ln.ev[b], ln.st) != cudaSuccess) { failed = 1; break; }
            used++;
        }
        if (cudaEventRecord(ln.done, ln.st) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> th;
    for (int l = 1; l < L; l++) th.emplace_back(lane, l);
    lane(0);
    for (std::thread &t : th) t.join();
    if (failed) { tp_set_error("upload: a staging lane failed (%s)", cudaGetErrorString(cudaGetLastError())); return TP_ERR_CUDA; }
    for (int l = 0; l < L; l++) TP_CUDA(cudaStreamWaitEvent(st, ctx->lanes[(size_t)l].done, 0));
    return TP_OK;
}

static bool host_is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// A contiguous host range -> device: through the lanes when it is pageable and large, else one plain copy.
int tp_upload_range(tp_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return TP_OK;
    if (ctx->upload_lanes > 0 && bytes >= ((size_t)64 << 20) && host_is_pageable(src)) {
        std::vector<UploadPiece> pieces;
        for (size_t off = 0; off < bytes; off += LANE_BYTES) pieces.push_back({off, off, std::min(LANE_BYTES, bytes - off), 1});
        return upload_lanes_run(ctx, dst, src, 0, pieces, st);
    }
    TP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return TP_OK;
}

// Upper triangle of a host matrix -> dst, ~32 band copies (row bands from the diagonal to the right edge; column bands from
// the top to the diagonal for R's layout).  R > 1: only the bands dealt to `rank` (boustrophedon: equal bytes per rank).
static int upload_upper(tp_ctx *ctx, double *dst, const double *mat, int n, int colmajor, cudaStream_t st, int R, int rank) {
    const int band = upload_bands(n);
    const size_t pitch = (size_t)n * sizeof(double);
    auto owner_of = [&](int bi) { const int blk = bi / R, pos = bi % R; return (blk & 1) ? R - 1 - pos : pos; };
    int bi = 0;
    if (ctx->upload_lanes > 0 && (size_t)n * n * 4 / (size_t)R >= ((size_t)64 << 20) && host_is_pageable(mat)) {
        std::vector<UploadPiece> pieces;
        for (int r = 0; r < n; r += band, bi++) {
            if (R > 1 && owner_of(bi) != rank) continue;
            const int h = n - r < band ? n - r : band;
            // row-major: rows [r, r + h) from column r; column-major: columns [r, r + h) down to row r + h
            const size_t off = (colmajor ? (size_t)r * n : (size_t)r * n + r) * sizeof(double);
            const size_t width = (colmajor ? (size_t)(r + h) : (size_t)(n - r)) * sizeof(double);
            const int per = (int)std::max<size_t>(1, LANE_BYTES / width);
            for (int q = 0; q < h; q += per)
                pieces.push_back({off + (size_t)q * pitch, off + (size_t)q * pitch, width, std::min(per, h - q)});
        }
        return upload_lanes_run(ctx, dst, mat, pitch, pieces, st);
    }
    bi = 0;
    for (int r = 0; r < n; r += band, bi++) {
        if (R > 1 && owner_of(bi) != rank) continue;
        const int h = n - r < band ? n - r : band;
        if (!colmajor)   // rows [r, r + h), columns [r, n)
            TP_CUDA(cudaMemcpy2DAsync(dst + (size_t)r * n + r, pitch, mat + (size_t)r * n + r, pitch,
                                      (size_t)(n - r) * sizeof(double), h, cudaMemcpyHostToDevice, st));
        else             // columns [r, r + h), rows [0, r + h)
            TP_CUDA(cudaMemcpy2DAsync(dst + (size_t)r * n, pitch, mat + (size_t)r * n, pitch,
                                      (size_t)(r + h) * sizeof(double), h, cudaMemcpyHostToDevice, st));
    }
    return TP_OK;
}

// Batch pool: start the upload of the matrix of the NEXT call on the context's copy stream, into the second input buffer,
// while the current call computes; tp_filter adopts it when it is handed the same (pointer, n, layout).
int tp_stage_input(tp_ctx *ctx, const double *mat, int n, int colmajor) {
    TP_ARG(ctx && mat && n >= 2, "tp_stage_input: bad arguments");
    if (ctx->group || tp_nranks(ctx) > 1) return TP_OK;         // single-device contexts only
    TP_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) {
        TP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        TP_CUDA(cudaEventCreateWithFlags(&ctx->staged_ev, cudaEventDisableTiming));
    }
    ctx->staged_mat = nullptr;
    TP_TRY(ctx->raw_next.reserve((size_t)n * n * sizeof(double)));
    TP_TRY(upload_upper(ctx, ctx->raw_next.as<double>(), mat, n, colmajor, ctx->copy_stream, 1, 0));
    TP_CUDA(cudaEventRecord(ctx->staged_ev, ctx->copy_stream));
    ctx->staged_mat = mat; ctx->staged_n = n; ctx->staged_colmajor = colmajor ? 1 : 0;
    return TP_OK;
}

int tp_filter(tp_ctx *ctx, const double *mat, int n, int colmajor, int on_device, double bad_frac,
              uint8_t *bad_out, double *rowmeans_out, double *thr_out) {
    TP_ARG(ctx && mat && bad_out, "tp_filter: null argument");
    TP_ARG(n >= 2, "tp_filter: matrix must be at least 2 x 2");
    TP_ARG(bad_frac >= 0.0 && bad_frac <= 1.0, "tp_filter: bad_frac must be in [0, 1]");
    if (tp_group_dispatch(ctx))      // multi-device context: every member device runs the stage; rank 0 reports
        return tp_group_run(ctx, [&](tp_ctx *gc, int gr) -> int {
            std::vector<uint8_t> tmp;
            if (gr) tmp.resize((size_t)n);
            return tp_filter(gc, mat, n, colmajor, on_device, bad_frac, gr ? tmp.data() : bad_out, gr ? nullptr : rowmeans_out,
                             gr ? nullptr : thr_out);
        });
    TP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->generation++;
    const size_t bytes = (size_t)n * n * sizeof(double);
    const int R = tp_nranks(ctx), rank = tp_rank(ctx);
    if (on_device && ctx->group && R > 1) {
        // multi-device context: `mat` lives on the first device; the others receive it over NVLink
        ctx->ingested_n = rank == 0 ? ctx->ingested_n : 0;
        double *dst = const_cast<double *>(mat);
        if (rank != 0) { TP_TRY(ctx->raw_own.reserve(bytes)); dst = ctx->raw_own.as<double>(); }
        TP_TRY(tp_comm_bcast_bytes(ctx, dst, bytes, 0));
        ctx->raw = dst;
    } else if (on_device) {
        ctx->raw = mat;
    } else if (ctx->staged_mat == mat && ctx->staged_n == n && ctx->staged_colmajor == (colmajor ? 1 : 0) && R == 1) {
        // the batch worker uploaded this matrix under the previous call's compute: swap the input buffers
        ctx->ingested_n = 0;
        ctx->staged_mat = nullptr;
        std::swap(ctx->raw_own.p, ctx->raw_next.p);
        std::swap(ctx->raw_own.cap, ctx->raw_next.cap);
        TP_CUDA(cudaStreamWaitEvent(st, ctx->staged_ev, 0));
        ctx->raw = ctx->raw_own.as<double>();
    } else {
        // Only the upper triangle is ever read (forceSymmetric(uplo = 'U')), so only it crosses PCIe: ~32 band copies
        // (row bands from the diagonal to the right edge; column bands from the top to the diagonal for R's layout),
        // 52 % of the bytes of the full matrix.
        // Several ranks working on the same matrix (a call spread over GPUs) upload it ONCE between them: the bands are
        // dealt out boustrophedon (equal bytes per rank), every rank pulls its bands over its own PCIe link, then each
        // band is broadcast from its owner over NVLink.
        ctx->ingested_n = 0;                       // raw_own is about to be overwritten
        TP_TRY(ctx->raw_own.reserve(bytes));
        double *dst = ctx->raw_own.as<double>();
        const int band = upload_bands(n);
        const bool share = R > 1 && n >= ctx->dist_min_n;
        auto owner_of = [&](int bi) { const int blk = bi / R, pos = bi % R; return (blk & 1) ? R - 1 - pos : pos; };
        int bi = 0;
        TP_TRY(upload_upper(ctx, dst, mat, n, colmajor, st, share ? R : 1, rank));
        if (share) {
            // the memory between the first and the last uploaded element of a band is one contiguous range
            TP_TRY(tp_comm_group_begin(ctx));
            bi = 0;
            for (int r = 0; r < n; r += band, bi++) {
                const int h = n - r < band ? n - r : band;
                double *b0 = colmajor ? dst + (size_t)r * n : dst + (size_t)r * n + r;
                double *b1 = colmajor ? dst + (size_t)(r + h - 1) * n + (r + h) : dst + (size_t)(r + h) * n;
                TP_TRY(tp_comm_bcast_bytes(ctx, b0, (size_t)(b1 - b0) * sizeof(double), owner_of(bi)));
            }
            TP_TRY(tp_comm_group_end(ctx));
        }
        ctx->raw = dst;
    }
    ctx->n = n;
    ctx->colmajor = colmajor ? 1 : 0;
    ctx->have_X = ctx->have_C = ctx->have_scores = ctx->have_sweep = false;
    const int upper = colmajor ? 0 : 1;
    TP_TRY(ctx->rowmean.reserve((size_t)n * sizeof(double)));
    TP_TRY(ctx->flags.reserve((size_t)2 * n));
    TP_TRY(ctx->qtmp.reserve(4 * sizeof(double)));
    double *rm = ctx->rowmean.as<double>();
    unsigned char *diag0 = ctx->flags.as<unsigned char>();
    unsigned char *bad = diag0 + n;
    double *xs = ctx->qtmp.as<double>();

    TP_MARK(ctx, EV_FILTER0);
    // rows per CTA: keep the grid at >= ~2 waves of 148 SMs when the matrix allows it
    const int want = 2 * ctx->sm_count;
    tp_prof_begin(ctx, PC_ROWMEAN);
    if (n / 64 >= want) launch_rowmean<64>(ctx, ctx->raw, n, upper, rm, diag0);
    else if (n / 32 >= want) launch_rowmean<32>(ctx, ctx->raw, n, upper, rm, diag0);
    else if (n / 16 >= want) launch_rowmean<16>(ctx, ctx->raw, n, upper, rm, diag0);
    else launch_rowmean<8>(ctx, ctx->raw, n, upper, rm, diag0);
    tp_prof_end(ctx);
    ctx->launches += 1;
    int use_q = bad_frac != 0.0;
    double index = 1.0;
    int lo = 1, hi = 1;
    if (use_q) {
        index = 1.0 + (double)(n - 1) * bad_frac;      // R: 1 + max(n - 1, 0) * probs
        lo = (int)floor(index);
        hi = (int)ceil(index);
        order_stat_kernel<<<(n + 255) / 256, 256, 0, st>>>(rm, n, lo - 1, hi - 1, xs);
        ctx->launches += 1;
    }
    flags_kernel<<<(n + 255) / 256, 256, 0, st>>>(rm, diag0, n, use_q, index, lo, xs, bad, xs + 2);
    ctx->launches += 1;
    TP_CUDA(cudaGetLastError());
    TP_MARK(ctx, EV_FILTER1);
    TP_CUDA(cudaMemcpyAsync(bad_out, bad, n, cudaMemcpyDeviceToHost, st));
    if (rowmeans_out) TP_CUDA(cudaMemcpyAsync(rowmeans_out, rm, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (thr_out) TP_CUDA(cudaMemcpyAsync(thr_out, xs + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
    TP_CUDA(tp_stream_sync(ctx));
    return TP_OK;
}

int tp_compact(tp_ctx *ctx, const int *keep, int nf) {
    TP_ARG(ctx && keep, "tp_compact: null argument");
    if (tp_group_dispatch(ctx)) return tp_group_run(ctx, [&](tp_ctx *gc, int) -> int { return tp_compact(gc, keep, nf); });
    TP_ARG(ctx->raw && ctx->n > 0, "tp_compact: call tp_filter first");
    TP_ARG(nf >= 2 && nf <= ctx->n, "tp_compact: bad keep count");
    ctx->generation++;
    for (int i = 0; i < nf; i++)
        TP_ARG(keep[i] >= 0 && keep[i] < ctx->n, "tp_compact: keep index out of range");
    TP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int ldx = round_up(nf, 8);
    TP_TRY(ctx->keep.reserve((size_t)nf * sizeof(int)));
    TP_TRY(ctx->X.reserve((size_t)nf * ldx * sizeof(double)));
    TP_CUDA(cudaMemcpyAsync(ctx->keep.p, keep, (size_t)nf * sizeof(int), cudaMemcpyHostToDevice, st));
    TP_MARK(ctx, EV_COMPACT0);
    const int nt = (nf + 31) / 32;
    tp_prof_begin(ctx, PC_COMPACT);
    compact_kernel<<<dim3(nt, nt), 256, 0, st>>>(ctx->raw, ctx->n, ctx->colmajor ? 0 : 1, ctx->keep.as<int>(), nf,
                                                 ctx->X.as<double>(), ldx);
    tp_prof_end(ctx);
    if (ldx > nf) zero_pad_kernel<<<(nf + 127) / 128, 128, 0, st>>>(ctx->X.as<double>(), nf, ldx);
    ctx->launches += (ldx > nf) ? 2 : 1;
    TP_CUDA(cudaGetLastError());
    TP_MARK(ctx, EV_COMPACT1);
    ctx->nf = nf;
    ctx->ldx = ldx;
    ctx->have_X = true;
    ctx->have_C = ctx->have_scores = ctx->have_sweep = false;
    return TP_OK;
}
