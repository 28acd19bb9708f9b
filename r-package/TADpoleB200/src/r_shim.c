/* r_shim.c -- the .Call layer between R and libtadpole_b200 (include/tadpole_b200.h).
 *
 * The only translation unit that includes Rinternals.h.  Every entry point converts R objects to plain pointers
 * and sizes, calls the C ABI, and converts back.  Rules:
 *   - R owns every host buffer; every SEXP allocated here is PROTECTed until it is reachable from the result;
 *   - error() (a longjmp) is raised only after the core has returned, when no C++ frame is live;
 *   - the core never calls the R API and never keeps a host pointer past the call;
 *   - device memory belongs to the tp_ctx behind an external pointer whose finalizer destroys it.
 * Replaces, in the reference: R/TADpole.R:17 (read.big.matrix), :19-22,35-37 (bad columns), :362-374 / :448-460
 * (cor, prcomp, find_params, final chclust), :470-497 (per-level tables), R/DiffT.R:41-49 (diffT loop) and
 * R/DiffT.R:61-73 (random_bed, batched).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>
#include "tadpole_b200.h"

static tp_ctx *ctx_of(SEXP p) {
    tp_ctx *c = (tp_ctx *)R_ExternalPtrAddr(p);
    if (!c) error("TADpole: the GPU context has been released");
    return c;
}

static void ctx_finalizer(SEXP p) {
    tp_ctx *c = (tp_ctx *)R_ExternalPtrAddr(p);
    if (c) {
        tp_ctx_destroy(c);
        R_ClearExternalPtr(p);
    }
}

/* devices: one index, or several -> one context over those GPUs driven from this one thread (options(tadpole.gpus=)) */
SEXP C_tp_ctx(SEXP devices) {
    tp_ctx *c = NULL;
    int ndev = length(devices);
    if (ndev < 1) error("TADpole: no device given");
    int *devs = (int *)R_alloc((size_t)ndev, sizeof(int));
    if (ndev == 1) devs[0] = asInteger(devices);
    else for (int i = 0; i < ndev; i++) devs[i] = isReal(devices) ? (int)REAL(devices)[i] : INTEGER(devices)[i];
    if (tp_ctx_create_multi(devs, ndev, &c) != TP_OK) error("%s", tp_last_error());
    SEXP p = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* the counter that changes whenever the state resident in the context is replaced (kept in attr(, 'resident')) */
SEXP C_tp_generation(SEXP ctx) {
    SEXP g = PROTECT(allocVector(REALSXP, 1));
    REAL(g)[0] = (double)tp_ctx_generation(ctx_of(ctx));
    UNPROTECT(1);
    return g;
}

/* stop() unless the context still holds the state the R object was made from */
static void check_generation(tp_ctx *c, SEXP generation) {
    if ((double)tp_ctx_generation(c) != asReal(generation))
        error("TADpole: the GPU context has been used for another matrix since this call; its resident state is gone");
}

/* read.big.matrix(mat_file, type = 'double', sep = '\t') on the device (R/TADpole.R:17): returns N */
SEXP C_tp_ingest(SEXP ctx, SEXP path) {
    int n = 0;
    if (tp_ingest_tsv_file(ctx_of(ctx), CHAR(STRING_ELT(path, 0)), '\t', &n) != TP_OK) error("%s", tp_last_error());
    return ScalarInteger(n);
}

/* Upper-triangle pixels (bin1, bin2, count) -> the dense matrix in HBM (tp_ingest_coo): a Matrix::sparseMatrix, or the
 * columns of a `cooler dump`, without the dense N x N matrix ever existing in R.  Returns c(N, pixels below the
 * diagonal that were ignored). */
SEXP C_tp_ingest_coo(SEXP ctx, SEXP bin1, SEXP bin2, SEXP count, SEXP n, SEXP index_base) {
    if (!isInteger(bin1) || !isInteger(bin2) || !isReal(count) || XLENGTH(bin1) != XLENGTH(bin2) || XLENGTH(bin1) != XLENGTH(count))
        error("TADpole: bin1, bin2 (integer) and count (numeric) of equal length are expected");
    unsigned long long below = 0;
    if (tp_ingest_coo(ctx_of(ctx), INTEGER(bin1), INTEGER(bin2), REAL(count), (size_t)XLENGTH(bin1), asInteger(n),
                      asInteger(index_base), &below) != TP_OK)
        error("%s", tp_last_error());
    SEXP out = PROTECT(allocVector(REALSXP, 2));
    REAL(out)[0] = (double)asInteger(n);
    REAL(out)[1] = (double)below;
    UNPROTECT(1);
    return out;
}

/* the same from a three-column text file "bin1 <tab> bin2 <tab> count" parsed on the device; n <= 0: largest bin + 1.
 * Returns c(N, pixels, pixels below the diagonal). */
SEXP C_tp_ingest_coo_file(SEXP ctx, SEXP path, SEXP n, SEXP index_base) {
    int nn = 0;
    unsigned long long nnz = 0, below = 0;
    if (tp_ingest_coo_file(ctx_of(ctx), CHAR(STRING_ELT(path, 0)), '\t', asInteger(n), asInteger(index_base), &nn, &nnz,
                           &below) != TP_OK)
        error("%s", tp_last_error());
    SEXP out = PROTECT(allocVector(REALSXP, 3));
    REAL(out)[0] = (double)nn;
    REAL(out)[1] = (double)nnz;
    REAL(out)[2] = (double)below;
    UNPROTECT(1);
    return out;
}

/* the ingested matrix as an R matrix (for the plots of load_mat, R/TADpole.R:24-53); symmetric use only needs the
 * upper triangle, so the row-major device matrix is handed back transposed in place of a copy loop */
SEXP C_tp_ingested_matrix(SEXP ctx) {
    const double *dev = NULL;
    int n = 0;
    if (tp_ingested(ctx_of(ctx), &dev, &n) != TP_OK) error("%s", tp_last_error());
    SEXP m = PROTECT(allocMatrix(REALSXP, n, n));
    if (tp_get_ingested(ctx_of(ctx), REAL(m)) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    /* row-major n x n read as column-major is the transpose: fix it up so that m[i, j] is the file's field (i, j) */
    double *a = REAL(m);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            double t = a[i + (size_t)j * n];
            a[i + (size_t)j * n] = a[j + (size_t)i * n];
            a[j + (size_t)i * n] = t;
        }
    UNPROTECT(1);
    return m;
}

/* numeric core of load_mat (R/TADpole.R:19-22,35-37): logical vector of bad columns.
 * mat = an R numeric matrix (column-major, host), or NULL to use the matrix ingested by C_tp_ingest. */
SEXP C_tp_filter(SEXP ctx, SEXP mat, SEXP bad_frac) {
    tp_ctx *c = ctx_of(ctx);
    const double *ptr;
    int n, colmajor, on_device;
    if (mat == R_NilValue) {
        if (tp_ingested(c, &ptr, &n) != TP_OK) error("%s", tp_last_error());
        colmajor = 0; on_device = 1;
    } else {
        if (!isReal(mat) || nrows(mat) != ncols(mat)) error("TADpole: a square numeric matrix is expected");
        ptr = REAL(mat); n = nrows(mat); colmajor = 1; on_device = 0;
    }
    SEXP bad = PROTECT(allocVector(LGLSXP, n));
    unsigned char *tmp = (unsigned char *)R_alloc(n, 1);
    int rc = tp_filter(c, ptr, n, colmajor, on_device, asReal(bad_frac), tmp, NULL, NULL);
    if (rc != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    for (int i = 0; i < n; i++) LOGICAL(bad)[i] = tmp[i];
    UNPROTECT(1);
    return bad;
}

/* list(n_pcs, optimal_n_clusters, seqdist, scores) from the outputs of tp_call_arm / tp_recall */
static SEXP pack_call(tp_ctx *c, int rc, int k, int npcs, int ncl, int maxlev, int ld, double *sc, int kmax, SEXP seq) {
    if (rc == TP_ERR_ARG && maxlev > ld) {       /* more levels than the buffer: everything else is filled in */
        ld = maxlev;
        sc = (double *)R_alloc((size_t)kmax * ld, sizeof(double));
        rc = tp_get_sweep_scores(c, sc, ld);
    }
    if (rc != TP_OK) error("%s", tp_last_error());
    SEXP scores = PROTECT(allocMatrix(REALSXP, k, maxlev));        /* column-major for R, NaN padding -> NA */
    for (int r = 0; r < k; r++)
        for (int l = 0; l < maxlev; l++) {
            double v = sc[(size_t)r * ld + l];
            REAL(scores)[r + (size_t)l * k] = ISNAN(v) ? NA_REAL : v;
        }
    SEXP out = PROTECT(allocVector(VECSXP, 5));
    SET_VECTOR_ELT(out, 0, ScalarInteger(npcs));
    SET_VECTOR_ELT(out, 1, ScalarInteger(ncl));
    SET_VECTOR_ELT(out, 2, seq);
    SET_VECTOR_ELT(out, 3, scores);
    SEXP gen = PROTECT(allocVector(REALSXP, 1));
    REAL(gen)[0] = (double)tp_ctx_generation(c);          /* the state this result was read from */
    SET_VECTOR_ELT(out, 4, gen);
    UNPROTECT(3);
    return out;
}

/* cor -> prcomp -> find_params -> final chclust for one matrix or one arm (R/TADpole.R:362-374, 448-460);
 * keep = 0-based original indices of the rows load_mat keeps */
SEXP C_tp_call_arm(SEXP ctx, SEXP keep, SEXP max_pcs, SEXP min_clusters) {
    tp_ctx *c = ctx_of(ctx);
    int nf = length(keep), k = 0, npcs = 0, ncl = 0, maxlev = 0, ld = 256;
    if (nf < 3) error("TADpole: fewer than 3 good bins left after filtering");
    int kmax = asInteger(max_pcs) < nf ? asInteger(max_pcs) : nf;
    SEXP seq = PROTECT(allocVector(REALSXP, nf - 1));
    double *sc = (double *)R_alloc((size_t)kmax * ld, sizeof(double));
    int rc = tp_call_arm(c, INTEGER(keep), nf, asInteger(max_pcs), asInteger(min_clusters), &k, &npcs, &ncl, sc, ld,
                         &maxlev, REAL(seq));
    SEXP out = pack_call(c, rc, k, npcs, ncl, maxlev, ld, sc, kmax, seq);
    UNPROTECT(1);
    return out;
}

/* the sweep again on the resident PC scores with another max_pcs / min_clusters (no reference counterpart).  The sizes
 * come from the context, never from the R object: `generation` proves the object and the context still belong together. */
SEXP C_tp_recall(SEXP ctx, SEXP generation, SEXP max_pcs, SEXP min_clusters) {
    tp_ctx *c = ctx_of(ctx);
    check_generation(c, generation);
    int nf = 0, kfull = 0, k = 0, npcs = 0, ncl = 0, maxlev = 0, ld = 256;
    if (tp_ctx_dims(c, NULL, &nf, NULL, &kfull, NULL) != TP_OK) error("%s", tp_last_error());
    if (nf < 3 || kfull < 1) error("TADpole: the context holds no PC scores");
    int kmax = asInteger(max_pcs) < nf ? asInteger(max_pcs) : nf;
    if (kmax < 1) error("TADpole: max_pcs must be positive");
    SEXP seq = PROTECT(allocVector(REALSXP, nf - 1));
    double *sc = (double *)R_alloc((size_t)kmax * ld, sizeof(double));
    int rc = tp_recall(c, asInteger(max_pcs), asInteger(min_clusters), &k, &npcs, &ncl, sc, ld, &maxlev, REAL(seq));
    SEXP out = pack_call(c, rc, k, npcs, ncl, maxlev, ld, sc, kmax, seq);
    UNPROTECT(1);
    return out;
}

/* seqdist of the candidate that clusters on the first n_pcs PCs (1-based), for CH_map / plot_hierarchy browsing */
SEXP C_tp_dendro(SEXP ctx, SEXP generation, SEXP n_pcs) {
    tp_ctx *c = ctx_of(ctx);
    check_generation(c, generation);
    int nf = 0, k = 0, maxlev = 0;
    if (tp_ctx_dims(c, NULL, &nf, &k, NULL, &maxlev) != TP_OK) error("%s", tp_last_error());
    if (nf < 3 || maxlev < 1) error("TADpole: the context holds no sweep");
    if (asInteger(n_pcs) < 1 || asInteger(n_pcs) > k) error("TADpole: n_pcs must be between 1 and %d", k);
    SEXP seq = PROTECT(allocVector(REALSXP, nf - 1));
    if (tp_get_dendro(c, asInteger(n_pcs) - 1, REAL(seq), NULL) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    UNPROTECT(1);
    return seq;
}

/* mat[keep, keep] after NA -> 0 and forceSymmetric(uplo = 'U'): the matrix load_mat() returns (R/TADpole.R:85,88-90);
 * keep = 0-based original indices.  Symmetric, so the row-major device copy is also the column-major R matrix. */
SEXP C_tp_get_filtered(SEXP ctx, SEXP keep) {
    tp_ctx *c = ctx_of(ctx);
    int nf = length(keep);
    if (nf < 2) error("TADpole: fewer than 2 good bins left after filtering");
    if (tp_compact(c, INTEGER(keep), nf) != TP_OK) error("%s", tp_last_error());
    SEXP m = PROTECT(allocMatrix(REALSXP, nf, nf));
    if (tp_get_filtered(c, REAL(m)) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    UNPROTECT(1);
    return m;
}

/* dendro$merge of the chclust object: rioja's .find.groups on seqdist, in C (O(n log n)) */
SEXP C_tp_find_groups(SEXP seqdist) {
    int n1 = length(seqdist);
    SEXP m = PROTECT(allocMatrix(INTSXP, n1, 2));
    if (n1 > 0 && tp_find_groups(REAL(seqdist), n1, INTEGER(m)) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    UNPROTECT(1);
    return m;
}

/* both arms of a centromere_search call at once (R/TADpole.R:357-374); on a multi-device context the two halves of the
 * devices work on the two arms at the same time.  list(p = <as C_tp_call_arm>, q = ...) */
SEXP C_tp_call_arms(SEXP ctx, SEXP keep_p, SEXP keep_q, SEXP max_pcs, SEXP min_clusters) {
    tp_ctx *c = ctx_of(ctx);
    SEXP keep[2] = {keep_p, keep_q};
    int nf[2] = {length(keep_p), length(keep_q)}, kmax[2], k[2], npcs[2], ncl[2], maxlev[2] = {0, 0}, ld = 256, rc;
    if (nf[0] < 3 || nf[1] < 3) error("TADpole: fewer than 3 good bins left in an arm after filtering");
    SEXP seq0 = PROTECT(allocVector(REALSXP, nf[0] - 1)), seq1 = PROTECT(allocVector(REALSXP, nf[1] - 1));
    SEXP seq[2] = {seq0, seq1};
    double *sc[2];
    for (;;) {
        for (int a = 0; a < 2; a++) {
            kmax[a] = asInteger(max_pcs) < nf[a] ? asInteger(max_pcs) : nf[a];
            sc[a] = (double *)R_alloc((size_t)kmax[a] * ld, sizeof(double));
        }
        rc = tp_call_arms(c, INTEGER(keep[0]), nf[0], INTEGER(keep[1]), nf[1], asInteger(max_pcs), asInteger(min_clusters),
                          k, npcs, ncl, sc[0], sc[1], ld, maxlev, REAL(seq[0]), REAL(seq[1]));
        int need = maxlev[0] > maxlev[1] ? maxlev[0] : maxlev[1];
        if (rc == TP_ERR_ARG && need > ld) { ld = need; continue; }       /* rare: more levels than the buffers hold */
        break;
    }
    if (rc != TP_OK) {
        UNPROTECT(2);
        error("%s", tp_last_error());
    }
    SEXP out = PROTECT(allocVector(VECSXP, 2));
    for (int a = 0; a < 2; a++)
        SET_VECTOR_ELT(out, a, pack_call(c, TP_OK, k[a], npcs[a], ncl[a], maxlev[a], ld, sc[a], kmax[a], seq[a]));
    UNPROTECT(3);
    return out;
}

/* TADpole() on a list of matrices, `inflight` calls per device kept in flight by the library's own threads over every
 * device of the context (tp_call_batch); this thread only waits.  Per matrix: list(bad, n_pcs, optimal_n_clusters,
 * seqdist, scores, levels, tables = list of 2-column (start, end) matrices), or a character error message. */
SEXP C_tp_call_batch(SEXP ctx, SEXP mats, SEXP max_pcs, SEXP min_clusters, SEXP bad_frac, SEXP inflight) {
    tp_ctx *c = ctx_of(ctx);
    int nc = length(mats);
    const double **ptr = (const double **)R_alloc((size_t)(nc ? nc : 1), sizeof(double *));
    int *n = (int *)R_alloc((size_t)(nc ? nc : 1), sizeof(int));
    for (int i = 0; i < nc; i++) {
        SEXP m = VECTOR_ELT(mats, i);
        if (!isReal(m) || nrows(m) != ncols(m)) error("TADpole: element %d is not a square numeric matrix", i + 1);
        ptr[i] = REAL(m);
        n[i] = nrows(m);
    }
    tp_batch *b = NULL;
    if (tp_call_batch(c, nc, ptr, n, 1, 0, asInteger(max_pcs), asInteger(min_clusters), asReal(bad_frac), asInteger(inflight),
                      1, &b) != TP_OK)
        error("%s", tp_last_error());
    /* (an R allocation failure below longjmps past tp_batch_free: the batch, plain host memory, is then leaked) */
    SEXP out = PROTECT(allocVector(VECSXP, nc));
    for (int i = 0; i < nc; i++) {
        if (tp_batch_status(b, i) != TP_OK) {
            SEXP msg = PROTECT(allocVector(STRSXP, 1));
            SET_STRING_ELT(msg, 0, mkChar(tp_batch_error(b, i)));
            SET_VECTOR_ELT(out, i, msg);
            UNPROTECT(1);
            continue;
        }
        int nn, nf, k, maxlev, nlev, nrow;
        tp_batch_dims(b, i, &nn, &nf, &k, &maxlev, &nlev, &nrow);
        SEXP bad = PROTECT(allocVector(LGLSXP, nn)), seq = PROTECT(allocVector(REALSXP, nf - 1));
        SEXP scores = PROTECT(allocMatrix(REALSXP, k, maxlev)), levels = PROTECT(allocVector(INTSXP, nlev));
        SEXP tables = PROTECT(allocVector(VECSXP, nlev));
        unsigned char *tb = (unsigned char *)R_alloc((size_t)nn, 1);
        double *sc = (double *)R_alloc((size_t)k * maxlev + 1, sizeof(double));
        int *off = (int *)R_alloc((size_t)nlev + 1, sizeof(int)), *st = (int *)R_alloc((size_t)nrow + 1, sizeof(int)),
            *en = (int *)R_alloc((size_t)nrow + 1, sizeof(int));
        int npcs = 0, ncl = 0;
        tp_batch_get(b, i, tb, &npcs, &ncl, sc, REAL(seq), INTEGER(levels), off, st, en, NULL);
        for (int j = 0; j < nn; j++) LOGICAL(bad)[j] = tb[j];
        for (int r = 0; r < k; r++)
            for (int l = 0; l < maxlev; l++) {
                double v = sc[(size_t)r * maxlev + l];
                REAL(scores)[r + (size_t)l * k] = ISNAN(v) ? NA_REAL : v;
            }
        for (int l = 0; l < nlev; l++) {
            int rows = off[l + 1] - off[l];
            SEXP m = PROTECT(allocMatrix(INTSXP, rows, 2));
            for (int r = 0; r < rows; r++) { INTEGER(m)[r] = st[off[l] + r]; INTEGER(m)[r + rows] = en[off[l] + r]; }
            SET_VECTOR_ELT(tables, l, m);
            UNPROTECT(1);
        }
        SEXP item = PROTECT(allocVector(VECSXP, 7));
        SET_VECTOR_ELT(item, 0, bad);
        SET_VECTOR_ELT(item, 1, ScalarInteger(npcs));
        SET_VECTOR_ELT(item, 2, ScalarInteger(ncl));
        SET_VECTOR_ELT(item, 3, seq);
        SET_VECTOR_ELT(item, 4, scores);
        SET_VECTOR_ELT(item, 5, levels);
        SET_VECTOR_ELT(item, 6, tables);
        SET_VECTOR_ELT(out, i, item);
        UNPROTECT(6);
    }
    tp_batch_free(b);
    UNPROTECT(1);
    return out;
}

/* the cutree / fix_values / rle loop of R/TADpole.R:470-497 for many levels at once: list of 2-column integer
 * matrices (start, end), one per requested level.  bad = integer(0) with nbad_null = TRUE means attr NULL. */
SEXP C_tp_levels(SEXP seqdist, SEXP levels, SEXP names, SEXP bad, SEXP bad_is_null) {
    int nf = length(names), nlev = length(levels);
    int nbad = asLogical(bad_is_null) ? -1 : length(bad);
    size_t cap = 0;
    for (int i = 0; i < nlev; i++) cap += (size_t)INTEGER(levels)[i] + (nbad > 0 ? nbad : 0) + 2;
    int *st = (int *)R_alloc(cap ? cap : 1, sizeof(int)), *en = (int *)R_alloc(cap ? cap : 1, sizeof(int));
    int *off = (int *)R_alloc((size_t)nlev + 1, sizeof(int));
    if (tp_assemble_levels(REAL(seqdist), nf, INTEGER(levels), nlev, INTEGER(names), INTEGER(bad), nbad, st, en, off) != TP_OK)
        error("%s", tp_last_error());
    SEXP out = PROTECT(allocVector(VECSXP, nlev));
    for (int i = 0; i < nlev; i++) {
        int rows = off[i + 1] - off[i];
        SEXP m = PROTECT(allocMatrix(INTSXP, rows, 2));
        for (int r = 0; r < rows; r++) {
            INTEGER(m)[r] = st[off[i] + r];
            INTEGER(m)[r + rows] = en[off[i] + r];
        }
        SET_VECTOR_ELT(out, i, m);
        UNPROTECT(1);
    }
    UNPROTECT(1);
    return out;
}

/* one level: cutree + bad bins re-inserted as 0 + fix_values (R/TADpole.R:411-431): the per-bin label vector */
SEXP C_tp_labels(SEXP seqdist, SEXP n_clusters, SEXP names, SEXP bad, SEXP bad_is_null) {
    int nf = length(names), k = asInteger(n_clusters), nrows_out = 0;
    int nbad = asLogical(bad_is_null) ? -1 : length(bad);
    size_t cap = (size_t)k + (nbad > 0 ? nbad : 0) + 2;
    int *st = (int *)R_alloc(cap, sizeof(int)), *en = (int *)R_alloc(cap, sizeof(int));
    SEXP lab = PROTECT(allocVector(INTSXP, nf + (nbad > 0 ? nbad : 0)));
    if (tp_assemble(REAL(seqdist), nf, k, INTEGER(names), INTEGER(bad), nbad, st, en, &nrows_out, INTEGER(lab)) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    UNPROTECT(1);
    return lab;
}

/* the loop of diffT (R/DiffT.R:41-49) for npairs padded label vectors of length L (row-major npairs x L) */
SEXP C_tp_difft(SEXP ctx, SEXP lx, SEXP ly, SEXP L, SEXP npairs) {
    size_t n = (size_t)asInteger(L) * asInteger(npairs);
    if ((size_t)length(lx) != n || (size_t)length(ly) != n) error("TADpole: label vectors do not match L x npairs");
    SEXP out = PROTECT(allocVector(REALSXP, n));
    if (tp_difft_batch(ctx_of(ctx), INTEGER(lx), INTEGER(ly), asInteger(L), asInteger(npairs), 0, REAL(out)) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    UNPROTECT(1);
    return out;
}

/* nperm draws of random_bed (R/DiffT.R:61-73) on the device, each scored against labels_x:
 * list(borders [ (ntads-1) x nperm ], totals [nperm], curves [L x nperm]) -- column per permutation */
SEXP C_tp_difft_null(SEXP ctx, SEXP lx, SEXP pad_left, SEXP pad_right, SEXP ntads, SEXP bad, SEXP seed, SEXP nperm) {
    int L = length(lx), T = asInteger(ntads), P = asInteger(nperm);
    SEXP borders = PROTECT(allocMatrix(INTSXP, T > 1 ? T - 1 : 0, P));
    SEXP totals = PROTECT(allocVector(REALSXP, P));
    SEXP curves = PROTECT(allocMatrix(REALSXP, L, P));
    int rc = tp_difft_null(ctx_of(ctx), INTEGER(lx), L, asInteger(pad_left), asInteger(pad_right), T,
                           length(bad) ? INTEGER(bad) : NULL, length(bad), (unsigned long long)asReal(seed), P,
                           T > 1 ? INTEGER(borders) : NULL, NULL, REAL(curves), REAL(totals));
    if (rc != TP_OK) {
        UNPROTECT(3);
        error("%s", tp_last_error());
    }
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, borders);
    SET_VECTOR_ELT(out, 1, totals);
    SET_VECTOR_ELT(out, 2, curves);
    UNPROTECT(4);
    return out;
}

/* several GPUs: one R process per GPU; rank 0 makes the id, every rank joins (csrc/comm.cu) */
SEXP C_tp_comm_unique_id(void) {
    SEXP id = PROTECT(allocVector(RAWSXP, 128));
    if (tp_comm_unique_id(RAW(id)) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    UNPROTECT(1);
    return id;
}

SEXP C_tp_ctx_comm_init(SEXP ctx, SEXP id, SEXP rank, SEXP nranks, SEXP slot) {
    if (length(id) != 128) error("TADpole: the communicator id must be 128 raw bytes");
    if (tp_ctx_comm_init(ctx_of(ctx), RAW(id), asInteger(rank), asInteger(nranks), asInteger(slot)) != TP_OK)
        error("%s", tp_last_error());
    return R_NilValue;
}

SEXP C_tp_ctx_comm_select(SEXP ctx, SEXP slot) {
    if (tp_ctx_comm_select(ctx_of(ctx), asInteger(slot)) != TP_OK) error("%s", tp_last_error());
    return R_NilValue;
}

static const R_CallMethodDef call_table[] = {
    {"C_tp_ctx", (DL_FUNC)&C_tp_ctx, 1},
    {"C_tp_ingest", (DL_FUNC)&C_tp_ingest, 2},
    {"C_tp_ingested_matrix", (DL_FUNC)&C_tp_ingested_matrix, 1},
    {"C_tp_ingest_coo", (DL_FUNC)&C_tp_ingest_coo, 6},
    {"C_tp_ingest_coo_file", (DL_FUNC)&C_tp_ingest_coo_file, 4},
    {"C_tp_filter", (DL_FUNC)&C_tp_filter, 3},
    {"C_tp_call_arm", (DL_FUNC)&C_tp_call_arm, 4},
    {"C_tp_recall", (DL_FUNC)&C_tp_recall, 4},
    {"C_tp_dendro", (DL_FUNC)&C_tp_dendro, 3},
    {"C_tp_generation", (DL_FUNC)&C_tp_generation, 1},
    {"C_tp_get_filtered", (DL_FUNC)&C_tp_get_filtered, 2},
    {"C_tp_find_groups", (DL_FUNC)&C_tp_find_groups, 1},
    {"C_tp_call_arms", (DL_FUNC)&C_tp_call_arms, 5},
    {"C_tp_call_batch", (DL_FUNC)&C_tp_call_batch, 6},
    {"C_tp_levels", (DL_FUNC)&C_tp_levels, 5},
    {"C_tp_labels", (DL_FUNC)&C_tp_labels, 5},
    {"C_tp_difft", (DL_FUNC)&C_tp_difft, 5},
    {"C_tp_difft_null", (DL_FUNC)&C_tp_difft_null, 8},
    {"C_tp_comm_unique_id", (DL_FUNC)&C_tp_comm_unique_id, 0},
    {"C_tp_ctx_comm_init", (DL_FUNC)&C_tp_ctx_comm_init, 5},
    {"C_tp_ctx_comm_select", (DL_FUNC)&C_tp_ctx_comm_select, 2},
    {NULL, NULL, 0}};

void R_init_TADpoleB200(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_table, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
