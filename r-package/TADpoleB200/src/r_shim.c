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

SEXP C_tp_ctx(SEXP device) {
    tp_ctx *c = NULL;
    if (tp_ctx_create(asInteger(device), &c) != TP_OK) error("%s", tp_last_error());
    SEXP p = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* read.big.matrix(mat_file, type = 'double', sep = '\t') on the device (R/TADpole.R:17): returns N */
SEXP C_tp_ingest(SEXP ctx, SEXP path) {
    int n = 0;
    if (tp_ingest_tsv_file(ctx_of(ctx), CHAR(STRING_ELT(path, 0)), '\t', &n) != TP_OK) error("%s", tp_last_error());
    return ScalarInteger(n);
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
    SEXP out = PROTECT(allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, ScalarInteger(npcs));
    SET_VECTOR_ELT(out, 1, ScalarInteger(ncl));
    SET_VECTOR_ELT(out, 2, seq);
    SET_VECTOR_ELT(out, 3, scores);
    UNPROTECT(2);
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

/* the sweep again on the resident PC scores with another max_pcs / min_clusters (no reference counterpart) */
SEXP C_tp_recall(SEXP ctx, SEXP nf_, SEXP max_pcs, SEXP min_clusters) {
    tp_ctx *c = ctx_of(ctx);
    int nf = asInteger(nf_), k = 0, npcs = 0, ncl = 0, maxlev = 0, ld = 256;
    int kmax = asInteger(max_pcs) < nf ? asInteger(max_pcs) : nf;
    SEXP seq = PROTECT(allocVector(REALSXP, nf - 1));
    double *sc = (double *)R_alloc((size_t)kmax * ld, sizeof(double));
    int rc = tp_recall(c, asInteger(max_pcs), asInteger(min_clusters), &k, &npcs, &ncl, sc, ld, &maxlev, REAL(seq));
    SEXP out = pack_call(c, rc, k, npcs, ncl, maxlev, ld, sc, kmax, seq);
    UNPROTECT(1);
    return out;
}

/* seqdist of the candidate that clusters on the first n_pcs PCs (1-based), for CH_map / plot_hierarchy browsing */
SEXP C_tp_dendro(SEXP ctx, SEXP nf_, SEXP n_pcs) {
    int nf = asInteger(nf_);
    SEXP seq = PROTECT(allocVector(REALSXP, nf - 1));
    if (tp_get_dendro(ctx_of(ctx), asInteger(n_pcs) - 1, REAL(seq), NULL) != TP_OK) {
        UNPROTECT(1);
        error("%s", tp_last_error());
    }
    UNPROTECT(1);
    return seq;
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
    {"C_tp_filter", (DL_FUNC)&C_tp_filter, 3},
    {"C_tp_call_arm", (DL_FUNC)&C_tp_call_arm, 4},
    {"C_tp_recall", (DL_FUNC)&C_tp_recall, 4},
    {"C_tp_dendro", (DL_FUNC)&C_tp_dendro, 3},
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
