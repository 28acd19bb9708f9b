/* CPU oracle, plain C -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see tadpole_oracle.py header).
 *
 * oracle_coniss_lw restates what the reference does per candidate at R/TADpole.R:108
 *     clust <- rioja::chclust(dist(pcs))
 * in the reference's own algorithmic shape: stats::dist builds every pairwise Euclidean
 * distance (O(n^2 p)); rioja::chclust(method = "coniss") (rioja >= 0.9-21, source not under
 * /root/reference; algorithm after Grimm 1987) squares them and, n-1
 * times, scans ADJACENT cluster pairs for the smallest dispersion increase (strict '<', so
 * the lowest index wins ties), adds it to a running total, records that total for the
 * boundary it removed (seqdist) and updates squared dissimilarities to every other live
 * cluster with the Ward / Lance-Williams recurrence
 *     d(r, p+q) = [(n_r+n_p) d(r,p) + (n_r+n_q) d(r,q) - n_r d(p,q)] / (n_r+n_p+n_q),
 * the increase of a pair being d/2.  Parity with rioja is UNPINNED (no R here).
 *
 * oracle_difft restates the O(L^2) loop of R/DiffT.R:41-49.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* x: n x p row-major with leading dimension ld.  seqdist[n-1], order[n-1] (boundary removed
 * at each step, 0-based: boundary j separates objects j and j+1). */
int oracle_coniss_lw(const double *x, int n, int ld, int p, double *seqdist, int *order) {
    if (n < 2) return 0;
    size_t nn = (size_t)n;
    double *d = (double *)malloc(nn * nn * sizeof(double));
    int *cnt = (int *)malloc(nn * sizeof(int));
    int *nxt = (int *)malloc(nn * sizeof(int));
    int *prv = (int *)malloc(nn * sizeof(int));
    int *last = (int *)malloc(nn * sizeof(int));
    if (!d || !cnt || !nxt || !prv || !last) { free(d); free(cnt); free(nxt); free(prv); free(last); return 1; }
    /* stats::dist: sqrt(sum (xi-xj)^2); chclust squares it again */
    for (size_t i = 0; i < nn; i++) {
        d[i * nn + i] = 0.0;
        for (size_t j = i + 1; j < nn; j++) {
            double s = 0.0;
            const double *a = x + i * (size_t)ld, *b = x + j * (size_t)ld;
            for (int c = 0; c < p; c++) { double t = a[c] - b[c]; s += t * t; }
            double e = sqrt(s);
            d[i * nn + j] = d[j * nn + i] = e * e;
        }
    }
    for (int i = 0; i < n; i++) { cnt[i] = 1; nxt[i] = i + 1; prv[i] = i - 1; last[i] = i; }
    double total = 0.0;
    for (int step = 0; step < n - 1; step++) {
        int best = -1; double bestv = 0.0;
        for (int a = 0; nxt[a] < n; a = nxt[a]) {
            double v = 0.5 * d[(size_t)a * nn + nxt[a]];
            if (best < 0 || v < bestv) { best = a; bestv = v; }
        }
        int pa = best, q = nxt[pa];
        total += bestv;
        seqdist[last[pa]] = total;
        order[step] = last[pa];
        double dpq = d[(size_t)pa * nn + q];
        int np_ = cnt[pa], nq = cnt[q];
        for (int r = 0; r < n; r = nxt[r]) {
            if (r == pa || r == q) { continue; }
            int nr = cnt[r];
            double v = ((double)(nr + np_) * d[(size_t)r * nn + pa] + (double)(nr + nq) * d[(size_t)r * nn + q]
                        - (double)nr * dpq) / (double)(nr + np_ + nq);
            d[(size_t)r * nn + pa] = d[(size_t)pa * nn + r] = v;
        }
        cnt[pa] = np_ + nq;
        last[pa] = last[q];
        nxt[pa] = nxt[q];
        if (nxt[q] < n) prv[nxt[q]] = pa;
    }
    free(d); free(cnt); free(nxt); free(prv); free(last);
    return 0;
}

/* R/DiffT.R:41-49 on padded label vectors; out = cumulative score, normalised by its last
 * value unless every per-bin score is zero. */
int oracle_difft(const int *tx, const int *ty, int L, double *out) {
    long long cum = 0, mx = 0;
    for (int b = 0; b < L; b++) {
        long long s = 0;
        int bx = tx[b], by = ty[b];
        for (int j = 0; j < L; j++) {
            int x = (bx != tx[j]) | (bx == 0);
            int y = (by != ty[j]) | (by == 0);
            s += x ^ y;
        }
        if (s > mx) mx = s;
        cum += s;
        out[b] = (double)cum;
    }
    if (mx != 0) { double t = out[L - 1]; for (int b = 0; b < L; b++) out[b] /= t; }
    return 0;
}

/* ---- one candidate of the find_params sweep, start to end (R/TADpole.R:105-122) --------------------------------
 *     pcs <- pca$x[, 1:i];  clust <- rioja::chclust(dist(pcs))                      -> oracle_coniss_lw above
 *     bs <- rioja::bstick(clust, ng = nrow(mat) - 1);  n_cluster = first TRUE run of dispersion > bstick  (quirk Q1)
 *     for (n in min(min_clusters, n_cluster):n_cluster)
 *         score[n] <- fpc::calinhara(pca$x, cutree(clust, n), n)                   on ALL k columns (quirk Q2)
 * in plain C so that the timed CPU arm (bench.py --impl reference, cpu_baseline) runs every candidate in full on every
 * host core (ctypes releases the GIL).  calinhara is evaluated in trace form, two passes per level (cluster means, then
 * squared deviations): the reference's fpc forms k x k covariance matrices per cluster, O(n k^2) per level, so this port
 * UNDER-states the reference's cost; it never over-states it.
 * x: n x ld row-major PC scores, k_all columns in use.  seqdist[n-1]; scores[cap] (NaN padded).
 * returns 0, 1 = out of memory, 2 = no significant broken-stick level (the reference errors), 3 = cap too small
 * (*ncl_out tells the width needed). */
static const double *g_sort_key;
static int cmp_desc(const void *a, const void *b) {
    const int ia = *(const int *)a, ib = *(const int *)b;
    if (g_sort_key[ia] > g_sort_key[ib]) return -1;
    if (g_sort_key[ia] < g_sort_key[ib]) return 1;
    return ib - ia;                                   /* later boundary first: the merge order is (value, index) ascending */
}

int oracle_candidate(const double *x, int n, int ld, int k_all, int i, int min_clusters, double *seqdist, double *scores,
                     int cap, int *ncl_out) {
    const int n1 = n - 1;
    int *order = (int *)malloc((size_t)n1 * sizeof(int));
    double *bs = (double *)malloc((size_t)n1 * sizeof(double));
    double *height = (double *)malloc((size_t)n1 * sizeof(double));
    int *rank = (int *)malloc((size_t)n1 * sizeof(int));
    int *idx = (int *)malloc((size_t)n1 * sizeof(int));
    double *mean = (double *)malloc((size_t)n * k_all * sizeof(double));     /* at most n clusters */
    int *lab = (int *)malloc((size_t)n * sizeof(int)), *cnt = (int *)malloc((size_t)n * sizeof(int));
    int rc = 1;
    if (!order || !bs || !height || !rank || !idx || !mean || !lab || !cnt) goto done;
    if (oracle_coniss_lw(x, n, ld, i, seqdist, order)) goto done;
    /* height = sort(seqdist) = the running total in merge order */
    for (int t = 0; t < n1; t++) height[t] = seqdist[order[t]];
    {   /* vegan::bstick.default(n1, tot) = rev(cumsum(tot / n1:1) / n1) */
        const double tot = height[n1 - 1];
        double c = 0.0;
        for (int m = n1; m >= 1; m--) { c += tot / (double)m; bs[m - 1] = c / (double)n1; }
    }
    int first = -1, run = 0;
    for (int j = 1; j <= n1 - 1; j++) {               /* dispersion_j = |disp[j+1] - disp[j]|, disp = rev(height) */
        const int f = fabs(height[n1 - j - 1] - height[n1 - j]) > bs[j - 1];
        if (first < 0) { if (f) { first = j; run = 1; } }
        else if (f) run++;
        else break;
    }
    if (first < 0) { rc = 2; goto done; }
    *ncl_out = run;
    if (run > cap) { rc = 3; goto done; }
    for (int l = 0; l < cap; l++) scores[l] = NAN;
    /* cutree: the boundaries ranked by (seqdist descending, index descending); level m keeps the first m - 1 */
    for (int j = 0; j < n1; j++) idx[j] = j;
    g_sort_key = seqdist;
    qsort(idx, (size_t)n1, sizeof(int), cmp_desc);
    for (int r = 0; r < n1; r++) rank[idx[r]] = r;
    {
        double trs = 0.0;                              /* tr(S): total SS about the column means, all k columns */
        for (int c = 0; c < k_all; c++) {
            double m = 0.0;
            for (int r = 0; r < n; r++) m += x[(size_t)r * ld + c];
            m /= (double)n;
            for (int r = 0; r < n; r++) { const double t = x[(size_t)r * ld + c] - m; trs += t * t; }
        }
        const int mc = min_clusters < run ? min_clusters : run;
        for (int lev = mc; lev <= run; lev++) {
            int cl = 0;
            for (int r = 0; r < n; r++) { lab[r] = cl; if (r < n1 && rank[r] < lev - 1) cl++; }
            memset(mean, 0, (size_t)lev * k_all * sizeof(double));
            memset(cnt, 0, (size_t)lev * sizeof(int));
            for (int r = 0; r < n; r++) {
                double *m = mean + (size_t)lab[r] * k_all;
                const double *row = x + (size_t)r * ld;
                for (int c = 0; c < k_all; c++) m[c] += row[c];
                cnt[lab[r]]++;
            }
            for (int g = 0; g < lev; g++) for (int c = 0; c < k_all; c++) mean[(size_t)g * k_all + c] /= (double)cnt[g];
            double trw = 0.0;
            for (int r = 0; r < n; r++) {
                if (cnt[lab[r]] < 2) continue;         /* clusters of one object contribute 0 */
                const double *m = mean + (size_t)lab[r] * k_all, *row = x + (size_t)r * ld;
                for (int c = 0; c < k_all; c++) { const double t = row[c] - m[c]; trw += t * t; }
            }
            scores[lev - 1] = (double)(n - lev) * (trs - trw) / ((double)(lev - 1) * trw);
        }
    }
    rc = 0;
done:
    free(order); free(bs); free(height); free(rank); free(idx); free(mean); free(lab); free(cnt);
    return rc;
}
