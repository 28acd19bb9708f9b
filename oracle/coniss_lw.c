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
