/*
 * lsap_oracle.c -- CPU restatement of the rectangular linear-sum-assignment solver the reference's
 * matcher calls.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as rdetr_oracle.c): the product path never links this file.
 *
 * Where the algorithm lives: NOT in the reference tree.  models/matcher/hungarian_matcher.py:80,87
 * call scipy.optimize.linear_sum_assignment on the cost matrix copied to the host.  SciPy is a
 * third-party dependency (unpinned in the reference's requirements; 1.18.1 in this image).  Its solver
 * is the shortest-augmenting-path algorithm of D. F. Crouse, "On implementing 2D rectangular
 * assignment algorithms", IEEE Trans. Aerospace and Electronic Systems 52(4), 2016, in double
 * precision, with three documented conventions that fix WHICH optimum is returned when several exist:
 *   (1) the matrix is transposed when it has more rows than columns, and the result is re-sorted by row;
 *   (2) the set of unscanned columns is kept as a vector filled in REVERSE order (so that a constant
 *       matrix yields the identity) and a scanned column is removed by moving the last entry into its
 *       slot;
 *   (3) among columns of equal reduced cost an unassigned column is preferred, the LAST such column in
 *       vector order; without an unassigned candidate the FIRST column of minimal cost wins.
 * This file restates that algorithm from the paper and those conventions.  Parity status: PINNED against
 * scipy.optimize.linear_sum_assignment itself (tests/test_lsap_oracle.py: random, integer/tie-heavy,
 * duplicated-column, both orientations, empty, infeasible) and against index fixtures produced by the
 * reference's HungarianMatcher (tests/golden/matcher_*.npz, oracle/make_golden_matcher.py).
 *
 * Return codes: 0 ok, -1 infeasible (a row with no finite entry left), -2 invalid entry (NaN or -inf),
 * matching SciPy's two ValueErrors.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* One augmenting-path search from row `start` over the nr x nc matrix c (nr <= nc). */
static int64_t lsap_search(int64_t nc, const double *c, const double *u, const double *v, int64_t *path,
                           const int64_t *row4col, double *dist, int64_t start, unsigned char *SR,
                           unsigned char *SC, int64_t *todo, int64_t nr, double *min_out)
{
    double min_val = 0.0;
    int64_t n_todo = nc, sink = -1, i = start;
    for (int64_t t = 0; t < nc; ++t) { todo[t] = nc - t - 1; SC[t] = 0; dist[t] = INFINITY; }
    for (int64_t r = 0; r < nr; ++r) SR[r] = 0;

    while (sink < 0) {
        int64_t pick = -1;
        double lowest = INFINITY;
        SR[i] = 1;
        for (int64_t t = 0; t < n_todo; ++t) {
            const int64_t j = todo[t];
            const double r = min_val + c[i * nc + j] - u[i] - v[j];
            if (r < dist[j]) { path[j] = i; dist[j] = r; }
            if (dist[j] < lowest || (dist[j] == lowest && row4col[j] < 0)) { lowest = dist[j]; pick = t; }
        }
        min_val = lowest;
        if (min_val == INFINITY) return -1;
        const int64_t j = todo[pick];
        if (row4col[j] < 0) sink = j; else i = row4col[j];
        SC[j] = 1;
        todo[pick] = todo[--n_todo];
    }
    *min_out = min_val;
    return sink;
}

/* cost: n_rows x n_cols row-major doubles.  Writes min(n_rows, n_cols) pairs (row ascending). */
int rdetr_oracle_lsap(const double *cost, int64_t n_rows, int64_t n_cols, int64_t *row_ind, int64_t *col_ind)
{
    if (n_rows <= 0 || n_cols <= 0) return 0;
    const int transposed = n_cols < n_rows;
    const int64_t nr = transposed ? n_cols : n_rows, nc = transposed ? n_rows : n_cols;

    double *c = (double *)malloc(sizeof(double) * (size_t)(nr * nc));
    for (int64_t i = 0; i < nr; ++i)
        for (int64_t j = 0; j < nc; ++j) {
            const double x = transposed ? cost[j * n_cols + i] : cost[i * n_cols + j];
            if (x != x || x == -INFINITY) { free(c); return -2; }
            c[i * nc + j] = x;
        }

    double *u = (double *)calloc((size_t)nr, sizeof(double)), *v = (double *)calloc((size_t)nc, sizeof(double));
    double *dist = (double *)malloc(sizeof(double) * (size_t)nc);
    int64_t *path = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc), *todo = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
    int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * (size_t)nr), *row4col = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
    unsigned char *SR = (unsigned char *)malloc((size_t)nr), *SC = (unsigned char *)malloc((size_t)nc);
    for (int64_t i = 0; i < nr; ++i) col4row[i] = -1;
    for (int64_t j = 0; j < nc; ++j) { row4col[j] = -1; path[j] = -1; }

    int rc = 0;
    for (int64_t cur = 0; cur < nr && rc == 0; ++cur) {
        double min_val = 0.0;
        int64_t j = lsap_search(nc, c, u, v, path, row4col, dist, cur, SR, SC, todo, nr, &min_val);
        if (j < 0) { rc = -1; break; }
        u[cur] += min_val;                                       /* dual update */
        for (int64_t i = 0; i < nr; ++i)
            if (SR[i] && i != cur) u[i] += min_val - dist[col4row[i]];
        for (int64_t k = 0; k < nc; ++k)
            if (SC[k]) v[k] -= min_val - dist[k];
        for (;;) {                                               /* flip the path */
            const int64_t i = path[j];
            row4col[j] = i;
            const int64_t prev = col4row[i];
            col4row[i] = j;
            j = prev;
            if (i == cur) break;
        }
    }

    if (rc == 0) {
        if (!transposed) {
            for (int64_t i = 0; i < nr; ++i) { row_ind[i] = i; col_ind[i] = col4row[i]; }
        } else {
            /* pairs (col4row[g], g) sorted by their first member: every original row is used at most once */
            int64_t k = 0;
            for (int64_t q = 0; q < nc; ++q)
                if (row4col[q] >= 0) { row_ind[k] = q; col_ind[k] = row4col[q]; ++k; }
        }
    }
    free(c); free(u); free(v); free(dist); free(path); free(todo); free(col4row); free(row4col); free(SR); free(SC);
    return rc;
}
