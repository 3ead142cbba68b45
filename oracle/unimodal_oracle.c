/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's unimodal regression
 *   /root/reference/src/matcouply/_unimodal_regression.py
 *     prefix_isotonic_regression      :24-69   (PAVA with `<=` pooling, running-sum errors)
 *     _compute_isotonic_from_index    :72-81
 *     _get_best_unimodality_index     :84-92   (first strict minimum)
 *     _unimodal_regression            :95-104
 *     unimodal_regression             :107-141 (column-wise over a row-major matrix)
 *
 * The floating-point operation ORDER follows the reference line by line so that results are
 * bit-identical to the NumPy/numba reference (build with -ffp-contract=off, no -ffast-math).
 * Pinned by tests/test_oracle.py against golden vectors produced from the reference itself
 * (oracle/gen_golden.py) and against scikit-learn's IsotonicRegression.
 *
 * Build: see oracle/Makefile  ->  oracle/_build/liboracle.so
 */
#include <stdlib.h>
#include <string.h>

/* Prefix isotonic regression of y[0..n) read with stride `stride` (so the reversed vector is stride<0).
 * level[n], start[n] (block start index of the block ending at i), err[n+1]. Weights are all 1
 * (the reference is only ever called with weights=None -> ones). scratch: 3*n doubles (+n when nn). */
static void prefix_isotonic(const double *y, long stride, int n, int non_negativity, double *level, int *start,
                            double *err, double *scratch)
{
    double *sumwy = scratch, *sumwy2 = scratch + n, *sumw = scratch + 2 * n;
    double *cumsumwy2 = scratch + 3 * n;
    unsigned char *thr = NULL;
    int i;
    for (i = 0; i < n; ++i) {
        double yi = y[(long)i * stride];
        sumwy[i] = 1.0 * yi;       /* weights * y           (:31) */
        sumwy2[i] = 1.0 * yi * yi; /* weights * y * y       (:32) */
        sumw[i] = 1.0;             /* weights.copy()        (:33) */
        level[i] = 0.0;
        start[i] = 0;
    }
    for (i = 0; i <= n; ++i) err[i] = 0.0;
    level[0] = y[0];
    start[0] = 0;
    if (non_negativity) {
        double acc = 0.0; /* np.cumsum is a sequential running sum (:44) */
        thr = (unsigned char *)calloc((size_t)n, 1);
        for (i = 0; i < n; ++i) {
            acc += sumwy2[i];
            cumsumwy2[i] = acc;
        }
        if (level[0] < 0) { /* (:46-48) */
            thr[0] = 1;
            err[1] = cumsumwy2[0];
        }
    }
    for (i = 1; i < n; ++i) {
        double levelerror;
        level[i] = y[(long)i * stride];
        start[i] = i;
        /* pool while the new block's level is <= the previous block's level (:53-55).
         * NB: the reference evaluates level_set[index_range[i]-1] before testing index_range[i] != 0;
         * with index 0 NumPy wraps to level_set[-1]; the conjunction makes the value irrelevant. */
        while (start[i] != 0 && level[i] <= level[start[i] - 1]) {
            int prev = start[i] - 1;
            sumwy[i] += sumwy[prev];
            sumwy2[i] += sumwy2[prev];
            sumw[i] += sumw[prev];
            level[i] = sumwy[i] / sumw[i];
            start[i] = start[prev];
        }
        levelerror = sumwy2[i] - (sumwy[i] * sumwy[i] / sumw[i]); /* (:57)  (a**2)/b */
        if (non_negativity && level[i] < 0) {
            thr[i] = 1;
            err[i + 1] = cumsumwy2[i]; /* (:58-60) */
        } else {
            err[i + 1] = levelerror + err[start[i]]; /* (:62) */
        }
    }
    if (non_negativity) {
        for (i = 0; i < n; ++i)
            if (thr[i]) level[i] = 0.0; /* (:64-67) */
        free(thr);
    }
}

/* Unimodal regression of one strided column. out has stride out_stride. Returns the peak index t*. */
static int unimodal_column(const double *y, long stride, int n, int non_negativity, double *out, long out_stride,
                           double *best_error_out)
{
    double *levelL = (double *)malloc(sizeof(double) * (size_t)n * 2);
    double *levelR = levelL + n;
    int *startL = (int *)malloc(sizeof(int) * (size_t)n * 2);
    int *startR = startL + n;
    double *errL = (double *)malloc(sizeof(double) * (size_t)(n + 1) * 2);
    double *errR = errL + n + 1;
    double *scratch = (double *)malloc(sizeof(double) * (size_t)n * 4);
    double best_error;
    int best_idx = 0, i, idx;

    prefix_isotonic(y, stride, n, non_negativity, levelL, startL, errL, scratch);
    /* reversed vector y[::-1] (:97) */
    prefix_isotonic(y + (long)(n - 1) * stride, -stride, n, non_negativity, levelR, startR, errR, scratch);

    /* (:84-92) error_left has n+1 entries; first strict minimum */
    best_error = errR[n];
    for (i = 0; i < n + 1; ++i) {
        double e = errL[i] + errR[(n + 1) - i - 1];
        if (e < best_error) {
            best_error = e;
            best_idx = i;
        }
    }
    /* left part: isotonic fit of y[:best_idx] (:72-81) */
    idx = best_idx - 1;
    while (idx >= 0) {
        int s = startL[idx], j;
        for (j = s; j <= idx; ++j) out[(long)j * out_stride] = levelL[idx];
        idx = s - 1;
    }
    /* right part: isotonic fit of reversed y[: n-best_idx], then reversed back (:100-104) */
    idx = (n - best_idx) - 1;
    while (idx >= 0) {
        int s = startR[idx], j;
        for (j = s; j <= idx; ++j) out[(long)(n - 1 - j) * out_stride] = levelR[idx];
        idx = s - 1;
    }
    if (best_error_out) *best_error_out = best_error;
    free(levelL);
    free(startL);
    free(errL);
    free(scratch);
    return best_idx;
}

/* Column-wise unimodal regression of a row-major (n_rows x n_cols) matrix.
 * peaks (may be NULL): n_cols ints receiving t*; errors (may be NULL): n_cols doubles. */
void oracle_unimodal_regression(const double *y, int n_rows, int n_cols, int non_negativity, double *out, int *peaks,
                                double *errors)
{
    int r;
    for (r = 0; r < n_cols; ++r) {
        double e;
        int t = unimodal_column(y + r, n_cols, n_rows, non_negativity, out + r, n_cols, &e);
        if (peaks) peaks[r] = t;
        if (errors) errors[r] = e;
    }
}

/* Exposed for tests: prefix isotonic regression of a contiguous vector. */
void oracle_prefix_isotonic(const double *y, int n, int non_negativity, double *level, int *start, double *err)
{
    double *scratch = (double *)malloc(sizeof(double) * (size_t)n * 4);
    prefix_isotonic(y, 1, n, non_negativity, level, start, err, scratch);
    free(scratch);
}
