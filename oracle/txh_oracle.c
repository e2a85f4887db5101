/*
 * txh_oracle.c -- CPU restatement of tx-fast-hydrology's routing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: a plain-C, scalar
 * restatement of the reference's numba kernels.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path (tx_fast_hydrology_b200/) never links, imports or calls anything here.
 *
 * Parity status: PINNED.  The reference ships no golden vectors of its own
 * (SURVEY.md section 4), so every function below is pinned against the reference
 * itself, executed live in the build container (tests/golden/make_golden.py imports
 * /root/reference and writes tests/golden/ (npz files); tests/test_oracle.py replays them).
 *
 * Each function cites the reference file:line it follows.  Paths are relative to
 * the reference repository root (tx_fast_hydrology/...).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TXO_API __attribute__((visibility("default")))

/* ---- nutils.py:64-89  _ax_bu ------------------------------------------------
 * One routing step with lateral forcing.  Walk downstream from every headwater
 * (ascending index) while all upstream reaches of the current one are done.
 * `work` is an int64[n] scratch (the reference's indegree.copy()). */
TXO_API void txo_ax_bu(int64_t n, int64_t h, const int64_t *startnodes,
                       const int64_t *endnodes, const double *alpha,
                       const double *beta, const double *chi, const double *gamma,
                       const double *i_t_prev, const double *o_t_prev,
                       const double *q_t_next, const int64_t *indegree,
                       double *i_t_next, double *o_t_next, int64_t *work)
{
    memset(i_t_next, 0, (size_t)n * sizeof(double));
    memset(o_t_next, 0, (size_t)n * sizeof(double));
    memcpy(work, indegree, (size_t)n * sizeof(int64_t));
    for (int64_t k = 0; k < h; ++k) {
        int64_t s = startnodes[k];
        int64_t e = endnodes[s];
        while (work[s] == 0) {
            /* nutils.py:79-82: ((a*i_next + b*i_prev) + c*o_prev) + g*q, no fma */
            double t0 = alpha[s] * i_t_next[s];
            double t1 = beta[s] * i_t_prev[s];
            double t2 = chi[s] * o_t_prev[s];
            double t3 = gamma[s] * q_t_next[s];
            o_t_next[s] += ((t0 + t1) + t2) + t3;
            if (s != e)
                i_t_next[e] += o_t_next[s];
            work[e] -= 1;
            s = e;
            e = endnodes[s];
        }
    }
}

/* ---- nutils.py:91-114  _ax  (homogeneous operator, no gamma*q term) ---------- */
TXO_API void txo_ax(int64_t n, int64_t h, const int64_t *startnodes,
                    const int64_t *endnodes, const double *alpha, const double *beta,
                    const double *chi, const double *i_t_prev, const double *o_t_prev,
                    const int64_t *indegree, double *i_t_next, double *o_t_next,
                    int64_t *work)
{
    memset(i_t_next, 0, (size_t)n * sizeof(double));
    memset(o_t_next, 0, (size_t)n * sizeof(double));
    memcpy(work, indegree, (size_t)n * sizeof(int64_t));
    for (int64_t k = 0; k < h; ++k) {
        int64_t s = startnodes[k];
        int64_t e = endnodes[s];
        while (work[s] == 0) {
            double t0 = alpha[s] * i_t_next[s];
            double t1 = beta[s] * i_t_prev[s];
            double t2 = chi[s] * o_t_prev[s];
            o_t_next[s] += (t0 + t1) + t2;
            if (s != e)
                i_t_next[e] += o_t_next[s];
            work[e] -= 1;
            s = e;
            e = endnodes[s];
        }
    }
}

/* ---- nutils.py:116-134  _apply_gain ------------------------------------------ */
TXO_API void txo_apply_gain(int64_t n, int64_t h, const int64_t *startnodes,
                            const int64_t *endnodes, const double *gain,
                            const int64_t *indegree, double *i_t_next,
                            double *o_t_next, int64_t *work)
{
    memset(i_t_next, 0, (size_t)n * sizeof(double));
    memset(o_t_next, 0, (size_t)n * sizeof(double));
    memcpy(work, indegree, (size_t)n * sizeof(int64_t));
    for (int64_t k = 0; k < h; ++k) {
        int64_t s = startnodes[k];
        int64_t e = endnodes[s];
        while (work[s] == 0) {
            o_t_next[s] += gain[s];
            if (s != e)
                i_t_next[e] += o_t_next[s];
            work[e] -= 1;
            s = e;
            e = endnodes[s];
        }
    }
}

/* ---- nutils.py:136-141  numba_init_inflows  (NO self-loop guard) ------------- */
TXO_API void txo_init_inflows(int64_t n, double *a, const int64_t *indices,
                              const double *b)
{
    for (int64_t i = 0; i < n; ++i)
        a[indices[i]] += b[i];
}

/* ---- nutils.py:143-169  _ap / _ap_par ----------------------------------------
 * out[:, c] = A . P[:, c] for every column c of the square C-order matrix P.
 * `threads` > 1 fans columns out the way numba's prange does. */
TXO_API void txo_ap(int64_t n, int64_t h, const double *P, double *out,
                    const int64_t *startnodes, const int64_t *endnodes,
                    const double *alpha, const double *beta, const double *chi,
                    const int64_t *indegree, int threads)
{
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
#endif
    {
        double *col = (double *)malloc((size_t)n * 4 * sizeof(double));
        int64_t *work = (int64_t *)malloc((size_t)n * sizeof(int64_t));
        double *o_prev = col, *i_prev = col + n, *i_next = col + 2 * n,
               *o_next = col + 3 * n;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t c = 0; c < n; ++c) {
            for (int64_t r = 0; r < n; ++r)
                o_prev[r] = P[r * n + c];
            memset(i_prev, 0, (size_t)n * sizeof(double));
            txo_init_inflows(n, i_prev, endnodes, o_prev);
            txo_ax(n, h, startnodes, endnodes, alpha, beta, chi, i_prev, o_prev,
                   indegree, i_next, o_next, work);
            for (int64_t r = 0; r < n; ++r)
                out[r * n + c] = o_next[r];
        }
        free(col);
        free(work);
    }
}

/* ---- nutils.py:172-214  _aqat / _aqat_par ------------------------------------
 * Two passes of _ap with `out = out.T` between (nutils.py:184,206).  The
 * reference returns the transposed VIEW of the caller's buffer; here `res`
 * receives the logical matrix the caller sees, res[i][j] = (returned view)[i, j],
 * in C order.  `tmp` is an n*n scratch. */
TXO_API void txo_aqat(int64_t n, int64_t h, const double *P, double *res,
                      double *tmp, const int64_t *startnodes,
                      const int64_t *endnodes, const double *alpha,
                      const double *beta, const double *chi,
                      const int64_t *indegree, int threads)
{
    /* pass 1: buf = A P  (buf is the caller's `out`, C order) */
    txo_ap(n, h, P, tmp, startnodes, endnodes, alpha, beta, chi, indegree, threads);
    /* out = out.T : view V with V[i,j] = buf[j,i].  Materialise V in C order. */
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j)
            res[i * n + j] = tmp[j * n + i];
    /* pass 2: V[:, c] = A . V[:, c] for every c -- in place, column by column
     * (each column is read fully before it is written: nutils.py:208-213). */
    memcpy(tmp, res, (size_t)n * n * sizeof(double));
    txo_ap(n, h, tmp, res, startnodes, endnodes, alpha, beta, chi, indegree, threads);
}

/* ---- nutils.py:5-39  interpolate_sample --------------------------------------
 * fp is C-order [T x m]; method 1 = linear, 0 = nearest (ties -> earlier row). */
TXO_API void txo_interpolate_sample(double x, int64_t T, int64_t m, const double *xp,
                                    const double *fp, int method, double *result)
{
    /* np.searchsorted(xp, x), side='left': first ix with xp[ix] >= x */
    int64_t lo = 0, hi = T;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (xp[mid] < x) lo = mid + 1; else hi = mid;
    }
    int64_t ix = lo;
    if (ix == 0) {
        memcpy(result, fp, (size_t)m * sizeof(double));
    } else if (ix >= T) {
        memcpy(result, fp + (T - 1) * m, (size_t)m * sizeof(double));
    } else {
        double dx_0 = x - xp[ix - 1];
        double dx_1 = xp[ix] - x;
        if (method == 1) {
            double frac = dx_0 / (dx_0 + dx_1);
            double w0 = 1 - frac;
            for (int64_t j = 0; j < m; ++j) {
                double a = w0 * fp[(ix - 1) * m + j];
                double b = frac * fp[ix * m + j];
                result[j] = a + b;
            }
        } else {
            const double *row = (fabs(dx_0) <= fabs(dx_1)) ? fp + (ix - 1) * m
                                                          : fp + ix * m;
            memcpy(result, row, (size_t)m * sizeof(double));
        }
    }
}

/* ---- muskingum.py:332-360  compute_alpha/beta/chi/gamma ---------------------- */
TXO_API void txo_coeffs(int64_t n, const double *K, const double *X, double dt,
                        double *alpha, double *beta, double *chi, double *gamma)
{
    for (int64_t j = 0; j < n; ++j) {
        double k = K[j], x = X[j];
        alpha[j] = (dt - 2 * k * x) / (2 * k * (1 - x) + dt);
        beta[j] = (dt + 2 * k * x) / (2 * k * (1 - x) + dt);
        chi[j] = (2 * k * (1 - x) - dt) / (2 * k * (1 - x) + dt);
        gamma[j] = dt / (k * (1 - x) + dt / 2);
    }
}

/* ---- muskingum.py:322-330  compute_indegree ---------------------------------- */
TXO_API void txo_indegree(int64_t n, const int64_t *startnodes,
                          const int64_t *endnodes, int64_t *indegree)
{
    memset(indegree, 0, (size_t)n * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i)
        indegree[endnodes[i]] += 1;
    for (int64_t i = 0; i < n; ++i)
        if (endnodes[i] == startnodes[i])
            indegree[i] -= 1;
}

/* ---- muskingum.py:410-419  init_states (i from o; self-loop included) -------- */
TXO_API void txo_init_states(int64_t n, const int64_t *startnodes,
                             const int64_t *endnodes, const double *o_t_next,
                             double *i_t_next)
{
    memset(i_t_next, 0, (size_t)n * sizeof(double));
    for (int64_t i = 0; i < n; ++i)
        i_t_next[endnodes[i]] += o_t_next[startnodes[i]];
}

/* ---- nutils.py:72-88 visit sequence (test hook) -------------------------------
 * Emits the order in which the reference's walk evaluates reaches; returns the
 * count (== n for a valid forest). */
TXO_API int64_t txo_visit_order(int64_t n, int64_t h, const int64_t *startnodes,
                                const int64_t *endnodes, const int64_t *indegree,
                                int64_t *order, int64_t *work)
{
    int64_t c = 0;
    memcpy(work, indegree, (size_t)n * sizeof(int64_t));
    for (int64_t k = 0; k < h; ++k) {
        int64_t s = startnodes[k];
        int64_t e = endnodes[s];
        while (work[s] == 0) {
            if (c < n) order[c] = s;
            ++c;
            work[e] -= 1;
            s = e;
            e = endnodes[s];
        }
    }
    return c;
}

/* ---- muskingum.py:499-533 simulate loop, members fanned out over threads -----
 * CPU baseline for the member-batched run: every member owns contiguous state
 * vectors (as each prange column does in nutils.py:157-169) and is stepped
 * `nsteps` times with _ax_bu; forcing rows are interpolated per step exactly as
 * muskingum.py:528-531 + nutils.py:5-39 do.  Member k's forcing is
 *   w0*wmul[r0*M+k]*fp[r0] + w1*wmul[r1*M+k]*fp[r1]   (wmul == NULL -> shared).
 * o_state / i_state are [M][n], updated in place.  Returns updates performed. */
TXO_API int64_t txo_run_members(int64_t n, int64_t h, int64_t M, int64_t nsteps,
                                const int64_t *startnodes, const int64_t *endnodes,
                                const double *alpha, const double *beta,
                                const double *chi, const double *gamma,
                                const int64_t *indegree, int64_t R, const double *xp,
                                const double *fp, const double *wmul, double t0,
                                double dt_ns, double *o_state, double *i_state,
                                int threads)
{
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
#endif
    {
        double *buf = (double *)malloc((size_t)n * 5 * sizeof(double));
        int64_t *work = (int64_t *)malloc((size_t)n * sizeof(int64_t));
        double *q = buf, *ia = buf + n, *oa = buf + 2 * n, *ib = buf + 3 * n,
               *ob = buf + 4 * n;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int64_t k = 0; k < M; ++k) {
            memcpy(oa, o_state + k * n, (size_t)n * sizeof(double));
            memcpy(ia, i_state + k * n, (size_t)n * sizeof(double));
            double *ip = ia, *op = oa, *in = ib, *on = ob;
            for (int64_t s = 0; s < nsteps; ++s) {
                double x = t0 + (double)(s + 1) * dt_ns;
                if (wmul == NULL) {
                    txo_interpolate_sample(x, R, n, xp, fp, 1, q);
                } else {
                    int64_t lo = 0, hi = R;
                    while (lo < hi) {
                        int64_t mid = lo + (hi - lo) / 2;
                        if (xp[mid] < x) lo = mid + 1; else hi = mid;
                    }
                    int64_t ix = lo;
                    if (ix == 0) {
                        for (int64_t j = 0; j < n; ++j) q[j] = wmul[k] * fp[j];
                    } else if (ix >= R) {
                        for (int64_t j = 0; j < n; ++j)
                            q[j] = wmul[(R - 1) * M + k] * fp[(R - 1) * n + j];
                    } else {
                        double dx_0 = x - xp[ix - 1], dx_1 = xp[ix] - x;
                        double frac = dx_0 / (dx_0 + dx_1);
                        double w0 = (1 - frac) * wmul[(ix - 1) * M + k];
                        double w1 = frac * wmul[ix * M + k];
                        for (int64_t j = 0; j < n; ++j)
                            q[j] = w0 * fp[(ix - 1) * n + j] + w1 * fp[ix * n + j];
                    }
                }
                txo_ax_bu(n, h, startnodes, endnodes, alpha, beta, chi, gamma, ip, op,
                          q, indegree, in, on, work);
                double *t;
                t = ip; ip = in; in = t;
                t = op; op = on; on = t;
            }
            memcpy(o_state + k * n, op, (size_t)n * sizeof(double));
            memcpy(i_state + k * n, ip, (size_t)n * sizeof(double));
        }
        free(buf);
        free(work);
    }
    return n * M * nsteps;
}

TXO_API int txo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
