// Spline order converter (second half): the reference's
// SmoothingSpline.generate_new_control_points (TG/spline_order_converter.py:22-34) as one more problem kind for the
// batched SQP of tg_sqp.h.
//
// Problem (one per old spline): control points Q [d][N] of a B-spline of order k that minimise
//     f(Q) = sum_c sum_t (Y[c][t] - p_Q,c(t))^2        over `resolution` samples t = linspace(0, N - k, R)   (:36-43)
// subject to 6 d linear equality rows: position, velocity and acceleration of the old spline at both ends (:45-63,
// rows = old - new, in the reference's order: positions [d][2] flattened, velocities, accelerations).
// The reference hands it to scipy SLSQP with default options (maxiter 100, ftol 1e-6) and 2-point finite-difference
// derivatives from the arc-length initial guess of create_initial_control_points (:83-112); here the SLSQP iteration
// is tg_sqp.h's (same line search, BFGS and QP stages as the trajectory problems), the objective gradient is
// scipy's forward difference (one perturbed objective per variable, h = 1.4901161193847656e-08) and the constant
// constraint rows are written once.
//
// Same source for the CUDA kernel (tg_smooth.cu, one warp per problem) and the single-lane host build of the test
// harness (tests/hostsim).
#ifndef TG_SMOOTH_H
#define TG_SMOOTH_H

#include "tg_sqp.h"

#define TG_SMOOTH_MAX_ORDER 5

// basis matrices of TG/matrix_evaluation.py:234-262 (orders 2 .. 5), element (l, col) with the scalar factor applied
// element-wise as numpy does
TG_HD double tg_basis_m(int order, int l, int col)
{
    const double M2[3][3] = {{1.0, -2.0, 1.0}, {-2.0, 2.0, 1.0}, {1.0, 0.0, 0.0}};
    const double M3[4][4] = {{-2.0, 6.0, -6.0, 2.0}, {6.0, -12.0, 0.0, 8.0}, {-6.0, 6.0, 6.0, 2.0}, {2.0, 0.0, 0.0, 0.0}};
    const double M4[5][5] = {{1.0, -4.0, 6.0, -4.0, 1.0}, {-4.0, 12.0, -6.0, -12.0, 11.0}, {6.0, -12.0, -6.0, 12.0, 11.0},
                             {-4.0, 4.0, 6.0, 4.0, 1.0}, {1.0, 0.0, 0.0, 0.0, 0.0}};
    const double M5[6][6] = {{-1.0, 5.0, -10.0, 10.0, -5.0, 1.0}, {5.0, -20.0, 20.0, 20.0, -50.0, 26.0},
                             {-10.0, 30.0, 0.0, -60.0, 0.0, 66.0}, {10.0, -20.0, -20.0, 20.0, 50.0, 26.0},
                             {-5.0, 5.0, 10.0, 10.0, 5.0, 1.0}, {1.0, 0.0, 0.0, 0.0, 0.0, 0.0}};
    return order == 2 ? 0.5 * M2[l][col] : order == 3 ? M3[l][col] / 12.0 : order == 4 ? M4[l][col] / 24.0 : M5[l][col] / 120.0;
}

// weights of the order + 1 control points of an interval at local parameter tau for the r-th derivative:
// w[l] = sum_col M[l][col] (order-col)! / (order-r-col)! tau^(order-r-col) / scale^r   (TG/matrix_evaluation.py:26-29, 127-130, 175-180)
TG_HD void tg_basis_weights(int order, double tau, int r, double scale, double *w)
{
    double sr = 1.0;
    for (int q = 0; q < r; q++) sr *= scale;
    for (int l = 0; l <= order; l++) {
        double h = 0;
        for (int col = 0; col <= order - r; col++) {
            double fac = 1.0, pw = 1.0;
            for (int q = 0; q < r; q++) fac *= (double)(order - col - q);
            for (int q = 0; q < order - r - col; q++) pw *= tau;
            h += tg_basis_m(order, l, col) * (r == 0 ? pw : fac / sr * pw);
        }
        w[l] = h;
    }
}

struct TgSmoothShape {
    int d, N, k, R;          // dimension, new control points, new order, resolution
    double scale;            // new scale factor (TG/spline_order_converter.py:77-82)
};

TG_HD int tg_smooth_table_doubles(const TgSmoothShape &S) { return S.R * (S.k + 2) + 6 * (S.k + 1); }
TG_HD int tg_smooth_par_doubles(const TgSmoothShape &S) { return S.d * S.R + 6 * S.d; }

// Shape-wide table: per sample t its interval (as a double) and the k + 1 position weights, then the end-point weights
// [r = 0, 1, 2][end e = 0, 1][k + 1].  Sample times are numpy.linspace(0, N - k, R) with the reference's interval
// assignment ((t >= i) & (t < i + 1), the last interval also takes t == i + 1).
TG_HD void tg_smooth_table_entry(const TgSmoothShape &S, int t, double *tab)
{
    const int nint = S.N - S.k, div = S.R - 1;
    const double step = div > 0 ? (double)nint / (double)div : 0.0;
#if defined(__CUDA_ARCH__)
    const double tt = (div > 0 && t == div) ? (double)nint : __dmul_rn((double)t, step);
#else
    const double tt = (div > 0 && t == div) ? (double)nint : (double)t * step;
#endif
    int i = (int)tt;
    if (i > nint - 1) i = nint - 1;
    double *row = tab + (size_t)t * (S.k + 2);
    row[0] = (double)i;
    tg_basis_weights(S.k, tt - (double)i, 0, 1.0, row + 1);
}

TG_HD void tg_smooth_end_entry(const TgSmoothShape &S, int q /* r * 2 + e */, double *tab)
{
    const int r = q >> 1, e = q & 1;
    tg_basis_weights(S.k, e ? 1.0 : 0.0, r, S.scale, tab + (size_t)S.R * (S.k + 2) + (size_t)q * (S.k + 1));
}

// synthetic layout for the SQP stages: n = d N variables without bounds, 6 d equality rows, nothing else
TG_HD void tg_smooth_layout(const TgSmoothShape &S, TgLayout *L)
{
    int *p = (int *)L;
    for (int i = 0; i < (int)(sizeof(TgLayout) / sizeof(int)); i++) p[i] = 0;
    L->d = S.d; L->N = S.N; L->nint = S.N - S.k;
    L->n = S.d * S.N;
    L->ia = L->n; L->it0 = L->n; L->is0 = -1; L->is1 = -1;          // no scale factor, scalars or times among the variables
    L->meq = 6 * S.d; L->m = 6 * S.d; L->mineq = 0;
    L->r_sfcl = L->m; L->r_sfcu = L->m; L->r_obs = L->m;              // no corridor rows
}

struct TgSmoothEval {
    TgSmoothShape S;
    const double *tab;       // shape table
    const double *par;       // [Y (d x R) | b (d x 6: pos e0 e1, vel e0 e1, acc e0 e1)]

    // objective at x (lane-parallel over samples)
    TG_MEMBER double objective(const double *x) const
    {
        const int k1 = S.k + 1;
        double h = 0;
        #pragma unroll 1
        for (int t = TG_LANE(); t < S.R; t += TG_NL) {
            const double *row = tab + (size_t)t * (S.k + 2);
            const int i = (int)row[0];
            #pragma unroll 1
            for (int c = 0; c < S.d; c++) {
                const double *q = x + c * S.N + i;
                double p = q[0] * row[1];
                for (int l = 1; l < k1; l++) p = p + q[l] * row[1 + l];
                const double e = par[c * S.R + t] - p;
                h += e * e;
            }
        }
        return tg_wsum(h);
    }

    // constant rows of A (old - new: minus the end-point weights), written once
    TG_MEMBER void init(const TgSqpWs &W) const
    {
        const int k1 = S.k + 1, nint = S.N - S.k;
        const double *cend = tab + (size_t)S.R * (S.k + 2);
        #pragma unroll 1
        for (int item = TG_LANE(); item < 6 * S.d * k1; item += TG_NL) {
            const int l = item % k1, rc = item / k1;          // rc = r * 2 d + c * 2 + e
            const int r = rc / (2 * S.d), ce = rc - r * 2 * S.d, c = ce >> 1, e = ce & 1;
            const int i = c * S.N + (e ? nint - 1 : 0) + l;
            W.A[i * W.lda + rc] = -cend[(r * 2 + e) * k1 + l];
        }
        TG_SYNC();
    }

    TG_MEMBER double value(const TgSqpWs &W) const
    {
        const int k1 = S.k + 1, nint = S.N - S.k;
        const double *cend = tab + (size_t)S.R * (S.k + 2);
        const double *b = par + (size_t)S.d * S.R;
        #pragma unroll 1
        for (int rc = TG_LANE(); rc < 6 * S.d; rc += TG_NL) {
            const int r = rc / (2 * S.d), ce = rc - r * 2 * S.d, c = ce >> 1, e = ce & 1;
            const double *q = W.x + c * S.N + (e ? nint - 1 : 0), *w = cend + (r * 2 + e) * k1;
            double p = q[0] * w[0];
            for (int l = 1; l < k1; l++) p = p + q[l] * w[l];
            W.c[rc] = b[c * 6 + r * 2 + e] - p;
        }
        const double f = objective(W.x);
        TG_SYNC();
        return f;
    }

    // scipy's approx_derivative('2-point', abs_step = 1.4901161193847656e-08) of the objective at W.x (f = its value there)
    TG_MEMBER void gradient(const TgSqpWs &W, double f) const
    {
        const int n = S.d * S.N;
        #pragma unroll 1
        for (int i = 0; i < n; i++) {
            const double xi = W.x[i];
            TG_SYNC();
            if (TG_LANE() == 0) W.x[i] = xi + TG_FD_STEP;
            TG_SYNC();
            const double dx = W.x[i] - xi;
            const double f1 = objective(W.x);
            TG_SYNC();
            if (TG_LANE() == 0) { W.g[i] = (f1 - f) / dx; W.x[i] = xi; }
            TG_SYNC();
        }
    }
};

// TG/spline_order_converter.py:83-112: equal arc-length steps along the old control polygon (one lane; old_pts [d][oldN]).
// scr: oldN doubles.
TG_HD void tg_smooth_initial_points(int d, const double *old_pts, int oldN, int N, double *out /* [d][N] */, double *dist)
{
    for (int i = 0; i < oldN - 1; i++) {
        double s = 0;
        for (int c = 0; c < d; c++) { const double v = old_pts[c * oldN + i + 1] - old_pts[c * oldN + i]; s += v * v; }
        dist[i] = sqrt(s);
    }
    for (int i = 0; i < oldN - 2; i++) dist[i + 1] = dist[i + 1] + dist[i];
    const int nseg = N - 1;
    const double step_len = dist[oldN - 2] / (double)nseg;
    int seg = 0;
    double cur = 0.0, step = 0.0, prev[3];
    for (int c = 0; c < d; c++) prev[c] = old_pts[c * oldN];
    for (int i = 0; i < nseg; i++) {
        double v[3], s = 0;
        for (int c = 0; c < d; c++) { v[c] = old_pts[c * oldN + seg + 1] - old_pts[c * oldN + seg]; s += v[c] * v[c]; }
        const double nv = sqrt(s);
        for (int c = 0; c < d; c++) { out[c * N + i] = prev[c] + v[c] / nv * step; prev[c] = out[c * N + i]; }
        step = step_len;
        cur = cur + step;
        if (dist[seg] < cur) {
            // first segment whose cumulative length is not below the distance walked (argmin of the non-negative differences)
            int best = 0;
            double bv = INFINITY;
            for (int j = 0; j < oldN - 1; j++) {
                double tdiff = dist[j] - cur;
                if (tdiff < 0) tdiff = INFINITY;
                if (tdiff < bv) { bv = tdiff; best = j; }
            }
            seg = best;
            // (numpy indexes distances[-1] when seg == 0: the last cumulative length)
            step = cur - dist[seg > 0 ? seg - 1 : oldN - 2];
            for (int c = 0; c < d; c++) prev[c] = old_pts[c * oldN + seg];
        }
    }
    for (int c = 0; c < d; c++) out[c * N + N - 1] = old_pts[c * oldN + oldN - 1];
}

// one problem, all stages until done (one warp on the device, one lane on the host).  x: in = initial control points,
// out = solution; ws: tg_sqp_workspace_doubles(L) doubles.
TG_FN void tg_smooth_solve(const TgSmoothShape &S, const double *tab, const double *par, double *x, double *ws, int maxiter,
                           double acc, TgSqpResult *res)
{
    TgLayout L;
    tg_smooth_layout(S, &L);
    TgSqpWs W;
    tg_sqp_carve(L, ws, &W);
    tg_sqp_begin(L, W, x, maxiter, acc, TG_SQP_FD_JACOBIAN);
    const TgSmoothEval ev = {S, tab, par};
    #pragma unroll 1
    for (;;) {
        const int st = W.ctl->state;
        if (st == TG_ST_DONE) break;
        if (st == TG_ST_INIT || st == TG_ST_LS) {
            tg_sqp_stage_ls_t(L, W, ev, (double *)0, 0);
            if (W.ctl->need_der) {
                ev.gradient(W, W.ctl->f);
                if (TG_LANE() == 0) { W.ctl->need_der = 0; W.ctl->nfev += L.n; }
                TG_SYNC();
            }
        } else tg_sqp_stage_qp<false>(L, W);
    }
    #pragma unroll 1
    for (int i = TG_LANE(); i < L.n; i += TG_NL) x[i] = W.x[i];
    TG_SYNC();
    if (res && TG_LANE() == 0) {
        res->status = W.ctl->status; res->nit = W.ctl->iter > maxiter ? maxiter : W.ctl->iter;
        res->nfev = W.ctl->nfev; res->f = W.ctl->f;
    }
}

#endif  // TG_SMOOTH_H
