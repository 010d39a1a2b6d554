// TEST INFRASTRUCTURE -- single-lane host build of the device headers
// (csrc/tg_eval.h, csrc/tg_sqp.h) so that the evaluation maths and the SQP
// logic can be unit-tested in the GPU-less build container against the oracle
// and scipy.  Never loaded by the product: the product path is the CUDA
// library built from csrc/tg_api.cu and fails loudly without it.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../trajectory_generator_b200/csrc/tg_eval.h"
#ifdef TG_WITH_SQP
#include "../../trajectory_generator_b200/csrc/tg_sqp.h"
#include "../../trajectory_generator_b200/csrc/tg_smooth.h"
#endif

extern "C" int hs_layout(const int *spec, int *out, int cap)
{
    TgLayout L;
    tg_make_layout(spec, &L);
    const int cnt = (int)(sizeof(TgLayout) / sizeof(int));
    if (out && cap >= cnt) memcpy(out, &L, sizeof(TgLayout));
    return cnt;
}

// f, g[n], c[m], J[m*n] row-major (all rows, linear ones included)
extern "C" void hs_eval(const int *spec, const double *par, const double *x, double *f, double *g, double *c, double *J)
{
    TgLayout L;
    tg_make_layout(spec, &L);
    std::vector<double> scratch(tg_scratch_doubles(L));
    *f = tg_objective(L, spec, x, g);
    TgJac sink = {J, L.n, 1, 0};
    if (J) {
        if (L.d == 2) tg_linear_jacobian_d<2>(L, spec, par, sink); else tg_linear_jacobian_d<3>(L, spec, par, sink);
    }
    if (L.d == 2) tg_constraints_d<2>(L, spec, par, x, c, J ? &sink : nullptr, scratch.data());
    else tg_constraints_d<3>(L, spec, par, x, c, J ? &sink : nullptr, scratch.data());
}

#ifdef TG_WITH_SQP
extern "C" int hs_solve(const int *spec, const double *par, double *x, int maxiter, double ftol, int flags, double *fout,
                        int *status, int *nit, int *nfev, double *trace, int trace_cap)
{
    TgLayout L;
    tg_make_layout(spec, &L);
    size_t nd = tg_sqp_workspace_doubles(L);
    std::vector<double> ws(nd);
    TgSqpResult res;
    if (L.d == 2) tg_sqp_solve<2>(L, spec, par, x, ws.data(), maxiter, ftol, flags, &res, trace, trace_cap);
    else tg_sqp_solve<3>(L, spec, par, x, ws.data(), maxiter, ftol, flags, &res, trace, trace_cap);
    *fout = res.f; *status = res.status; *nit = res.nit; *nfev = res.nfev;
    return res.status;
}
#endif

#ifdef TG_WITH_SQP
// the solver's finite-difference derivatives at x (scipy's approx_derivative emulation): g[n], J[m*n] row-major
// (linear rows hold their constant coefficients)
extern "C" void hs_fd_derivatives(const int *spec, const double *par, const double *x, const double *xl, const double *xu,
                                  double *g, double *J)
{
    TgLayout L;
    tg_make_layout(spec, &L);
    std::vector<double> ws(tg_sqp_workspace_doubles(L));
    TgSqpWs W;
    tg_sqp_carve(L, ws.data(), &W);
    tg_sqp_begin(L, W, x, 100, 1e-6, TG_SQP_FD_JACOBIAN);
    for (int i = 0; i < L.n; i++) { W.xl[i] = xl[i]; W.xu[i] = xu[i]; }
    TgJac sink = {W.A, 1, W.lda, 2};          // the solver's A: every row except the corridor rows
    TgJac full = {J, L.n, 1, 0};              // corridor rows (constant) straight into the dense output
    double f;
    if (L.n_sfc) { if (L.d == 2) tg_jac_sfc<2>(L, spec, par, full); else tg_jac_sfc<3>(L, spec, par, full); }
    if (L.d == 2) {
        tg_linear_jacobian_d<2>(L, spec, par, sink);
        f = tg_sqp_evaluate<2>(L, spec, par, W, false);
        tg_sqp_fd_derivatives<2>(L, spec, par, W, f);
    } else {
        tg_linear_jacobian_d<3>(L, spec, par, sink);
        f = tg_sqp_evaluate<3>(L, spec, par, W, false);
        tg_sqp_fd_derivatives<3>(L, spec, par, W, f);
    }
    for (int i = 0; i < L.n; i++) {
        g[i] = W.g[i];
        for (int j = 0; j < L.m; j++)
            if (!tg_is_sfc_row(W, j)) J[j * L.n + i] = W.A[i * W.lda + tg_arow(W, j)];
    }
}
#endif

#ifdef TG_WITH_SQP
// the BFGS factor update on its own: L D L' <- L D L' + sigma z z'.  Lm: n x n, unit lower factor stored as the solver
// stores it (column i at Lm[i*n + j], j > i); Dd: n.  Both are updated in place.
extern "C" void hs_ldl_update(int n, double sigma, const double *z, double *Lm, double *Dd)
{
    std::vector<double> zz(z, z + n), w(n + 1), sc(5 * (size_t)n + 5);
    tg_ldl_update(n, sigma, zz.data(), Lm, Dd, w.data(), sc.data(), n);
}
#endif


#ifdef TG_WITH_SQP
// spline order converter (csrc/tg_smooth.h) on the host: table + one solve.  par = [Y | b], x in/out
extern "C" int hs_smooth_solve(int d, int N, int order, int resolution, double scale, const double *par, double *x,
                               double *fout, int *nit)
{
    const TgSmoothShape S = {d, N, order, resolution, scale};
    std::vector<double> tab(tg_smooth_table_doubles(S));
    for (int t = 0; t < resolution; t++) tg_smooth_table_entry(S, t, tab.data());
    for (int q = 0; q < 6; q++) tg_smooth_end_entry(S, q, tab.data());
    TgLayout L;
    tg_smooth_layout(S, &L);
    std::vector<double> ws(tg_sqp_workspace_doubles(L));
    TgSqpResult res;
    tg_smooth_solve(S, tab.data(), par, x, ws.data(), 100, 1e-6, &res);
    *fout = res.f; *nit = res.nit;
    return res.status;
}

extern "C" void hs_smooth_initial(int d, const double *old_pts, int oldN, int N, double *out)
{
    std::vector<double> scr(oldN);
    tg_smooth_initial_points(d, old_pts, oldN, N, out, scr.data());
}
#endif
