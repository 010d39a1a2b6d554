/*
 * TEST INFRASTRUCTURE -- CPU oracle for the native geometry chain.
 *
 * Plain-C restatement of the algorithms in the reference's
 * trajectory_generation/constraint_functions/TrajectoryConstraintsCCode (CC/)
 * library.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this; the product (trajectory_generator_b200/csrc) never does.
 *
 * It deliberately keeps the reference's control flow (time-domain evaluation
 * with the scale factor, atan2/sin/cos rotations, exact-zero cascades, DBL_MAX
 * sentinels) so that it can be pinned against the reference's gtest golden
 * vectors (tests/golden/native_kats.json) and against the reference itself
 * compiled unmodified (oracle/_ref, see oracle/build_ref.sh).  Parity status:
 * PINNED by both (tests/test_oracle_native.py).
 *
 * Exports the same 24 C symbols as the reference library plus
 * a few tgo_* helpers that expose internal steps to the golden-vector tests.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TGO_MAX DBL_MAX

/* ---- CC/src/CubicEquationSolver.cpp:39-51 ---- */
static double solve_linear(double c, double d) { return c == 0 ? TGO_MAX : -d / c; }

/* ---- CC/src/CubicEquationSolver.cpp:53-72 ---- */
static void solve_quadratic(double b, double c, double d, double r[2])
{
    double disc = c * c - 4 * b * d;
    if (disc == 0) { r[0] = -c / (2 * b); r[1] = TGO_MAX; }
    else if (disc < 0) { r[0] = TGO_MAX; r[1] = TGO_MAX; }
    else { r[0] = (-c + sqrt(disc)) / (2 * b); r[1] = (-c - sqrt(disc)) / (2 * b); }
}

/* ---- CC/src/CubicEquationSolver.cpp:74-117 ---- */
static void solve_cubic(double a, double b, double c, double d, double r[3])
{
    double disc = 18 * a * b * c * d - 4 * (b * b * b) * d + (b * b) * (c * c) - 4 * a * (c * c * c) -
                  27 * (a * a) * (d * d);
    r[0] = r[1] = r[2] = TGO_MAX;
    if (disc > 0) {
        double Q = (3 * (c / a) - pow(b / a, 2)) / 9;
        double R = (9 * (b / a) * (c / a) - 27 * (d / a) - 2 * pow(b / a, 3)) / 54;
        double theta = acos(R / pow(-Q, 3.0 / 2.0));
        r[0] = 2 * sqrt(-Q) * cos(theta / 3) - (b / a) / 3;
        r[1] = 2 * sqrt(-Q) * cos((theta + 2 * M_PI) / 3) - (b / a) / 3;
        r[2] = 2 * sqrt(-Q) * cos((theta + 4 * M_PI) / 3) - (b / a) / 3;
    } else if (disc < 0) {
        double P = b * b - 3 * a * c;
        double Q = 9 * a * b * c - 2 * (b * b * b) - 27 * (a * a) * d;
        double t1 = Q / 2 + sqrt((Q * Q) / 4 - pow(P, 3));
        double t2 = Q / 2 - sqrt((Q * Q) / 4 - pow(P, 3));
        double N = cbrt(t1) + cbrt(t2);
        r[0] = -b / (3 * a) + N / (3 * a);
    } else {
        double P = b * b - 3 * a * c;
        if (P == 0) r[0] = -b / (3 * a);
        else {
            r[0] = (9 * a * d - b * c) / (2 * P);
            r[1] = (4 * a * b * c - 9 * a * a * d - b * b * b) / (a * P);
        }
    }
}

/* ---- CC/src/CubicEquationSolver.cpp:8-36 ---- */
void tgo_solve_equation(double a, double b, double c, double d, double r[3])
{
    r[0] = r[1] = r[2] = TGO_MAX;
    if (a == 0) {
        if (b == 0) { if (c != 0) r[0] = solve_linear(c, d); }
        else solve_quadratic(b, c, d, r);
    } else solve_cubic(a, b, c, d, r);
}

/* interval control points: cp[c*4 + l], c < D, l < 4 (CC/src/CBindHelperFunctions.cpp:11-31) */
static void take_interval(const double *pts, int N, int j, int D, double *cp)
{
    for (int c = 0; c < D; c++)
        for (int l = 0; l < 4; l++) cp[c * 4 + l] = pts[c * N + j + l];
}

/* columns 0..2 of P*M, M from CC/src/DerivativeEvaluator.cpp:56-65 */
static void pm_columns(const double *cp, int D, double *k3, double *k2, double *k1)
{
    for (int c = 0; c < D; c++) {
        const double *p = cp + c * 4;
        k3[c] = p[0] * (-1 / 6.0) + p[1] * (1 / 2.0) + p[2] * (-1 / 2.0) + p[3] * (1 / 6.0);
        k2[c] = p[0] * (1 / 2.0) + p[1] * (-1.0) + p[2] * (1 / 2.0);
        k1[c] = p[0] * (-1 / 2.0) + p[2] * (1 / 2.0);
    }
}

/* ---- CC/src/DerivativeEvaluator.cpp:22-29,83-104 ---- */
static void velocity_at(const double *cp, int D, double t, double a, double *v)
{
    double k3[3], k2[3], k1[3];
    pm_columns(cp, D, k3, k2, k1);
    double T0 = 3 * t * t / (a * a * a), T1 = 2 * t / (a * a), T2 = 1 / a;
    for (int c = 0; c < D; c++) v[c] = k3[c] * T0 + k2[c] * T1 + k1[c] * T2;
}

/* ---- CC/src/DerivativeEvaluator.cpp:39-46,107-128 ---- */
static void acceleration_at(const double *cp, int D, double t, double a, double *acc)
{
    double k3[3], k2[3], k1[3];
    pm_columns(cp, D, k3, k2, k1);
    double T0 = 6 * t / (a * a * a), T1 = 2 / (a * a);
    for (int c = 0; c < D; c++) acc[c] = k3[c] * T0 + k2[c] * T1;
}

static double norm(const double *v, int D)
{
    double s = 0;
    for (int c = 0; c < D; c++) s += v[c] * v[c];
    return sqrt(s);
}
static double dotp(const double *a, const double *b, int D)
{
    double s = 0;
    for (int c = 0; c < D; c++) s += a[c] * b[c];
    return s;
}
static double speed_at(const double *cp, int D, double t, double a) { double v[3]; velocity_at(cp, D, t, a, v); return norm(v, D); }
static double accel_mag_at(const double *cp, int D, double t, double a) { double v[3]; acceleration_at(cp, D, t, a, v); return norm(v, D); }

/* ---- CC/src/DerivativeBounds.cpp:110-123: roots of d|v|^2/dtau, scaled to time ---- */
static void velocity_roots(const double *cp, int D, double a, double r[3])
{
    double k3[3], k2[3], k1[3];
    pm_columns(cp, D, k3, k2, k1);
    double J00 = dotp(k3, k3, D), J01 = dotp(k3, k2, D), J11 = dotp(k2, k2, D), J20 = dotp(k1, k3, D), J21 = dotp(k1, k2, D);
    tgo_solve_equation(36 * J00, 12 * J01 + 24 * J01, 8 * J11 + 12 * J20, 4 * J21, r);
    for (int i = 0; i < 3; i++) r[i] *= a;
}

/* ---- CC/src/DerivativeBounds.cpp:48-76 ---- */
static void min_velocity_and_time(const double *cp, int D, double a, double *vmin, double *tmin)
{
    double r[3];
    velocity_roots(cp, D, a, r);
    double best = speed_at(cp, D, 0, a), tb = 0;
    double vf = speed_at(cp, D, a, a);
    if (vf < best) { best = vf; tb = a; }
    for (int i = 0; i < 3; i++)
        if (r[i] > 0 && r[i] < a) {
            double v = speed_at(cp, D, r[i], a);
            if (v < best) { best = v; tb = r[i]; }
        }
    *vmin = best; *tmin = tb;
}

/* ---- CC/src/DerivativeBounds.cpp:128-142 ---- */
static double max_acceleration(const double *cp, int D, double a)
{
    double a0 = accel_mag_at(cp, D, 0, a), a1 = accel_mag_at(cp, D, a, a);
    return a1 > a0 ? a1 : a0;
}

static void cross3(const double *a, const double *b, int D, double *o)
{
    if (D == 2) { o[0] = a[0] * b[1] - a[1] * b[0]; o[1] = o[2] = 0; }
    else { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
}

/* ---- CC/src/CrossTermEvaluator.cpp:13-19,75-92 ---- */
static double cross_term_at(const double *cp, int D, double t, double a)
{
    double v[3], acc[3], x[3];
    velocity_at(cp, D, t, a, v);
    acceleration_at(cp, D, t, a, acc);
    cross3(v, acc, D, x);
    return D == 2 ? fabs(x[0]) : sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
}

/* ---- CC/src/CrossTermProperties.cpp:13-101 ----
 * The reference spells d|v x a|^2/dtau / 2 out as polynomials in the control
 * point coordinates.  With k3,k2,k1 = columns of P*M, v x a (times alpha^3) is
 * 6 U tau^2 + 6 V tau + 2 W,  U = k2 x k3, V = k1 x k3, W = k1 x k2, hence the
 * cubic (72 U.U, 108 U.V, 36 V.V + 24 U.W, 12 V.W) -- the same four numbers
 * (goldens: UnitTestCrossTermProperties.cpp:4-45). */
void tgo_cross_coefficients(const double *cp, int D, double c[4])
{
    double k3[3], k2[3], k1[3], U[3], V[3], W[3];
    pm_columns(cp, D, k3, k2, k1);
    cross3(k2, k3, D, U); cross3(k1, k3, D, V); cross3(k1, k2, D, W);
    c[0] = 72 * dotp(U, U, 3); c[1] = 108 * dotp(U, V, 3);
    c[2] = 36 * dotp(V, V, 3) + 24 * dotp(U, W, 3); c[3] = 12 * dotp(V, W, 3);
}

/* ---- CC/src/CrossTermBounds.cpp:168-200 ---- */
double tgo_max_cross_term(const double *cp, int D, double a)
{
    double c[4], r[3];
    tgo_cross_coefficients(cp, D, c);
    tgo_solve_equation(c[0], c[1], c[2], c[3], r);
    double best = cross_term_at(cp, D, 0, a), xf = cross_term_at(cp, D, a, a);
    if (xf > best) best = xf;
    for (int i = 0; i < 3; i++) {
        double t = r[i] * a;
        if (t > 0 && t < a) { double x = cross_term_at(cp, D, t, a); if (x > best) best = x; }
    }
    return best;
}

/* ---- CC/src/CrossTermBounds.cpp:64-152: kind 1 curvature (alpha forced to 1), 2 angular rate, 3 centripetal ---- */
double tgo_interval_turn_bound(const double *cp, int D, double a, int kind)
{
    if (kind == 1) a = 1;
    double vmin, tmin;
    min_velocity_and_time(cp, D, a, &vmin, &tmin);
    double cmax = tgo_max_cross_term(cp, D, a);
    double amax = max_acceleration(cp, D, a);
    if (vmin <= 1.0e-8) {
        if (kind == 3) return 0;
        return accel_mag_at(cp, D, tmin, a) <= 1.0e-8 ? 0 : TGO_MAX;
    }
    double b1, b2;
    if (kind == 1) { b1 = amax / (vmin * vmin); b2 = cmax / (vmin * vmin * vmin); }
    else if (kind == 2) { b1 = amax / vmin; b2 = cmax / (vmin * vmin); }
    else { b1 = amax; b2 = cmax / vmin; }
    return b2 < b1 ? b2 : b1;
}

/* ---- CC/src/CrossTermBounds.cpp:13-61 ---- */
static double spline_turn_bound(const double *pts, int N, int D, double a, int kind)
{
    double best = 0, cp[12];
    for (int j = 0; j < N - 3; j++) {
        take_interval(pts, N, j, D, cp);
        double b = tgo_interval_turn_bound(cp, D, a, kind);
        if (b > best) best = b;
    }
    return best;
}

/* ---- CC/src/DerivativeBounds.cpp:12-27 ---- */
static double spline_min_velocity(const double *pts, int N, int D, double a)
{
    double best = TGO_MAX, cp[12], v, t;
    for (int j = 0; j < N - 3; j++) {
        take_interval(pts, N, j, D, cp);
        min_velocity_and_time(cp, D, a, &v, &t);
        if (v < best) best = v;
    }
    return best;
}
double tgo_interval_min_velocity(const double *cp, int D, double a) { double v, t; min_velocity_and_time(cp, D, a, &v, &t); return v; }
double tgo_interval_max_acceleration(const double *cp, int D, double a) { return max_acceleration(cp, D, a); }

/* ---- CC/src/BsplineToMinvo.cpp:87-96 (third order), Q = P * Mc ---- */
static const double MINVO3[4][4] = {
    {0.18372189964688778830269864557208, 0.057009542139797595613306102386893, -0.015455156825262485566573649098775, -0.0053387946850481119404479942697845},
    {0.7017652268843997637057156686535, 0.66657381574108923111064205020873, 0.2918717989443756838876956809183, 0.11985166815376058497710386445935},
    {0.11985166815376058497710386445935, 0.2918717989443756838876956809183, 0.66657381574108923111064205020873, 0.7017652268843997637057156686535},
    {-0.0053387946850481119404479942697845, -0.015455156825262485566573649098775, 0.057009542139797595613306102386893, 0.18372189964688778830269864557208}};

static void to_minvo(const double *cp, int D, double *q)
{
    for (int c = 0; c < D; c++)
        for (int k = 0; k < 4; k++) {
            double s = 0;
            for (int l = 0; l < 4; l++) s += cp[c * 4 + l] * MINVO3[l][k];
            q[c * 4 + k] = s;
        }
}

/* ---- CC/src/SphereCollisionEvaluator.cpp:88-154, CC/src/RotationHelperFunctions.cpp:11-69 ---- */
static double hull_distance_to_sphere(const double *q, int D, const double *center, double radius)
{
    double w[3] = {0, 0, 0}, R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int c = 0; c < D; c++) w[c] = (q[c * 4] + q[c * 4 + 1] + q[c * 4 + 2] + q[c * 4 + 3]) / 4 - center[c];
    if (D == 2) {
        if (!(w[0] == 0 && w[1] == 0)) {
            double psi = -atan2(w[1], w[0]);
            R[0][0] = cos(psi); R[0][1] = -sin(psi); R[1][0] = sin(psi); R[1][1] = cos(psi);
        }
    } else if (!(w[0] == 0 && w[1] == 0 && w[2] == 0)) {
        /* angles = (psi, -theta); rotation = Rz(-psi) * Ry(theta) */
        double theta = atan2(w[2], w[0]);
        double ct = cos(theta), st = sin(theta);
        double Ry[3][3] = {{ct, 0, st}, {0, 1, 0}, {-st, 0, ct}};
        double w2x = Ry[0][0] * w[0] + Ry[0][2] * w[2], w2y = w[1];
        double psi = -atan2(w2y, w2x);
        double cp_ = cos(psi), sp = sin(psi);
        double Rz[3][3] = {{cp_, -sp, 0}, {sp, cp_, 0}, {0, 0, 1}};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0;
                for (int k = 0; k < 3; k++) s += Rz[i][k] * Ry[k][j];
                R[i][j] = s;
            }
    }
    double minx = TGO_MAX;
    for (int k = 0; k < 4; k++) {
        double x = 0;
        for (int c = 0; c < D; c++) x += R[0][c] * q[c * 4 + k];
        if (x < minx) minx = x;
    }
    double cx = 0;
    for (int c = 0; c < D; c++) cx += R[0][c] * center[c];
    return minx - (cx + radius);
}

/* ---- CC/src/SphereCollisionEvaluator.cpp:13-45 ---- */
static double *spline_distances_to_spheres(int D, const double *centers, const double *radii, int K, const double *pts, int N)
{
    double *out = (double *)malloc(sizeof(double) * (K > 0 ? K : 1));
    for (int i = 0; i < K; i++) {
        double c[3] = {0, 0, 0}, best = TGO_MAX, cp[12], q[12];
        for (int k = 0; k < D; k++) c[k] = centers[i + k * K];
        for (int j = 0; j < N - 3; j++) {
            take_interval(pts, N, j, D, cp);
            to_minvo(cp, D, q);
            double dist = hull_distance_to_sphere(q, D, c, radii[i]);
            if (best > dist) best = dist;
        }
        out[i] = best;
    }
    return out;
}

/* ---- CC/src/SphereCollisionEvaluator.cpp:70-86 ---- */
static double *interval_distances_to_sphere(int D, const double *pts, int N, double radius, const double *center)
{
    double *out = (double *)malloc(sizeof(double) * (N > 3 ? N - 3 : 1));
    double cp[12], q[12];
    for (int j = 0; j < N - 3; j++) {
        take_interval(pts, N, j, D, cp);
        to_minvo(cp, D, q);
        out[j] = hull_distance_to_sphere(q, D, center, radius);
    }
    return out;
}

/* ---- CC/src/MDMAlgorithmClass.cpp:11-71 (points[c*npts + i]) ---- */
double tgo_mdm_min_norm(const double *points, int D, int npts, int max_iterations, double tolerance)
{
    double p[16] = {0}, cur[3];
    int supp[16], nsupp = 1, iterations = 0;
    double delta_p = 1.0;
    for (int c = 0; c < D; c++) cur[c] = points[c * npts];
    supp[0] = 0; p[0] = 1;
    while (delta_p > 0.000001 && iterations < max_iterations && nsupp > 0) {
        int max_index = supp[0], min_index = 0;
        double best = DBL_MIN;   /* reference uses numeric_limits<double>::min() (smallest positive) */
        int first = 1;
        (void)first;
        max_index = supp[0];
        {
            int arg = 0;
            for (int i = 0; i < nsupp; i++) {
                double s = 0;
                for (int c = 0; c < D; c++) s += points[c * npts + supp[i]] * cur[c];
                if (s > best) { best = s; arg = i; }
            }
            max_index = supp[arg];
        }
        {
            double lo = DBL_MAX;
            for (int i = 0; i < npts; i++) {
                double s = 0;
                for (int c = 0; c < D; c++) s += points[c * npts + i] * cur[c];
                if (s < lo) { lo = s; min_index = i; }
            }
        }
        double diff[3], dn2 = 0;
        delta_p = 0;
        for (int c = 0; c < D; c++) {
            diff[c] = points[c * npts + max_index] - points[c * npts + min_index];
            delta_p += diff[c] * cur[c];
            dn2 += diff[c] * diff[c];
        }
        if (delta_p > tolerance) {
            double dn = sqrt(dn2);
            double t = delta_p / (p[max_index] * dn * dn);
            if (t >= 1) t = 1.0;
            for (int c = 0; c < D; c++) cur[c] = cur[c] - (t * p[max_index] * diff[c]);
            double t1 = t * p[max_index], t2 = 1 - t;
            p[min_index] += t1;
            p[max_index] *= t2;
            nsupp = 0;
            for (int i = 0; i < npts; i++) if (p[i] > tolerance) supp[nsupp++] = i;
            iterations += 1;
        }
    }
    return norm(cur, D);
}

/* ---- CC/src/ControlPointDerivativeBounds.cpp:13-46 ---- */
static double min_velocity_of_bez_vel_cont_pts(const double *pts, int n, int D)
{
    double best = TGO_MAX, tri[9];
    int nint = (n - 1) / 2;
    for (int i = 0; i < nint; i++) {
        for (int c = 0; c < D; c++)
            for (int l = 0; l < 3; l++) tri[c * 3 + l] = pts[c * n + 2 * i + l];
        double v = tgo_mdm_min_norm(tri, D, 3, 500, 0.000001);
        if (v < best) best = v;
    }
    return best;
}

/* ================= the reference's 24 extern "C" symbols =================
 * CC/include/CrossTermBounds.hpp:45-72, ObstacleConstraints.hpp:25-49,
 * ControlPointDerivativeBounds.hpp:26-33.  Handles are opaque and unused. */
static int g_handle;
#define DEFINE_DIM(D)                                                                                         \
    void *CrossTermBounds_##D(void) { return &g_handle; }                                                     \
    double get_spline_curvature_bound_##D(void *o, double *p, int N) { (void)o; return spline_turn_bound(p, N, D, 1.0, 1); } \
    double get_spline_angular_rate_bound_##D(void *o, double *p, int N, double a) { (void)o; return spline_turn_bound(p, N, D, a, 2); } \
    double get_spline_centripetal_acceleration_bound_##D(void *o, double *p, int N, double a) { (void)o; return spline_turn_bound(p, N, D, a, 3); } \
    void *DerivativeBounds_##D(void) { return &g_handle; }                                                    \
    double find_min_velocity_of_spline_##D(void *o, double *p, int N, double a) { (void)o; return spline_min_velocity(p, N, D, a); } \
    void *ObstacleConstraints_##D(void) { return &g_handle; }                                                 \
    double *getObstaclesConstraintsForSpline_##D(void *o, double *c, double *r, int K, double *p, int N) { (void)o; return spline_distances_to_spheres(D, c, r, K, p, N); } \
    double *getObstacleConstraintsForIntervals_##D(void *o, double *p, int N, double r, double *c) { (void)o; return interval_distances_to_sphere(D, p, N, r, c); } \
    double getObstacleConstraintForSpline_##D(void *o, double *p, int N, double r, double *c) {               \
        (void)o; double *a = interval_distances_to_sphere(D, p, N, r, c); double b = TGO_MAX;                  \
        for (int j = 0; j < N - 3; j++) if (b > a[j]) b = a[j];                                               \
        free(a); return b; }                                                                                  \
    void *ControlPointDerivativeBounds_##D(void) { return &g_handle; }                                        \
    double find_min_velocity_of_bez_vel_cont_pts_##D(void *o, double *p, int n) { (void)o; return min_velocity_of_bez_vel_cont_pts(p, n, D); }

DEFINE_DIM(2)
DEFINE_DIM(3)

/* helpers for golden-vector tests of internal steps */
double tgo_hull_distance(const double *q, int D, const double *center, double radius) { return hull_distance_to_sphere(q, D, center, radius); }
void tgo_to_minvo(const double *cp, int D, double *q) { to_minvo(cp, D, q); }
void tgo_free(void *p) { free(p); }
