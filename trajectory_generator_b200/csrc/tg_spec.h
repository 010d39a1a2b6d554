// Problem "shape" descriptor shared by the host C-ABI, the CUDA kernels and
// the host-simulation test build.  One descriptor describes every problem of a
// batch: all per-problem numbers live in a flat double parameter row.
//
// It encodes which constraint blocks TrajectoryGenerator.__get_constraints
// (reference TG/trajectory_generator.py:171-250) would build and in which
// order scipy's SLSQP sees their rows after new_constraint_to_old
// (scipy/optimize/_constraints.py:506-601): equality blocks first, then
// inequality blocks, two-sided blocks as all (y-lb) rows then all (ub-y) rows.
#ifndef TG_SPEC_H
#define TG_SPEC_H

#if defined(__CUDACC__)
#define TG_HD __host__ __device__ __forceinline__
#define TG_MEMBER __host__ __device__ __forceinline__       // member functions (no `static`)
#else
#define TG_HD static inline
#define TG_MEMBER inline
#endif

#define TG_MAX_CORRIDORS 8

// indices into the int32 spec array (the C-ABI passes `const int *spec`)
enum TgSpecField {
    TG_SP_DIM = 0,      // 2 or 3
    TG_SP_NCP,          // number of control points N (>= 4)
    TG_SP_OBJECTIVE,    // TgObjective
    TG_SP_START_KIND,   // 0: location rows (d)         1: zero-velocity rows (3d)
    TG_SP_END_KIND,     // 0: location rows  1: zero-velocity rows  2: target rows (alpha column)
    TG_SP_START_DIR,    // 0 none, 1: s*(P2-P0)/2, 2: zero-velocity variant s*(P3-P0)/2
    TG_SP_START_VEL,    // velocity rows (|v|>0)
    TG_SP_START_ACC,
    TG_SP_END_DIR,
    TG_SP_END_VEL,
    TG_SP_END_ACC,
    TG_SP_NIW,          // number of intermediate waypoints
    TG_SP_IW_VEL,       // intermediate velocity rows present
    TG_SP_DB_MINV,      // DerivativeBounds fields present (reference DS/dynamic_bounds.py:32-41)
    TG_SP_DB_MAXV,
    TG_SP_DB_UP,
    TG_SP_DB_HORIZ,
    TG_SP_DB_MAXA,
    TG_SP_DB_GRAV,
    TG_SP_DB_JERK,
    TG_SP_TANG,         // tangential acceleration rows
    TG_SP_TURN,         // TgTurn
    TG_SP_NCORR,        // number of safe-flight corridors (0: no SFC block)
    TG_SP_IPC0,         // intervals per corridor [TG_MAX_CORRIDORS]
    TG_SP_NOBST = TG_SP_IPC0 + TG_MAX_CORRIDORS,
    TG_SP_COUNT
};

enum TgObjective {   // reference TG/trajectory_generator.py:114-132 -> TG/objectives/objective_functions.py
    TG_OBJ_TIME = 0,              // "minimal_time_path"                 alpha^2
    TG_OBJ_DIST,                  // "minimal_distance_path"             sum |D1|^2
    TG_OBJ_VEL,                   // "minimal_velocity_path"             sum |D2|^2
    TG_OBJ_ACC,                   // "minimal_acceleration_path"         sum |D3|^2
    TG_OBJ_DIST_TIME,             // "minimal_distance_and_time_path"    alpha sum |D1|^2
    TG_OBJ_VEL_TIME,              // "minimal_velocity_and_time_path"    alpha sum |D2|^2
    TG_OBJ_ACC_TIME,              // "minimal_acceleration_and_time_path" alpha sum |D3|^2
    TG_OBJ_TIME_VEL_PENALTY       // "minimal_time_path_velocity_penalty" 100 alpha^2 - sum |D1|^2
};

enum TgTurn { TG_TURN_NONE = 0, TG_TURN_CURVATURE, TG_TURN_ANGULAR_RATE, TG_TURN_CENTRIPETAL };

// Everything derived from a spec: variable count, row counts / offsets in the
// SLSQP-ordered constraint vector, offsets in the nonlinear-Jacobian row list
// and in the per-problem parameter row.
struct TgLayout {
    int d, N, nint;           // dimension, control points, intervals (N-3)
    int n;                    // number of optimisation variables
    int ia;                   // index of alpha in x (= d*N)
    int is0, is1;             // index of start / end direction scalar in x (-1 if none)
    int it0;                  // index of first intermediate scale time in x
    int nws, niw;
    // ---- rows (SLSQP order) ----
    int meq, mineq, m;
    int r_start, n_start;     // linear
    int r_end, n_end;         // linear
    int r_sder, n_sder;       // nonlinear eq
    int r_eder, n_eder;
    int r_iwl, n_iwl;
    int r_iwv, n_iwv;
    int r_db, n_db;           // inequality blocks (>= 0 form)
    int r_tanl, r_tanu, n_tan;  // n_tan rows in each of the two blocks
    int r_turn, n_turn;
    int r_sfcl, r_sfcu, n_sfc;  // n_sfc rows in each of the two blocks
    int r_obs, n_obs;
    int m_nl;                 // number of nonlinear rows
    int m_lin;                // number of linear rows (= m - m_nl)
    // ---- turning block works on a trimmed control point window ----
    int turn_first, turn_ncp; // first control point and count passed to the turning bound
    // ---- parameter row offsets ----
    int p_start_loc, p_end_loc, p_target_vel;
    int p_sdir, p_svel, p_sacc, p_edir, p_evel, p_eacc;
    int p_iwl, p_iwv;
    int p_minv, p_maxv, p_up, p_horiz, p_maxa, p_grav, p_jerk;
    int p_tanmin, p_tanmax, p_turn;
    int p_sfc;                // per corridor: rotT[d*d], lb[d], ub[d]
    int p_obs_c, p_obs_r;     // centers[d*K] (row-major d x K), radii[K]
    int P;                    // parameter row length
};

TG_HD int tg_sfc_stride(int d) { return d * d + 2 * d; }

TG_HD void tg_make_layout(const int *sp, TgLayout *L)
{
    const int d = sp[TG_SP_DIM], N = sp[TG_SP_NCP];
    L->d = d; L->N = N; L->nint = N - 3;
    L->ia = d * N;
    int nx = d * N + 1;
    L->is0 = -1; L->is1 = -1;
    L->nws = 0;
    if (sp[TG_SP_START_DIR]) { L->is0 = nx++; L->nws++; }
    if (sp[TG_SP_END_DIR]) { L->is1 = nx++; L->nws++; }
    L->niw = sp[TG_SP_NIW];
    L->it0 = nx;
    nx += L->niw;
    L->n = nx;

    int r = 0, nl = 0;
    L->r_start = r; L->n_start = sp[TG_SP_START_KIND] == 1 ? 3 * d : d; r += L->n_start;
    L->r_end = r; L->n_end = sp[TG_SP_END_KIND] == 1 ? 3 * d : d; r += L->n_end;
    L->r_sder = r; L->n_sder = d * ((sp[TG_SP_START_DIR] ? 1 : 0) + (sp[TG_SP_START_VEL] ? 1 : 0) + (sp[TG_SP_START_ACC] ? 1 : 0)); r += L->n_sder;
    L->r_eder = r; L->n_eder = d * ((sp[TG_SP_END_DIR] ? 1 : 0) + (sp[TG_SP_END_VEL] ? 1 : 0) + (sp[TG_SP_END_ACC] ? 1 : 0)); r += L->n_eder;
    L->r_iwl = r; L->n_iwl = d * L->niw; r += L->n_iwl;
    L->r_iwv = r; L->n_iwv = sp[TG_SP_IW_VEL] ? d * L->niw : 0; r += L->n_iwv;
    L->meq = r;
    nl = L->n_sder + L->n_eder + L->n_iwl + L->n_iwv;
    L->r_db = r;
    L->n_db = (sp[TG_SP_DB_MINV] ? 1 : 0) + (sp[TG_SP_DB_MAXV] ? 1 : 0) +
              ((sp[TG_SP_DB_MAXV] && sp[TG_SP_DB_UP]) ? 1 : 0) + ((sp[TG_SP_DB_MAXV] && sp[TG_SP_DB_HORIZ]) ? 1 : 0) +
              (sp[TG_SP_DB_MAXA] ? 1 : 0) + (sp[TG_SP_DB_JERK] ? 1 : 0);
    r += L->n_db;
    L->n_tan = sp[TG_SP_TANG] ? 2 * L->nint : 0;
    L->r_tanl = r; r += L->n_tan;
    L->r_tanu = r; r += L->n_tan;
    L->r_turn = r; L->n_turn = sp[TG_SP_TURN] ? 1 : 0; r += L->n_turn;
    nl += L->n_db + 2 * L->n_tan + L->n_turn;
    int tot_int = 0;
    for (int c = 0; c < sp[TG_SP_NCORR]; c++) tot_int += sp[TG_SP_IPC0 + c];
    L->n_sfc = sp[TG_SP_NCORR] > 0 ? d * 4 * L->nint : 0;
    (void)tot_int;
    L->r_sfcl = r; r += L->n_sfc;
    L->r_sfcu = r; r += L->n_sfc;
    L->r_obs = r; L->n_obs = sp[TG_SP_NOBST]; r += L->n_obs;
    nl += L->n_obs;
    L->m = r; L->mineq = r - L->meq;
    L->m_nl = nl; L->m_lin = L->m - nl;

    L->turn_first = sp[TG_SP_START_KIND] == 1 ? 1 : 0;
    L->turn_ncp = N - L->turn_first - (sp[TG_SP_END_KIND] == 1 ? 1 : 0);

    int p = 0;
    L->p_start_loc = p; p += d;
    L->p_end_loc = p; p += d;
    L->p_target_vel = p; if (sp[TG_SP_END_KIND] == 2) p += d;
    L->p_sdir = p; if (sp[TG_SP_START_DIR]) p += d;
    L->p_svel = p; if (sp[TG_SP_START_VEL]) p += d;
    L->p_sacc = p; if (sp[TG_SP_START_ACC]) p += d;
    L->p_edir = p; if (sp[TG_SP_END_DIR]) p += d;
    L->p_evel = p; if (sp[TG_SP_END_VEL]) p += d;
    L->p_eacc = p; if (sp[TG_SP_END_ACC]) p += d;
    L->p_iwl = p; p += d * L->niw;
    L->p_iwv = p; if (sp[TG_SP_IW_VEL]) p += d * L->niw;
    L->p_minv = p; if (sp[TG_SP_DB_MINV]) p++;
    L->p_maxv = p; if (sp[TG_SP_DB_MAXV]) p++;
    L->p_up = p; if (sp[TG_SP_DB_UP]) p++;
    L->p_horiz = p; if (sp[TG_SP_DB_HORIZ]) p++;
    L->p_maxa = p; if (sp[TG_SP_DB_MAXA]) p++;
    L->p_grav = p; if (sp[TG_SP_DB_GRAV]) p++;
    L->p_jerk = p; if (sp[TG_SP_DB_JERK]) p++;
    L->p_tanmin = p; if (sp[TG_SP_TANG]) p++;
    L->p_tanmax = p; if (sp[TG_SP_TANG]) p++;
    L->p_turn = p; if (sp[TG_SP_TURN]) p++;
    L->p_sfc = p; p += sp[TG_SP_NCORR] * tg_sfc_stride(d);
    L->p_obs_c = p; p += d * sp[TG_SP_NOBST];
    L->p_obs_r = p; p += sp[TG_SP_NOBST];
    L->P = p;
}

// corridor that owns interval j (reference CF/sfc_constraints.py:53-77)
TG_HD int tg_corridor_of_interval(const int *sp, int j)
{
    int acc = 0;
    for (int c = 0; c < sp[TG_SP_NCORR]; c++) {
        acc += sp[TG_SP_IPC0 + c];
        if (j < acc) return c;
    }
    return sp[TG_SP_NCORR] - 1;
}

#endif  // TG_SPEC_H
