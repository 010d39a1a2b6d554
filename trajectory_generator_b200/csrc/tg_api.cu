// CUDA kernels (sm_100a) and the C-ABI declared in include/trajectory_generator_b200.h.
//
// Kernels: one warp per trajectory problem.
//   tg_eval_kernel   M1: objective, gradient, constraint rows, analytic nonlinear Jacobian rows
//   tg_linear_kernel constant Jacobian of the linear rows
//   tg_solve_kernel  M2: the whole SLSQP iteration of one problem per warp (tg_sqp.h),
//                    persistent CTAs pulling problem indices from an atomic queue
//   tg_legacy_*      single-problem kernels behind the reference's 24 C symbols
// Data layout in HBM: row-major [B][n] variables, [B][P] parameters, [B][m] rows,
// [B][m_nl][n] Jacobians -- a warp reads/writes its problem's rows as contiguous,
// coalesced 8-byte accesses; per-problem working sets are staged in shared memory.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <atomic>

#include "tg_sqp.h"
#include "../../include/trajectory_generator_b200.h"

// ---------------------------------------------------------------------------
struct TgShape {
    int sp[TG_SP_COUNT];
    TgLayout L;
};

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

static int tg_fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (e != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof g_err, "%s", what);
    return code;
}

#define TG_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) return tg_fail(100 + (int)e_, #call, e_);       \
    } while (0)

static int tg_make_shape(const int *spec, TgShape *S)
{
    if (!spec) return tg_fail(1, "spec is NULL");
    memcpy(S->sp, spec, sizeof S->sp);
    const int d = spec[TG_SP_DIM], N = spec[TG_SP_NCP];
    if (d != 2 && d != 3) return tg_fail(2, "dimension must be 2 or 3");
    if (N < 4 || N > 512) return tg_fail(2, "number of control points out of range");
    if (spec[TG_SP_NCORR] < 0 || spec[TG_SP_NCORR] > TG_MAX_CORRIDORS) return tg_fail(2, "too many corridors");
    tg_make_layout(S->sp, &S->L);
    return 0;
}

// ---------------------------------------------------------------------------
// M1 evaluation kernel
// ---------------------------------------------------------------------------
constexpr int EVAL_WARPS = 4;
constexpr int SOLVE_MIN_CTAS = 4;      // 4 CTAs x 4 warps per SM -> at most 128 registers per thread

__device__ __forceinline__ int tg_eval_smem_doubles(const TgLayout &L)
{
    return L.n + L.P + L.m + tg_scratch_doubles(L) + 4;
}

template <int D>
__global__ void __launch_bounds__(EVAL_WARPS * 32)
tg_eval_kernel(const TgShape S, int B, const double *__restrict__ par, const double *__restrict__ x,
               double *__restrict__ f, double *__restrict__ g, double *__restrict__ c, double *__restrict__ jnl)
{
    extern __shared__ double smem[];
    const TgLayout &L = S.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = tg_eval_smem_doubles(L);
    double *sx = smem + warp * per, *sp_ = sx + L.n, *sc = sp_ + L.P, *scr = sc + L.m;
    const int stride = gridDim.x * EVAL_WARPS;
    for (int b = blockIdx.x * EVAL_WARPS + warp; b < B; b += stride) {
        for (int i = lane; i < L.n; i += 32) sx[i] = x[(size_t)b * L.n + i];
        for (int i = lane; i < L.P; i += 32) sp_[i] = par[(size_t)b * L.P + i];
        __syncwarp();
        const double fv = tg_objective(L, S.sp, sx, g ? g + (size_t)b * L.n : nullptr);
        if (f && lane == 0) f[b] = fv;
        if (c || jnl) {
            TgJac sink = {jnl ? jnl + (size_t)b * L.m_nl * L.n : nullptr, L.n, 1, 1};
            tg_constraints_d<D>(L, S.sp, sp_, sx, sc, jnl ? &sink : nullptr, scr);
            if (c)
                for (int j = lane; j < L.m; j += 32) c[(size_t)b * L.m + j] = sc[j];
        }
        __syncwarp();
    }
}

template <int D>
__global__ void __launch_bounds__(EVAL_WARPS * 32)
tg_linear_kernel(const TgShape S, int B, const double *__restrict__ par, double *__restrict__ alin)
{
    extern __shared__ double smem[];
    const TgLayout &L = S.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sp_ = smem + warp * (L.P + 1);
    const int stride = gridDim.x * EVAL_WARPS;
    for (int b = blockIdx.x * EVAL_WARPS + warp; b < B; b += stride) {
        for (int i = lane; i < L.P; i += 32) sp_[i] = par[(size_t)b * L.P + i];
        __syncwarp();
        TgJac sink = {alin + (size_t)b * L.m * L.n, L.n, 1, 0};
        tg_linear_jacobian_d<D>(L, S.sp, sp_, sink);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// M2 solve kernel
// ---------------------------------------------------------------------------
// the reference's is_violation: violation flags of the LAST constraint in its list, tolerance 10e-6
// (TG/trajectory_generator.py:252-261, DS/constraint_function_data.py:12,45-48)
__device__ int tg_last_block_violation(const TgLayout &L, const double *c)
{
    int r0, r1, eq = 0;
    if (L.n_obs) { r0 = L.r_obs; r1 = r0 + L.n_obs; }
    else if (L.n_sfc) { r0 = L.r_sfcl; r1 = r0 + 2 * L.n_sfc; }
    else if (L.n_turn) { r0 = L.r_turn; r1 = r0 + 1; }
    else if (L.n_tan) { r0 = L.r_tanl; r1 = r0 + 2 * L.n_tan; }
    else if (L.n_db) { r0 = L.r_db; r1 = r0 + L.n_db; }
    else if (L.n_iwv) { r0 = L.r_iwv; r1 = r0 + L.n_iwv; eq = 1; }
    else if (L.n_iwl) { r0 = L.r_iwl; r1 = r0 + L.n_iwl; eq = 1; }
    else if (L.n_eder) { r0 = L.r_eder; r1 = r0 + L.n_eder; eq = 1; }
    else if (L.n_sder) { r0 = L.r_sder; r1 = r0 + L.n_sder; eq = 1; }
    else { r0 = L.r_end; r1 = r0 + L.n_end; eq = 1; }
    int bad = 0;
    for (int j = r0 + (threadIdx.x & 31); j < r1; j += 32) {
        const double v = c[j];
        if (eq ? (fabs(v) > 10e-6) : (v < -10e-6)) bad = 1;
        if (v != v) bad = 1;
    }
    return __any_sync(0xffffffffu, bad);
}

template <int D>
__global__ void __launch_bounds__(128, SOLVE_MIN_CTAS)
tg_solve_kernel(const TgShape S, int B, const double *__restrict__ par, double *__restrict__ x,
                                double *__restrict__ fout, int *__restrict__ status, int *__restrict__ nit,
                                int *__restrict__ violation, int maxiter, double ftol, int flags,
                                double *gws, size_t ws_doubles, int warps_per_cta, int *queue)
{
    extern __shared__ double smem[];
    const TgLayout &L = S.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp slice: parameters + (shared-memory workspace | pointer into the global one)
    double *spar = smem + (size_t)warp * (L.P + (gws ? 0 : ws_doubles) + 2);
    double *ws = gws ? gws + ((size_t)blockIdx.x * warps_per_cta + warp) * ws_doubles : spar + L.P + 1;
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(queue, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= B) break;
        for (int i = lane; i < L.P; i += 32) spar[i] = par[(size_t)b * L.P + i];
        __syncwarp();
        TgSqpResult res;
        tg_sqp_solve<D>(L, S.sp, spar, x + (size_t)b * L.n, ws, maxiter, ftol, flags, &res, nullptr, 0);
        res.status = __shfl_sync(0xffffffffu, res.status, 0);
        int viol = 0;
        if (res.status != 0) {
            TgSqpWs W;
            tg_sqp_carve(L, ws, &W);
            viol = tg_last_block_violation(L, W.c);
        }
        if (lane == 0) {
            if (status) status[b] = res.status;
            if (nit) nit[b] = res.nit;
            if (fout) fout[b] = res.f;
            if (violation) violation[b] = viol;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
static int g_sm_count = 0, g_smem_optin = 0;

extern "C" int tg_device_check(void)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return tg_fail(10, "no CUDA device available (this library has no CPU path)", e);
    int dev = 0;
    TG_CUDA(cudaGetDevice(&dev));
    cudaFuncAttributes attr;
    e = cudaFuncGetAttributes(&attr, tg_eval_kernel<2>);
    if (e != cudaSuccess) return tg_fail(11, "no kernel image for this device (built for sm_100a)", e);
    TG_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    TG_CUDA(cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return 0;
}

extern "C" int tg_spec_count(void) { return TG_SP_COUNT; }

extern "C" int tg_layout(const int *spec, int *out, int cap)
{
    const int cnt = (int)(sizeof(TgLayout) / sizeof(int));
    if (spec && out && cap >= cnt) {
        TgLayout L;
        tg_make_layout(spec, &L);
        memcpy(out, &L, sizeof L);
    }
    return cnt;
}

extern "C" const char *tg_last_error(void) { return g_err; }
extern "C" unsigned long long tg_launch_count(void) { return g_launches.load(); }

extern "C" int tg_eval_batch(const int *spec, int B, const double *par, const double *x, double *f, double *g,
                             double *c, double *jnl, void *stream)
{
    TgShape S;
    int rc = tg_make_shape(spec, &S);
    if (rc) return rc;
    if (B <= 0) return 0;
    if ((rc = tg_device_check())) return rc;
    const size_t smem = (size_t)EVAL_WARPS * (S.L.n + S.L.P + S.L.m + tg_scratch_doubles(S.L) + 4) * sizeof(double);
    if (smem > (size_t)g_smem_optin) return tg_fail(3, "problem shape too large for the evaluation kernel's shared memory");
    if (S.L.d == 2) TG_CUDA(cudaFuncSetAttribute(tg_eval_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else TG_CUDA(cudaFuncSetAttribute(tg_eval_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (B + EVAL_WARPS - 1) / EVAL_WARPS;
    const int cap = g_sm_count * 16;
    if (grid > cap) grid = cap;
    if (S.L.d == 2) tg_eval_kernel<2><<<grid, EVAL_WARPS * 32, smem, (cudaStream_t)stream>>>(S, B, par, x, f, g, c, jnl);
    else tg_eval_kernel<3><<<grid, EVAL_WARPS * 32, smem, (cudaStream_t)stream>>>(S, B, par, x, f, g, c, jnl);
    g_launches++;
    TG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tg_linear_rows_batch(const int *spec, int B, const double *par, double *alin, void *stream)
{
    TgShape S;
    int rc = tg_make_shape(spec, &S);
    if (rc) return rc;
    if (B <= 0) return 0;
    if ((rc = tg_device_check())) return rc;
    const size_t smem = (size_t)EVAL_WARPS * (S.L.P + 1) * sizeof(double);
    int grid = (B + EVAL_WARPS - 1) / EVAL_WARPS;
    const int cap = g_sm_count * 16;
    if (grid > cap) grid = cap;
    if (S.L.d == 2) tg_linear_kernel<2><<<grid, EVAL_WARPS * 32, smem, (cudaStream_t)stream>>>(S, B, par, alin);
    else tg_linear_kernel<3><<<grid, EVAL_WARPS * 32, smem, (cudaStream_t)stream>>>(S, B, par, alin);
    g_launches++;
    TG_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// M2, lock-step form: the two stages of tg_sqp.h as separate kernels over the whole batch.  Every warp of
// the machine then runs the same few functions at the same time (the fused kernel is bound by
// instruction-cache misses: each warp sits in a different phase of a ~300 KB program).  Per-problem state
// lives in a global workspace and is staged through shared memory inside a stage when it fits.
// ---------------------------------------------------------------------------
constexpr int STAGE_WARPS = 4;

template <int D>
__global__ void __launch_bounds__(STAGE_WARPS * 32)
tg_sqp_begin_kernel(const TgShape S, int B, const double *__restrict__ x, double *pws, size_t np, int maxiter, double ftol,
                    int flags)
{
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * STAGE_WARPS + warp;
    if (b >= B) return;
    TgSqpWs W;
    size_t a, c;
    tg_sqp_carve2(S.L, pws + (size_t)b * np, nullptr, &W, &a, &c);
    tg_sqp_begin(S.L, W, x + (size_t)b * S.L.n, maxiter, ftol, flags);
}

// STAGE 0: line search / first evaluation.  STAGE 1: update + QP.
template <int D, int STAGE>
__global__ void __launch_bounds__(STAGE_WARPS * 32, 4)
tg_sqp_stage_kernel(const TgShape S, int B, const double *__restrict__ par, double *pws, size_t np, size_t ns, int staged,
                    int *counters)
{
    extern __shared__ double smem[];
    const TgLayout &L = S.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * STAGE_WARPS + warp;
    if (b >= B) return;
    double *gp = pws + (size_t)b * np;
    const int st = ((const TgSqpCtl *)gp)->state;
    if (STAGE == 0 ? !(st == TG_ST_INIT || st == TG_ST_LS) : !(st == TG_ST_UPDATE || st == TG_ST_QP)) return;
    const size_t per = (size_t)L.P + 1 + ns + (staged ? np : 0);
    double *spar = smem + (size_t)warp * per, *sscr = spar + L.P + 1, *spers = sscr + ns;
    if (STAGE == 0)
        for (int i = lane; i < L.P; i += 32) spar[i] = par[(size_t)b * L.P + i];
    if (staged)
        for (size_t i = lane; i < np; i += 32) spers[i] = gp[i];
    __syncwarp();
    TgSqpWs W;
    size_t a, c;
    tg_sqp_carve2(L, staged ? spers : gp, sscr, &W, &a, &c);
    if (STAGE == 0) tg_sqp_stage_ls<D>(L, S.sp, spar, W, nullptr, 0);
    else tg_sqp_stage_qp(L, W);
    __syncwarp();
    if (STAGE == 1 && lane == 0 && W.ctl->state == TG_ST_DONE) atomicAdd(counters, 1);
    if (staged)
        for (size_t i = lane; i < np; i += 32) gp[i] = spers[i];
}

__global__ void tg_sqp_finish_kernel(const TgShape S, int B, const double *pws, size_t np, double *__restrict__ x,
                                     double *__restrict__ fout, int *__restrict__ status, int *__restrict__ nit,
                                     int *__restrict__ violation)
{
    const TgLayout &L = S.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * STAGE_WARPS + warp;
    if (b >= B) return;
    TgSqpWs W;
    size_t a, c;
    tg_sqp_carve2(L, const_cast<double *>(pws) + (size_t)b * np, nullptr, &W, &a, &c);
    for (int i = lane; i < L.n; i += 32) x[(size_t)b * L.n + i] = W.x[i];
    const TgSqpCtl ctl = *W.ctl;
    int viol = 0;
    if (ctl.status != 0) viol = tg_last_block_violation(L, W.c);
    if (lane == 0) {
        if (status) status[b] = ctl.status;
        if (nit) nit[b] = ctl.iter > ctl.maxiter ? ctl.maxiter : ctl.iter;
        if (fout) fout[b] = ctl.f;
        if (violation) violation[b] = viol;
    }
}

// launch geometry for a shape
struct TgSolvePlan {
    // fused kernel
    int warps_per_cta, ctas, use_global;
    size_t ws_doubles, smem_bytes, global_bytes;
    // lock-step kernels
    size_t np, ns, stage_smem;
    int staged, chunk;
    size_t phased_bytes;
};

#define TG_PHASED_CHUNK_BYTES ((size_t)6 << 30)     // per-chunk cap of the global state (problems are solved in chunks)

static int tg_plan_solve(const TgShape &S, int B, TgSolvePlan *P)
{
    int rc = tg_device_check();
    if (rc) return rc;
    P->ws_doubles = tg_sqp_workspace_doubles(S.L);
    const size_t per_warp_shared = (S.L.P + P->ws_doubles + 2) * sizeof(double);
    const size_t budget = (size_t)g_smem_optin - 1024;
    // shared-memory workspace when at least 8 warps fit on an SM; otherwise the workspace lives in
    // global memory (L1/L2 resident) and only the parameter row is staged
    const size_t sm_total = 227 * 1024;
    if (per_warp_shared * 8 <= sm_total) {
        P->use_global = 0;
        P->warps_per_cta = 4;
        while (per_warp_shared * P->warps_per_cta > budget) P->warps_per_cta >>= 1;
        P->smem_bytes = per_warp_shared * P->warps_per_cta;
        int per_sm = (int)(sm_total / (P->smem_bytes + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm * P->warps_per_cta > 32) per_sm = 32 / P->warps_per_cta;
        P->ctas = g_sm_count * per_sm;
        P->global_bytes = 0;
    } else {
        P->use_global = 1;
        P->warps_per_cta = 4;
        P->smem_bytes = (size_t)P->warps_per_cta * (S.L.P + 2) * sizeof(double);
        P->ctas = g_sm_count * 4;     // 16 warps per SM
        P->global_bytes = (size_t)P->ctas * P->warps_per_cta * P->ws_doubles * sizeof(double);
    }
    const int need = (B + P->warps_per_cta - 1) / P->warps_per_cta;
    if (P->ctas > need) {
        P->ctas = need > 0 ? need : 1;
        if (P->use_global) P->global_bytes = (size_t)P->ctas * P->warps_per_cta * P->ws_doubles * sizeof(double);
    }
    // lock-step plan: stage the persistent state through shared memory when 16 warps still fit on an SM
    P->np = tg_sqp_persistent_doubles(S.L);
    P->ns = tg_sqp_scratch_doubles(S.L);
    const size_t with_state = ((size_t)S.L.P + 1 + P->ns + P->np) * sizeof(double) * STAGE_WARPS;
    const size_t without = ((size_t)S.L.P + 1 + P->ns) * sizeof(double) * STAGE_WARPS;
    P->staged = with_state * 4 <= sm_total - 4096;
    P->stage_smem = P->staged ? with_state : without;
    if (P->stage_smem > budget) return tg_fail(3, "problem shape too large for the solve kernels' shared memory");
    size_t chunk = TG_PHASED_CHUNK_BYTES / (P->np * sizeof(double));
    if (chunk < 1024) chunk = 1024;
    if (chunk > (size_t)B) chunk = (size_t)B;
    P->chunk = (int)chunk;
    P->phased_bytes = chunk * P->np * sizeof(double);
    return 0;
}

extern "C" size_t tg_solve_workspace_bytes(const int *spec, int B)
{
    TgShape S;
    TgSolvePlan P;
    if (tg_make_shape(spec, &S) || tg_plan_solve(S, B, &P)) return 0;
    const size_t need = P.global_bytes > P.phased_bytes ? P.global_bytes : P.phased_bytes;
    return need + 256;     // + counters
}

template <int D>
static int tg_solve_fused(const TgShape &S, const TgSolvePlan &P, int B, const double *par, double *x, double *f,
                          int *status, int *nit, int *violation, int maxiter, double ftol, int flags, int *queue,
                          double *gws, cudaStream_t st)
{
    TG_CUDA(cudaFuncSetAttribute(tg_solve_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes));
    tg_solve_kernel<D><<<P.ctas, P.warps_per_cta * 32, P.smem_bytes, st>>>(S, B, par, x, f, status, nit, violation, maxiter,
                                                                           ftol, flags, P.use_global ? gws : nullptr,
                                                                           P.ws_doubles, P.warps_per_cta, queue);
    g_launches++;
    TG_CUDA(cudaGetLastError());
    return 0;
}

template <int D>
static int tg_solve_phased(const TgShape &S, const TgSolvePlan &P, int B, const double *par, double *x, double *f,
                           int *status, int *nit, int *violation, int maxiter, double ftol, int flags, int *counters,
                           double *pws, cudaStream_t st)
{
    TG_CUDA(cudaFuncSetAttribute(tg_sqp_stage_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.stage_smem));
    TG_CUDA(cudaFuncSetAttribute(tg_sqp_stage_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.stage_smem));
    const TgLayout &L = S.L;
    for (int lo = 0; lo < B; lo += P.chunk) {
        const int nb = B - lo < P.chunk ? B - lo : P.chunk;
        const int grid = (nb + STAGE_WARPS - 1) / STAGE_WARPS;
        const double *cpar = par + (size_t)lo * L.P;
        double *cx = x + (size_t)lo * L.n;
        TG_CUDA(cudaMemsetAsync(counters, 0, 256, st));
        tg_sqp_begin_kernel<D><<<grid, STAGE_WARPS * 32, 0, st>>>(S, nb, cx, pws, P.np, maxiter, ftol, flags);
        g_launches++;
        int done = 0;
        // each round = one SLSQP major iteration of every unfinished problem; maxiter + 1 rounds finish everything
        for (int round = 0; round <= maxiter + 1 && done < nb; round++) {
            tg_sqp_stage_kernel<D, 0><<<grid, STAGE_WARPS * 32, P.stage_smem, st>>>(S, nb, cpar, pws, P.np, P.ns, P.staged, counters);
            tg_sqp_stage_kernel<D, 1><<<grid, STAGE_WARPS * 32, P.stage_smem, st>>>(S, nb, cpar, pws, P.np, P.ns, P.staged, counters);
            g_launches += 2;
            if ((round & 7) == 7) {      // poll the number of finished problems
                TG_CUDA(cudaMemcpyAsync(&done, counters, sizeof(int), cudaMemcpyDeviceToHost, st));
                TG_CUDA(cudaStreamSynchronize(st));
            }
        }
        tg_sqp_finish_kernel<<<grid, STAGE_WARPS * 32, 0, st>>>(S, nb, pws, P.np, cx, f ? f + lo : nullptr,
                                                               status ? status + lo : nullptr, nit ? nit + lo : nullptr,
                                                               violation ? violation + lo : nullptr);
        g_launches++;
        TG_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int tg_solve_batch(const int *spec, int B, const double *par, double *x, double *f, int *status, int *nit,
                              int *violation, int maxiter, double ftol, int flags, void *workspace,
                              size_t workspace_bytes, void *stream)
{
    TgShape S;
    TgSolvePlan P;
    int rc = tg_make_shape(spec, &S);
    if (rc) return rc;
    if (B <= 0) return 0;
    if ((rc = tg_plan_solve(S, B, &P))) return rc;
    if (S.L.n > 62) return tg_fail(3, "more than 62 optimisation variables are not supported by the solve kernel");
    const size_t need = (P.global_bytes > P.phased_bytes ? P.global_bytes : P.phased_bytes) + 256;
    if (!workspace || workspace_bytes < need) return tg_fail(4, "workspace too small (see tg_solve_workspace_bytes)");
    int *counters = (int *)workspace;
    double *gws = (double *)((char *)workspace + 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (flags & TG_SOLVE_FUSED) {
        TG_CUDA(cudaMemsetAsync(counters, 0, 256, st));
        return S.L.d == 2 ? tg_solve_fused<2>(S, P, B, par, x, f, status, nit, violation, maxiter, ftol, flags, counters, gws, st)
                          : tg_solve_fused<3>(S, P, B, par, x, f, status, nit, violation, maxiter, ftol, flags, counters, gws, st);
    }
    return S.L.d == 2 ? tg_solve_phased<2>(S, P, B, par, x, f, status, nit, violation, maxiter, ftol, flags, counters, gws, st)
                      : tg_solve_phased<3>(S, P, B, par, x, f, status, nit, violation, maxiter, ftol, flags, counters, gws, st);
}

// ---------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return tg_fail(100 + (int)e, "cudaMalloc", e);
        cap = bytes;
        return 0;
    }
};
static std::mutex g_host_mutex;
static DevBuf g_par, g_x, g_f, g_g, g_c, g_j, g_i, g_ws;

extern "C" int tg_eval_host(const int *spec, int B, const double *par, const double *x, double *f, double *g, double *c,
                            double *jnl)
{
    TgShape S;
    int rc = tg_make_shape(spec, &S);
    if (rc) return rc;
    if (B <= 0) return 0;
    if ((rc = tg_device_check())) return rc;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    const TgLayout &L = S.L;
    const size_t nb = sizeof(double);
    if ((rc = g_par.ensure((size_t)B * (L.P + 1) * nb)) || (rc = g_x.ensure((size_t)B * L.n * nb))) return rc;
    if (f && (rc = g_f.ensure((size_t)B * nb))) return rc;
    if (g && (rc = g_g.ensure((size_t)B * L.n * nb))) return rc;
    if (c && (rc = g_c.ensure((size_t)B * (L.m + 1) * nb))) return rc;
    if (jnl && (rc = g_j.ensure((size_t)B * (L.m_nl * L.n + 1) * nb))) return rc;
    TG_CUDA(cudaMemcpyAsync(g_par.p, par, (size_t)B * L.P * nb, cudaMemcpyHostToDevice, 0));
    TG_CUDA(cudaMemcpyAsync(g_x.p, x, (size_t)B * L.n * nb, cudaMemcpyHostToDevice, 0));
    rc = tg_eval_batch(spec, B, (const double *)g_par.p, (const double *)g_x.p, f ? (double *)g_f.p : nullptr,
                       g ? (double *)g_g.p : nullptr, c ? (double *)g_c.p : nullptr, jnl ? (double *)g_j.p : nullptr, 0);
    if (rc) return rc;
    if (f) TG_CUDA(cudaMemcpyAsync(f, g_f.p, (size_t)B * nb, cudaMemcpyDeviceToHost, 0));
    if (g) TG_CUDA(cudaMemcpyAsync(g, g_g.p, (size_t)B * L.n * nb, cudaMemcpyDeviceToHost, 0));
    if (c) TG_CUDA(cudaMemcpyAsync(c, g_c.p, (size_t)B * L.m * nb, cudaMemcpyDeviceToHost, 0));
    if (jnl) TG_CUDA(cudaMemcpyAsync(jnl, g_j.p, (size_t)B * L.m_nl * L.n * nb, cudaMemcpyDeviceToHost, 0));
    TG_CUDA(cudaStreamSynchronize(0));
    return 0;
}

extern "C" int tg_solve_host(const int *spec, int B, const double *par, double *x, double *f, int *status, int *nit,
                             int *violation, int maxiter, double ftol, int flags)
{
    TgShape S;
    int rc = tg_make_shape(spec, &S);
    if (rc) return rc;
    if (B <= 0) return 0;
    if ((rc = tg_device_check())) return rc;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    const TgLayout &L = S.L;
    const size_t nb = sizeof(double);
    const size_t wsb = tg_solve_workspace_bytes(spec, B);
    if (wsb == 0) return tg_fail(5, g_err[0] ? g_err : "cannot plan the solve kernel");
    if ((rc = g_par.ensure((size_t)B * (L.P + 1) * nb)) || (rc = g_x.ensure((size_t)B * L.n * nb)) ||
        (rc = g_f.ensure((size_t)B * nb)) || (rc = g_i.ensure((size_t)B * 3 * sizeof(int))) || (rc = g_ws.ensure(wsb)))
        return rc;
    int *di = (int *)g_i.p;
    TG_CUDA(cudaMemcpyAsync(g_par.p, par, (size_t)B * L.P * nb, cudaMemcpyHostToDevice, 0));
    TG_CUDA(cudaMemcpyAsync(g_x.p, x, (size_t)B * L.n * nb, cudaMemcpyHostToDevice, 0));
    rc = tg_solve_batch(spec, B, (const double *)g_par.p, (double *)g_x.p, (double *)g_f.p, di, di + B, di + 2 * B,
                        maxiter, ftol, flags, g_ws.p, wsb, 0);
    if (rc) return rc;
    TG_CUDA(cudaMemcpyAsync(x, g_x.p, (size_t)B * L.n * nb, cudaMemcpyDeviceToHost, 0));
    if (f) TG_CUDA(cudaMemcpyAsync(f, g_f.p, (size_t)B * nb, cudaMemcpyDeviceToHost, 0));
    if (status) TG_CUDA(cudaMemcpyAsync(status, di, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, 0));
    if (nit) TG_CUDA(cudaMemcpyAsync(nit, di + B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, 0));
    if (violation) TG_CUDA(cudaMemcpyAsync(violation, di + 2 * B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, 0));
    TG_CUDA(cudaStreamSynchronize(0));
    return 0;
}

// ---------------------------------------------------------------------------
// the reference's 24 symbols: single-problem launches (one warp)
// ---------------------------------------------------------------------------
enum { LG_TURN = 0, LG_MINV, LG_OBST, LG_INTERVALS, LG_BEZ };

// MDM min-norm point of the hull of 3 points (CC/src/MDMAlgorithmClass.cpp:11-71), one thread
template <int D>
__device__ double tg_mdm_min_norm3(const double *pts /* pts[c*3+i] */, int max_iterations, double tolerance)
{
    const int npts = 3;
    double p[3] = {1, 0, 0}, cur[D];
    int supp[3] = {0, 0, 0}, nsupp = 1, iterations = 0;
    double delta_p = 1.0;
    for (int c = 0; c < D; c++) cur[c] = pts[c * npts];
    while (delta_p > 0.000001 && iterations < max_iterations && nsupp > 0) {
        int max_index = supp[0], min_index = 0;
        double best = DBL_MIN;     // the reference starts from numeric_limits<double>::min()
        for (int i = 0; i < nsupp; i++) {
            double s = 0;
            for (int c = 0; c < D; c++) s += pts[c * npts + supp[i]] * cur[c];
            if (s > best) { best = s; max_index = supp[i]; }
        }
        double lo = DBL_MAX;
        for (int i = 0; i < npts; i++) {
            double s = 0;
            for (int c = 0; c < D; c++) s += pts[c * npts + i] * cur[c];
            if (s < lo) { lo = s; min_index = i; }
        }
        double diff[D], dn2 = 0;
        delta_p = 0;
        for (int c = 0; c < D; c++) {
            diff[c] = pts[c * npts + max_index] - pts[c * npts + min_index];
            delta_p += diff[c] * cur[c];
            dn2 += diff[c] * diff[c];
        }
        if (delta_p > tolerance) {
            const double dn = sqrt(dn2);
            double t = delta_p / (p[max_index] * dn * dn);
            if (t >= 1) t = 1.0;
            for (int c = 0; c < D; c++) cur[c] -= t * p[max_index] * diff[c];
            const double t1 = t * p[max_index], t2 = 1 - t;
            p[min_index] += t1;
            p[max_index] *= t2;
            nsupp = 0;
            for (int i = 0; i < npts; i++)
                if (p[i] > tolerance) supp[nsupp++] = i;
            iterations++;
        }
    }
    double s = 0;
    for (int c = 0; c < D; c++) s += cur[c] * cur[c];
    return sqrt(s);
}

template <int D>
__global__ void tg_legacy_kernel(int what, const double *pts, int N, double alpha, int kind, const double *centers,
                                 const double *radii, int K, double *out)
{
    extern __shared__ double sx[];
    const int lane = threadIdx.x;
    for (int i = lane; i < D * N; i += 32) sx[i] = pts[i];
    __syncwarp();
    const int nint = N - 3;
    if (what == LG_TURN) {
        // CC/src/CrossTermBounds.cpp:13-61
        double best = 0; int jb = 0x7fffffff;
        for (int j = lane; j < nint; j += 32) {
            TgInterval<D> I;
            tg_load_interval<D>(sx, N, j, I);
            const double b = tg_interval_turn_bound<D>(I, alpha, kind, nullptr);
            if (b > best) { best = b; jb = j; }
        }
        tg_wargmax(best, jb);
        if (lane == 0) out[0] = best;
    } else if (what == LG_MINV) {
        // CC/src/DerivativeBounds.cpp:12-27
        double best = DBL_MAX; int jb = 0x7fffffff;
        for (int j = lane; j < nint; j += 32) {
            TgInterval<D> I; double v, t;
            tg_load_interval<D>(sx, N, j, I);
            tg_min_velocity<D>(I, alpha, v, t);
            if (v < best) { best = v; jb = j; }
        }
        tg_wargmin(best, jb);
        if (lane == 0) out[0] = best;
    } else if (what == LG_OBST) {
        // CC/src/SphereCollisionEvaluator.cpp:13-45: one lane per sphere
        for (int i = lane; i < K; i += 32) {
            double ctr[D], best = DBL_MAX;
            for (int c = 0; c < D; c++) ctr[c] = centers[c * K + i];
            for (int j = 0; j < nint; j++) {
                const double dist = tg_hull_distance<D>(sx, N, j, ctr, radii[i], nullptr);
                if (best > dist) best = dist;
            }
            out[i] = best;
        }
    } else if (what == LG_INTERVALS) {
        // CC/src/SphereCollisionEvaluator.cpp:70-86: one lane per interval, sphere = (centers[0..D), radii[0])
        double ctr[D];
        for (int c = 0; c < D; c++) ctr[c] = centers[c];
        for (int j = lane; j < nint; j += 32) out[j] = tg_hull_distance<D>(sx, N, j, ctr, radii[0], nullptr);
    } else {
        // CC/src/ControlPointDerivativeBounds.cpp:13-46: N Bezier velocity points, triples at stride 2
        const int nseg = (N - 1) / 2;
        double best = DBL_MAX; int ib = 0x7fffffff;
        for (int i = lane; i < nseg; i += 32) {
            double tri[D * 3];
            for (int c = 0; c < D; c++)
                for (int l = 0; l < 3; l++) tri[c * 3 + l] = sx[c * N + 2 * i + l];
            const double v = tg_mdm_min_norm3<D>(tri, 500, 0.000001);
            if (v < best) { best = v; ib = i; }
        }
        tg_wargmin(best, ib);
        if (lane == 0) out[0] = best;
    }
}

struct LegacyHandle {
    int D;
    std::mutex mu;
    DevBuf din, dout;
    double *host_out = nullptr;
    size_t host_cap = 0;
};

static void *tg_new_handle(int D)
{
    LegacyHandle *h = new LegacyHandle();
    h->D = D;
    return h;
}

// runs one legacy query; returns pointer to nout doubles owned by the handle (NaN-filled on failure)
static double *tg_legacy_run(void *obj, int D, int what, const double *pts, int N, double alpha, int kind,
                             const double *centers, const double *radii, int K, int nout)
{
    static LegacyHandle fallback[2];
    LegacyHandle *h = obj ? (LegacyHandle *)obj : &fallback[D - 2];
    std::lock_guard<std::mutex> lock(h->mu);
    if ((size_t)nout + 1 > h->host_cap) {
        // the previous buffer is intentionally not freed: the reference hands out a fresh
        // `new double[]` per call and its callers may still hold the old view
        h->host_cap = (size_t)nout + 64;
        h->host_out = new double[h->host_cap];
    }
    double *res = h->host_out;
    for (int i = 0; i < nout; i++) res[i] = NAN;
    if (tg_device_check()) { fprintf(stderr, "libTrajectoryConstraints (B200): %s\n", g_err); return res; }
    if (N < 4 && what != LG_BEZ) { fprintf(stderr, "libTrajectoryConstraints (B200): need at least 4 control points\n"); return res; }
    const int nc = (what == LG_INTERVALS) ? D : D * K;
    const int nr = (what == LG_INTERVALS) ? 1 : K;
    const size_t nin = (size_t)D * N + nc + nr;
    if (h->din.ensure((nin + 1) * sizeof(double)) || h->dout.ensure(((size_t)nout + 1) * sizeof(double))) {
        fprintf(stderr, "libTrajectoryConstraints (B200): %s\n", g_err);
        return res;
    }
    double *dpts = (double *)h->din.p, *dctr = dpts + (size_t)D * N, *drad = dctr + nc;
    cudaMemcpyAsync(dpts, pts, sizeof(double) * D * N, cudaMemcpyHostToDevice, 0);
    if (centers && nc) cudaMemcpyAsync(dctr, centers, sizeof(double) * nc, cudaMemcpyHostToDevice, 0);
    if (radii && nr) cudaMemcpyAsync(drad, radii, sizeof(double) * nr, cudaMemcpyHostToDevice, 0);
    const size_t smem = sizeof(double) * ((size_t)D * N + 1);
    if (D == 2) tg_legacy_kernel<2><<<1, 32, smem, 0>>>(what, dpts, N, alpha, kind, dctr, drad, K, (double *)h->dout.p);
    else tg_legacy_kernel<3><<<1, 32, smem, 0>>>(what, dpts, N, alpha, kind, dctr, drad, K, (double *)h->dout.p);
    g_launches++;
    cudaError_t e = cudaMemcpy(res, h->dout.p, sizeof(double) * nout, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        fprintf(stderr, "libTrajectoryConstraints (B200): %s\n", cudaGetErrorString(e));
        for (int i = 0; i < nout; i++) res[i] = NAN;
    }
    return res;
}

#define TG_DEFINE_LEGACY(D)                                                                                            \
    extern "C" void *CrossTermBounds_##D(void) { return tg_new_handle(D); }                                            \
    extern "C" double get_spline_curvature_bound_##D(void *o, double p[], int N)                                       \
    { return tg_legacy_run(o, D, LG_TURN, p, N, 1.0, TG_TURN_CURVATURE, 0, 0, 0, 1)[0]; }                              \
    extern "C" double get_spline_angular_rate_bound_##D(void *o, double p[], int N, double a)                          \
    { return tg_legacy_run(o, D, LG_TURN, p, N, a, TG_TURN_ANGULAR_RATE, 0, 0, 0, 1)[0]; }                             \
    extern "C" double get_spline_centripetal_acceleration_bound_##D(void *o, double p[], int N, double a)              \
    { return tg_legacy_run(o, D, LG_TURN, p, N, a, TG_TURN_CENTRIPETAL, 0, 0, 0, 1)[0]; }                              \
    extern "C" void *DerivativeBounds_##D(void) { return tg_new_handle(D); }                                           \
    extern "C" double find_min_velocity_of_spline_##D(void *o, double p[], int N, double a)                            \
    { return tg_legacy_run(o, D, LG_MINV, p, N, a, 0, 0, 0, 0, 1)[0]; }                                                \
    extern "C" void *ObstacleConstraints_##D(void) { return tg_new_handle(D); }                                        \
    extern "C" double *getObstaclesConstraintsForSpline_##D(void *o, double c[], double r[], int K, double p[], int N) \
    { return tg_legacy_run(o, D, LG_OBST, p, N, 1.0, 0, c, r, K, K > 0 ? K : 1); }                                     \
    extern "C" double *getObstacleConstraintsForIntervals_##D(void *o, double p[], int N, double r, double c[])        \
    { return tg_legacy_run(o, D, LG_INTERVALS, p, N, 1.0, 0, c, &r, 1, N - 3 > 0 ? N - 3 : 1); }                       \
    extern "C" double getObstacleConstraintForSpline_##D(void *o, double p[], int N, double r, double c[])             \
    {                                                                                                                  \
        double *a = tg_legacy_run(o, D, LG_INTERVALS, p, N, 1.0, 0, c, &r, 1, N - 3 > 0 ? N - 3 : 1);                  \
        double b = DBL_MAX;                                                                                            \
        for (int j = 0; j < N - 3; j++) { if (a[j] != a[j]) return a[j]; if (b > a[j]) b = a[j]; }                     \
        return b;                                                                                                      \
    }                                                                                                                  \
    extern "C" void *ControlPointDerivativeBounds_##D(void) { return tg_new_handle(D); }                               \
    extern "C" double find_min_velocity_of_bez_vel_cont_pts_##D(void *o, double p[], int n)                            \
    { return tg_legacy_run(o, D, LG_BEZ, p, n, 1.0, 0, 0, 0, 0, 1)[0]; }

TG_DEFINE_LEGACY(2)
TG_DEFINE_LEGACY(3)
