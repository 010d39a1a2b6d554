// M2, fused form: one persistent kernel, each warp runs both SQP stages of a problem until it is done.  Kept for
// comparison with the lock-step stage kernels (profiles/README.md: it is instruction-cache bound); the QP-stage
// functions stay out of line here.
#define TG_GS 32
#define TG_SQP_NOINLINE
#include "tg_sqp.h"
#include "tg_shape.h"

// ---------------------------------------------------------------------------
// fused form: persistent CTAs, each warp pulls problem indices from an atomic queue and runs both stages
// until its problem is done (kept for comparison; see profiles/README.md)
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128, 4)
tg_solve_kernel(const TgShape S, int B, const double *__restrict__ par, double *__restrict__ x, double *__restrict__ fout,
                int *__restrict__ status, int *__restrict__ nit, int *__restrict__ violation, int maxiter, double ftol,
                int flags, double *gws, size_t ws_doubles, int warps_per_cta, int *queue)
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

cudaError_t tg_launch_fused_g32(const TgShape &S, int B, const double *par, double *x, double *f, int *status, int *nit,
                                int *violation, int maxiter, double ftol, int flags, double *gws, size_t ws_doubles,
                                int warps_per_cta, int ctas, size_t smem, int *queue, cudaStream_t st)
{
    cudaError_t e;
    if (S.L.d == 2) {
        if ((e = tg_allow_shared_memory(tg_solve_kernel<2>))) return e;
        tg_solve_kernel<2><<<ctas, warps_per_cta * 32, smem, st>>>(S, B, par, x, f, status, nit, violation, maxiter, ftol,
                                                                    flags, gws, ws_doubles, warps_per_cta, queue);
    } else {
        if ((e = tg_allow_shared_memory(tg_solve_kernel<3>))) return e;
        tg_solve_kernel<3><<<ctas, warps_per_cta * 32, smem, st>>>(S, B, par, x, f, status, nit, violation, maxiter, ftol,
                                                                    flags, gws, ws_doubles, warps_per_cta, queue);
    }
    return cudaGetLastError();
}
