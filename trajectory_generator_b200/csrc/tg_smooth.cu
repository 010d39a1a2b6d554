// Spline order converter on the GPU (TG/spline_order_converter.py): one warp per old spline runs
// the whole SLSQP solve of tg_smooth.h; workspaces live in global memory (one per resident warp), the shape-wide
// sample table is built by a small kernel first.
#define TG_GS 32
#define TG_SQP_NOINLINE
#include <cuda_runtime.h>
#include "tg_smooth.h"
#include "../../include/trajectory_generator_b200.h"

void tg_note_launch(int count);      // tg_api.cu

#define TG_SMOOTH_MAX_VARIABLES 160      // one warp per spline, workspace in global memory: a practical bound, not a structural one

namespace {

__global__ void tg_smooth_table_kernel(const TgSmoothShape S, double *tab)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < S.R) tg_smooth_table_entry(S, t, tab);
    else if (t < S.R + 6) tg_smooth_end_entry(S, t - S.R, tab);
}

__global__ void __launch_bounds__(128) tg_smooth_kernel(const TgSmoothShape S, int B, const double *tab, const double *par,
                                                        double *x, double *f, int *status, int *nit, double *ws,
                                                        size_t ws_doubles, int maxiter, double acc)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int P = tg_smooth_par_doubles(S), n = S.d * S.N;
    for (int b = warp; b < B; b += nwarps) {
        TgSqpResult res;
        tg_smooth_solve(S, tab, par + (size_t)b * P, x + (size_t)b * n, ws + (size_t)warp * ws_doubles, maxiter, acc, &res);
        if (TG_LANE() == 0) {
            if (f) f[b] = res.f;
            if (status) status[b] = res.status;
            if (nit) nit[b] = res.nit;
        }
        __syncwarp();
    }
}

__global__ void tg_smooth_initial_kernel(int d, int oldN, int N, int B, const double *old_cps, double *x0, double *scr)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) tg_smooth_initial_points(d, old_cps + (size_t)b * d * oldN, oldN, N, x0 + (size_t)b * d * N, scr + (size_t)b * oldN);
}

int tg_smooth_check(int d, int N, int order, int resolution)
{
    if ((d != 2 && d != 3) || order < 2 || order > TG_SMOOTH_MAX_ORDER || N < order + 1 || resolution < 2) return 2;
    if (d * N > TG_SMOOTH_MAX_VARIABLES) return 3;
    return 0;
}

size_t tg_smooth_warps(int B, int sms) { const size_t w = (size_t)sms * 16; return (size_t)B < w ? (size_t)B : w; }

}  // namespace

extern "C" size_t tg_smooth_workspace_bytes(int d, int N, int order, int resolution, int B)
{
    if (tg_smooth_check(d, N, order, resolution) || B <= 0) return 0;
    const TgSmoothShape S = {d, N, order, resolution, 1.0};
    TgLayout L;
    tg_smooth_layout(S, &L);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return ((size_t)tg_smooth_table_doubles(S) + 8 + tg_smooth_warps(B, sms) * (tg_sqp_workspace_doubles(L) + 8)) * sizeof(double);
}

extern "C" int tg_smooth_batch(int d, int N, int order, int resolution, double scale, int B, const double *par, double *x,
                               double *f, int *status, int *nit, int maxiter, double ftol, void *workspace,
                               size_t workspace_bytes, void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if ((rc = tg_smooth_check(d, N, order, resolution))) return rc;
    if (!par || !x || !workspace || workspace_bytes < tg_smooth_workspace_bytes(d, N, order, resolution, B)) return 4;
    const TgSmoothShape S = {d, N, order, resolution, scale};
    TgLayout L;
    tg_smooth_layout(S, &L);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = (cudaStream_t)stream;
    double *tab = (double *)workspace;
    double *ws = tab + ((tg_smooth_table_doubles(S) + 8) & ~7);
    const size_t wsd = tg_sqp_workspace_doubles(L) + 8;
    tg_smooth_table_kernel<<<(resolution + 6 + 127) / 128, 128, 0, st>>>(S, tab);
    const size_t warps = tg_smooth_warps(B, sms);
    tg_smooth_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(S, B, tab, par, x, f, status, nit, ws, wsd, maxiter, ftol);
    tg_note_launch(2);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}

extern "C" int tg_smooth_initial_batch(int d, int oldN, int N, int B, const double *old_cps, double *x0, double *scratch,
                                       void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if ((d != 2 && d != 3) || oldN < 2 || N < 2 || !old_cps || !x0 || !scratch) return 2;
    tg_smooth_initial_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d, oldN, N, B, old_cps, x0, scratch);
    tg_note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}
