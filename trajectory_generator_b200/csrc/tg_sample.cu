// Output sampling of solved trajectories: the reference's
// matrix_bspline_evaluation_for_dataset / ..._derivative_evaluation_for_dataset / ..._for_discrete_steps
// (TG/matrix_evaluation.py:5-173) for a whole batch of cubic splines.
//
// One thread per output sample.  The kernel is a pure stream: it reads d x N control points per trajectory
// (L1/L2 resident: 4 d doubles feed every sample of an interval) and writes d doubles per sample, coordinate-major
// [B][d][cap] exactly like the reference's spline_data[d, num_points], so consecutive threads write consecutive
// 8-byte words.  Bound: HBM writes, 8 d bytes per sample.
//
// Sample times follow numpy.linspace bit for bit (y_k = k * step + start with separate rounding of the product
// and the sum, last sample = stop); the interval of a sample is chosen with the reference's masks
// (time >= i) & (time < i + 1), the last interval also taking time == i + 1; a sample beyond the last knot is
// dropped by the reference (it can only be the last one), which leaves the zero the output array was created with.
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/trajectory_generator_b200.h"

void tg_note_launch(int count);      // tg_api.cu

namespace {

struct SampleArgs {
    int d, N, B;
    const double *cps; long cps_stride;        // control points of trajectory b at cps + b * cps_stride, row-major d x N
    const double *scale; long scale_stride;    // scale factor of trajectory b at scale[b * scale_stride]
    int rth;                                   // derivative order 0..3
    int mode;                                  // 0: num_points over [0, N-3] intervals; 1: discrete steps dt from offset[b]
    int num_points;
    const double *offset;                      // mode 1: starting offset per trajectory (may be NULL = 0)
    double dt;
    double *out; long cap;                     // out[b][c][k], k < cap
    double *times;                             // mode 1 (optional): time_data[b][k] (without start_time)
    int *counts;                               // mode 1: samples of trajectory b
};

// M of TG/matrix_evaluation.py:245-250 ( /12 applied element-wise as numpy does)
__device__ __forceinline__ double m3(int l, int col)
{
    const double M[4][4] = {{-2.0, 6.0, -6.0, 2.0}, {6.0, -12.0, 0.0, 8.0}, {-6.0, 6.0, 6.0, 2.0}, {2.0, 0.0, 0.0, 0.0}};
    return M[l][col] / 12.0;
}

__device__ __forceinline__ double ipow(double x, int e)
{
    // numpy's steps_array ** e for e = 0..3 (pow() is exact for e <= 1 and within an ulp of these products)
    return e == 0 ? 1.0 : e == 1 ? x : e == 2 ? x * x : x * x * x;
}

__global__ void __launch_bounds__(256) tg_sample_kernel(const SampleArgs a)
{
    const int nint = a.N - 3;
    const long per = (long)a.cap;
    // blockIdx.y strides over trajectories, blockIdx.x / threadIdx.x over the samples of one trajectory
    for (int b = blockIdx.y; b < a.B; b += gridDim.y)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.cap; k += gridDim.x * blockDim.x) {
        const double sf = a.scale ? a.scale[(long)b * a.scale_stride] : 1.0;
        int num;
        double t;            // sample time in units of intervals
        double tdata = 0;    // time_data entry (mode 1)
        if (a.mode == 0) {
            num = a.num_points;
            if (k >= num) continue;
            // np.linspace(0, nint, num)
            const int div = num - 1;
            const double step = div > 0 ? (double)nint / (double)div : 0.0;
            t = (div > 0 && k == div) ? (double)nint : __dmul_rn((double)k, step);
        } else {
            const double off = a.offset ? a.offset[b] : 0.0;
            const double duration = __dmul_rn(sf, (double)nint);
            num = (int)(__ddiv_rn(__dsub_rn(duration, off), a.dt)) + 1;          // int((duration - offset) / dt) + 1
            if (k == 0 && a.counts) a.counts[b] = num;
            if (k >= num) continue;
            const double last = __dadd_rn(__dmul_rn((double)(num - 1), a.dt), off);
            const int div = num - 1;
            const double step = div > 0 ? __ddiv_rn(__dsub_rn(last, off), (double)div) : 0.0;
            tdata = (div > 0 && k == div) ? last : __dadd_rn(__dmul_rn((double)k, step), off);
            if (a.times) a.times[(long)b * per + k] = tdata;
            t = __ddiv_rn(tdata, sf);
        }
        double *o = a.out + ((long)b * a.d) * per + k;
        if (!(t >= 0.0) || t > (double)nint) {          // dropped by the reference's masks: the array keeps its zero
            for (int c = 0; c < a.d; c++) o[(long)c * per] = 0.0;
            continue;
        }
        int i = (int)t;
        if (i > nint - 1) i = nint - 1;
        const double tau = t - (double)i;
        // (K L_r)[col] = (3-col)! / (3-r-col)! / sf^r * tau^(3-r-col), col <= 3 - r   (TG/matrix_evaluation.py:175-180)
        double w[4];
        const double sr = a.rth == 0 ? 1.0 : a.rth == 1 ? sf : a.rth == 2 ? sf * sf : sf * sf * sf;
#pragma unroll
        for (int col = 0; col < 4; col++) {
            double v = 0.0;
            if (col <= 3 - a.rth) {
                double fac = 1.0;
                for (int q = 0; q < a.rth; q++) fac *= (double)(3 - col - q);
                v = (a.rth == 0 ? 1.0 : fac / sr) * ipow(tau, 3 - a.rth - col);
            }
            w[col] = v;
        }
        const double *P = a.cps + (long)b * a.cps_stride;
        for (int c = 0; c < a.d; c++) {
            const double p0 = P[c * a.N + i], p1 = P[c * a.N + i + 1], p2 = P[c * a.N + i + 2], p3 = P[c * a.N + i + 3];
            double s = 0.0;
#pragma unroll
            for (int col = 0; col < 4; col++) {
                const double coef = ((p0 * m3(0, col) + p1 * m3(1, col)) + p2 * m3(2, col)) + p3 * m3(3, col);    // (P M)[c, col]
                s += coef * w[col];
            }
            o[(long)c * per] = s;
        }
    }
}

}  // namespace

extern "C" int tg_sample_batch(int d, int N, int B, const double *cps, long cps_stride, const double *scale,
                               long scale_stride, int derivative_order, int mode, int num_points, const double *offset,
                               double dt, double *out, long capacity, double *times, int *counts, void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if ((d != 2 && d != 3) || N < 4 || derivative_order < 0 || derivative_order > 3 || capacity < 1 || !cps || !out ||
        (mode == 0 && (num_points < 1 || num_points > capacity)) || (mode == 1 && !(dt > 0)) || (mode != 0 && mode != 1))
        return 2;
    SampleArgs a = {d, N, B, cps, cps_stride, scale, scale_stride, derivative_order, mode, num_points, offset, dt, out,
                    capacity, times, counts};
    long bx = (capacity + 255) / 256;
    if (bx > 1024) bx = 1024;
    const dim3 grid((unsigned)bx, (unsigned)(B < 65535 ? B : 65535));
    tg_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    tg_note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}
