// Output sampling of solved trajectories: the reference's
// matrix_bspline_evaluation_for_dataset / ..._derivative_evaluation_for_dataset / ..._for_discrete_steps
// (TG/matrix_evaluation.py:5-173) for a whole batch of cubic splines.
//
// A warp takes 64 consecutive samples of one trajectory at a time.  The kernel is a pure stream: it reads d x N control
// points per trajectory (L1/L2 resident: 4 d doubles feed every sample of an interval) and writes d doubles per sample, coordinate-major
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
    double step0;                              // mode 0: (N-3) / (num_points-1), the step of numpy.linspace
    const double *offset;                      // mode 1: starting offset per trajectory (may be NULL = 0)
    double dt;
    double *out; long cap;                     // out[b][c][k], k < cap
    double *times;                             // mode 1 (optional): time_data[b][k] (without start_time)
    int *counts;                               // mode 1: samples of trajectory b
};

// basis matrices of TG/matrix_evaluation.py:224-262 (orders 2 .. 5; the scalar factor applied element-wise as numpy does)
template <int ORD>
__device__ __forceinline__ double mk(int l, int col)
{
    if (ORD == 2) {
        const double M[3][3] = {{1.0, -2.0, 1.0}, {-2.0, 2.0, 1.0}, {1.0, 0.0, 0.0}};
        return 0.5 * M[l][col];
    } else if (ORD == 3) {
        const double M[4][4] = {{-2.0, 6.0, -6.0, 2.0}, {6.0, -12.0, 0.0, 8.0}, {-6.0, 6.0, 6.0, 2.0}, {2.0, 0.0, 0.0, 0.0}};
        return M[l][col] / 12.0;
    } else if (ORD == 4) {
        const double M[5][5] = {{1.0, -4.0, 6.0, -4.0, 1.0}, {-4.0, 12.0, -6.0, -12.0, 11.0}, {6.0, -12.0, -6.0, 12.0, 11.0},
                                {-4.0, 4.0, 6.0, 4.0, 1.0}, {1.0, 0.0, 0.0, 0.0, 0.0}};
        return M[l][col] / 24.0;
    } else {
        const double M[6][6] = {{-1.0, 5.0, -10.0, 10.0, -5.0, 1.0}, {5.0, -20.0, 20.0, 20.0, -50.0, 26.0},
                                {-10.0, 30.0, 0.0, -60.0, 0.0, 66.0}, {10.0, -20.0, -20.0, 20.0, 50.0, 26.0},
                                {-5.0, 5.0, 10.0, 10.0, 5.0, 1.0}, {1.0, 0.0, 0.0, 0.0, 0.0, 0.0}};
        return M[l][col] / 120.0;
    }
}

__device__ __forceinline__ double ipow(double x, int e)
{
    // numpy's steps_array ** e for e = 0..5 (pow() is exact for e <= 1 and within an ulp of these products)
    double r = 1.0;
#pragma unroll
    for (int q = 0; q < 5; q++) if (q < e) r = q == 0 ? x : r * x;
    return r;
}

// one sample: trajectory b (control points P, scale sf, derivative weights kd), sample index k
template <int MODE, int RTH, int D, int ORD>
__device__ __forceinline__ void tg_sample_one(const SampleArgs &a, int b, int k, const double *P, double sf, const double kd[ORD + 1],
                                              int num, double step, double off, double last, double v[D])
{
    const int nint = a.N - ORD, div = num - 1;
    constexpr int d = D;
    const long per = (long)a.cap;
    double t;            // sample time in units of intervals
    if (MODE == 0) {
        t = (div > 0 && k == div) ? (double)nint : __dmul_rn((double)k, step);          // np.linspace(0, nint, num)
    } else {
        const double tdata = (div > 0 && k == div) ? last : __dadd_rn(__dmul_rn((double)k, step), off);
        if (a.times) a.times[(long)b * per + k] = tdata;
        t = __ddiv_rn(tdata, sf);
    }
    if (!(t >= 0.0) || t > (double)nint) {          // dropped by the reference's masks: the array keeps its zero
#pragma unroll
        for (int c = 0; c < d; c++) v[c] = 0.0;
        return;
    }
    int i = (int)t;
    if (i > nint - 1) i = nint - 1;
    const double tau = t - (double)i;
    // weights of the order + 1 control points: M (K L_r)  (TG/matrix_evaluation.py:26-29, 127-130 with the products
    // associated as P (M L); the reference forms (P M) L -- the same sums up to the last place)
    double wl[ORD + 1], w[ORD + 1];
#pragma unroll
    for (int col = 0; col <= ORD; col++)
        w[col] = col <= ORD - RTH ? (RTH == 0 ? ipow(tau, ORD - col) : kd[col] * ipow(tau, ORD - RTH - col < 0 ? 0 : ORD - RTH - col)) : 0.0;
#pragma unroll
    for (int l = 0; l <= ORD; l++) {
        double h = mk<ORD>(l, 0) * w[0];
#pragma unroll
        for (int col = 1; col <= ORD; col++) h = h + mk<ORD>(l, col) * w[col];
        wl[l] = h;
    }
#pragma unroll
    for (int c = 0; c < d; c++) {
        const double *p = P + c * a.N + i;
        double h = p[0] * wl[0];
#pragma unroll
        for (int l = 1; l <= ORD; l++) h = h + p[l] * wl[l];
        v[c] = h;
    }
}

// Work item = (trajectory, chunk of 256 consecutive samples), handed to warps in a grid-stride loop: no block-level
// synchronisation, the per-trajectory set-up is shared by 8 samples per lane.  A lane takes PAIRS of consecutive samples
// and writes each coordinate row with one 16-byte store (st.global.v2.f64) when the rows are 16-byte aligned (even
// capacity), so a warp writes 512 contiguous bytes per coordinate and instruction.
template <int MODE, int RTH, int D, int ORD>
__global__ void __launch_bounds__(256) tg_sample_kernel(const SampleArgs a)
{
    const int nint = a.N - ORD;
    const int lane = threadIdx.x & 31;
    const int chunks = (int)((a.cap + 255) / 256);
    const long items = (long)a.B * chunks;
    const long warps = (long)gridDim.x * (blockDim.x >> 5);
    for (long item = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < items; item += warps) {
        const int b = (int)((unsigned)item / (unsigned)chunks), k0 = (int)(item - (long)b * chunks) * 256;
        const double *P = a.cps + (long)b * a.cps_stride;
        const double sf = a.scale ? a.scale[(long)b * a.scale_stride] : 1.0;
        // (K L_r)[col] = (p-col)! / (p-r-col)! / sf^r * tau^(p-r-col), col <= p - r, p = order   (TG/matrix_evaluation.py:175-180)
        double kd[ORD + 1];
        {
            const double sr = RTH == 0 ? 1.0 : RTH == 1 ? sf : RTH == 2 ? sf * sf : sf * sf * sf;
#pragma unroll
            for (int col = 0; col <= ORD; col++) {
                double fac = 1.0;
#pragma unroll
                for (int q = 0; q < RTH; q++) fac *= (double)(ORD - col - q);
                kd[col] = col <= ORD - RTH ? (RTH == 0 ? 1.0 : fac / sr) : 0.0;
            }
        }
        int num = a.num_points;
        double off = 0, last = 0, step = a.step0;
        if (MODE == 1) {
            off = a.offset ? a.offset[b] : 0.0;
            const double duration = __dmul_rn(sf, (double)nint);
            num = (int)(__ddiv_rn(__dsub_rn(duration, off), a.dt)) + 1;          // int((duration - offset) / dt) + 1
            if (k0 == 0 && lane == 0 && a.counts) a.counts[b] = num;
            last = __dadd_rn(__dmul_rn((double)(num - 1), a.dt), off);
            step = num > 1 ? __ddiv_rn(__dsub_rn(last, off), (double)(num - 1)) : 0.0;
        }
        const int kend = num < a.cap ? num : (int)a.cap;
        const bool vec = (a.cap & 1) == 0 && (reinterpret_cast<size_t>(a.out) & 15) == 0;
        double *orow = a.out + ((long)b * D) * (long)a.cap;
#pragma unroll 2
        for (int k = k0 + 2 * lane; k < k0 + 256 && k < kend; k += 64) {
            double v0[D], v1[D];
            tg_sample_one<MODE, RTH, D, ORD>(a, b, k, P, sf, kd, num, step, off, last, v0);
            const bool two = k + 1 < kend;
            if (two) tg_sample_one<MODE, RTH, D, ORD>(a, b, k + 1, P, sf, kd, num, step, off, last, v1);
#pragma unroll
            for (int c = 0; c < D; c++) {
                double *o = orow + (long)c * a.cap + k;
                if (vec && two) *reinterpret_cast<double2 *>(o) = make_double2(v0[c], v1[c]);
                else { o[0] = v0[c]; if (two) o[1] = v1[c]; }
            }
        }
    }
}


// One point of one interval per item (TG/matrix_evaluation.py:183-222: evaluate_point_on_interval /
// evaluate_point_derivative_on_interval): out[b][c] = cps[b][c][0 .. p] . M . T with
// T[i] = (t - tj)^(p-r-i) / sf^(p-i) * (p-i)! / (p-i-r)!   (r = 0: ((t - tj) / sf)^(p-i)).
template <int ORD>
__global__ void tg_interval_point_kernel(int d, int B, const double *cps, const double *t, const double *tj, const double *sf,
                                         int rth, double *out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double dtv = t[b] - tj[b], s = sf[b];
    double T[ORD + 1];
#pragma unroll
    for (int i = 0; i <= ORD; i++) {
        if (rth == 0) T[i] = pow(dtv / s, (double)(ORD - i));
        else if (i <= ORD - rth) {
            double fa = 1.0;
            for (int q = 0; q < rth; q++) fa *= (double)(ORD - i - q);
            T[i] = pow(dtv, (double)(ORD - rth - i)) / pow(s, (double)(ORD - i)) * fa;
        } else T[i] = 0.0;
    }
    for (int c = 0; c < d; c++) {
        const double *p = cps + ((long)b * d + c) * (ORD + 1);
        double pm[ORD + 1];
#pragma unroll
        for (int col = 0; col <= ORD; col++) {
            double h = p[0] * mk<ORD>(0, col);
#pragma unroll
            for (int l = 1; l <= ORD; l++) h = h + p[l] * mk<ORD>(l, col);
            pm[col] = h;
        }
        double h = pm[0] * T[0];
#pragma unroll
        for (int col = 1; col <= ORD; col++) h = h + pm[col] * T[col];
        out[(long)b * d + c] = h;
    }
}

}  // namespace

extern "C" int tg_interval_points_batch(int order, int d, int B, const double *cps, const double *t, const double *tj,
                                        const double *scale, int derivative_order, double *out, void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if (order < 2 || order > 5 || d < 1 || derivative_order < 0 || !cps || !t || !tj || !scale || !out) return 2;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (B + 127) / 128;
    if (order == 2) tg_interval_point_kernel<2><<<blocks, 128, 0, st>>>(d, B, cps, t, tj, scale, derivative_order, out);
    else if (order == 3) tg_interval_point_kernel<3><<<blocks, 128, 0, st>>>(d, B, cps, t, tj, scale, derivative_order, out);
    else if (order == 4) tg_interval_point_kernel<4><<<blocks, 128, 0, st>>>(d, B, cps, t, tj, scale, derivative_order, out);
    else tg_interval_point_kernel<5><<<blocks, 128, 0, st>>>(d, B, cps, t, tj, scale, derivative_order, out);
    tg_note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}

extern "C" int tg_sample_batch_order(int order, int d, int N, int B, const double *cps, long cps_stride, const double *scale,
                                     long scale_stride, int derivative_order, int mode, int num_points, const double *offset,
                                     double dt, double *out, long capacity, double *times, int *counts, void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if (order < 2 || order > 5 || (d != 2 && d != 3) || N < order + 1 || derivative_order < 0 || derivative_order > 3 ||
        derivative_order > order || capacity < 1 || !cps || !out ||
        (mode == 0 && (num_points < 1 || num_points > capacity)) || (mode == 1 && !(dt > 0)) || (mode != 0 && mode != 1))
        return 2;
    const double step0 = (mode == 0 && num_points > 1) ? (double)(N - order) / (double)(num_points - 1) : 0.0;
    SampleArgs a = {d, N, B, cps, cps_stride, scale, scale_stride, derivative_order, mode, num_points, step0, offset, dt,
                    out, capacity, times, counts};

    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long items = (long)B * ((capacity + 255) / 256);
    if (items > 0x7fffffffL) return 2;
    long blocks = (items + 7) / 8;
    if (blocks > (long)sms * 8) blocks = (long)sms * 8;
    cudaStream_t st = (cudaStream_t)stream;
#define TG_SAMPLE_LAUNCH(M, R, O)                                                                                    \
    do { if (d == 2) tg_sample_kernel<M, R, 2, O><<<(int)blocks, 256, 0, st>>>(a); else tg_sample_kernel<M, R, 3, O><<<(int)blocks, 256, 0, st>>>(a); } while (0)
#define TG_SAMPLE_RTH(M, O)                                                                                          \
    do {                                                                                                             \
        if (derivative_order == 0) TG_SAMPLE_LAUNCH(M, 0, O); else if (derivative_order == 1) TG_SAMPLE_LAUNCH(M, 1, O); \
        else if (derivative_order == 2) TG_SAMPLE_LAUNCH(M, 2, O); else TG_SAMPLE_LAUNCH(M, 3, O);                   \
    } while (0)
#define TG_SAMPLE_MODE(M)                                                                                            \
    do {                                                                                                             \
        if (order == 3) TG_SAMPLE_RTH(M, 3); else if (order == 2) TG_SAMPLE_RTH(M, 2);                               \
        else if (order == 4) TG_SAMPLE_RTH(M, 4); else TG_SAMPLE_RTH(M, 5);                                          \
    } while (0)
    if (mode == 0) TG_SAMPLE_MODE(0); else TG_SAMPLE_MODE(1);
    tg_note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}

extern "C" int tg_sample_batch(int d, int N, int B, const double *cps, long cps_stride, const double *scale,
                               long scale_stride, int derivative_order, int mode, int num_points, const double *offset,
                               double dt, double *out, long capacity, double *times, int *counts, void *stream)
{
    return tg_sample_batch_order(3, d, N, B, cps, cps_stride, scale, scale_stride, derivative_order, mode, num_points, offset,
                                 dt, out, capacity, times, counts, stream);
}
