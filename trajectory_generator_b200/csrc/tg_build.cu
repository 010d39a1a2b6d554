// Problem construction on the device: the parts of assembling a batch of problems that are
// arithmetic rather than bookkeeping --
//   * the initial variable vector  (TG/objectives/objective_variables.py:27-48, 63-105): control points on the
//     straight line / at equal arc-length steps along the point sequence, scale factor, waypoint scalars,
//     intermediate-waypoint times;
//   * safe-flight-corridor boxes fitted to consecutive points  (DS/safe_flight_corridor.py:109-146 and :13-16):
//     rotation, translation, rotated bounds, written into the parameter rows in the layout tg_rows_sfc reads.
// One thread per problem (per corridor for the boxes): the walks are sequential and tiny; the kernels stream
// [B][...] arrays and exist so that batches can be built where they will be solved.
#include <cuda_runtime.h>
#include <math.h>
#include "tg_spec.h"
#include "../../include/trajectory_generator_b200.h"

void tg_note_launch(int count);      // tg_api.cu

namespace {

// numpy.linalg.norm(v, 2) of a difference of two points, with numpy's arithmetic (products and sums rounded
// separately, no contraction): the walk below branches on comparisons of such lengths
__device__ __forceinline__ double seg_len(const double *s, int npts, int d, int q)
{
    double h = 0;
    for (int c = 0; c < d; c++) {
        const double v = __dsub_rn(s[c * npts + q + 1], s[c * npts + q]);
        h = c == 0 ? __dmul_rn(v, v) : __dadd_rn(h, __dmul_rn(v, v));
    }
    return sqrt(h);
}

// seq: [B][d][npts] point sequence; wseq: [B][d][nwp] waypoint locations (NULL when there are no intermediate
// waypoints); x0: [B][n]
__global__ void __launch_bounds__(128)
tg_initial_guess_kernel(const TgLayout L, int B, const double *__restrict__ seq, int npts, const double *__restrict__ wseq,
                        int nwp, double scale0, double *__restrict__ x0)
{
    const int d = L.d, N = L.N, n = L.n;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const double *s = seq + (size_t)b * d * npts;
        double *x = x0 + (size_t)b * n;
        const int nseg = npts - 1;
        if (nseg < 2) {
            // np.linspace(start_point, end_point, N).T : start + k * step, step = (stop - start) / (N - 1), last = stop
            for (int c = 0; c < d; c++) {
                const double a = s[c * npts], e = s[c * npts + 1];
                const double step = __ddiv_rn(__dsub_rn(e, a), (double)(N - 1));
                for (int k = 0; k < N; k++) x[c * N + k] = k == N - 1 ? e : __dadd_rn(__dmul_rn((double)k, step), a);
            }
        } else {
            // equal arc-length steps along the polyline (TG/objectives/objective_variables.py:69-92)
            double total = 0;
            for (int q = 0; q < nseg; q++) total = q == 0 ? seg_len(s, npts, d, 0) : __dadd_rn(seg_len(s, npts, d, q), total);
            const double spacing = __ddiv_rn(total, (double)(N - 1));
            int seg = 0;
            double walked = 0, step = 0, cum = seg_len(s, npts, d, 0), anchor[3], head[3];      // cum: length up to the end of `seg`
            for (int c = 0; c < d; c++) anchor[c] = s[c * npts];
            for (int i = 0; i < N - 1; i++) {
                const int sg = seg < nseg ? seg : nseg - 1;      // the reference would raise past the last segment
                for (int c = 0; c < d; c++) head[c] = __dsub_rn(s[c * npts + sg + 1], s[c * npts + sg]);
                const double hn = seg_len(s, npts, d, sg);
                for (int c = 0; c < d; c++) {
                    const double p = __dadd_rn(anchor[c], __dmul_rn(__ddiv_rn(head[c], hn), step));
                    x[c * N + i] = p;
                    anchor[c] = p;
                }
                step = spacing;
                walked = __dadd_rn(walked, step);
                if (cum < walked) {
                    step = __dsub_rn(walked, cum);
                    seg += 1;
                    const int sn = seg < nseg ? seg : nseg;
                    for (int c = 0; c < d; c++) anchor[c] = s[c * npts + sn];
                    if (seg < nseg) cum = __dadd_rn(seg_len(s, npts, d, seg), cum);
                }
            }
            for (int c = 0; c < d; c++) x[c * N + N - 1] = s[c * npts + npts - 1];
        }
        // scale factor, waypoint scalars (TG/objectives/objective_variables.py:37-44)
        x[L.ia] = scale0;
        for (int q = 0; q < L.nws; q++) x[L.ia + 1 + q] = 1.0;
        // intermediate waypoint times (TG/objectives/objective_variables.py:95-105)
        if (L.niw > 0) {
            const int ws = nwp - 1;
            if (ws <= 2 || !wseq) x[L.it0] = 0.5;
            else {
                const double *w = wseq + (size_t)b * d * nwp;
                double tot = 0;
                for (int q = 0; q < ws; q++) {
                    double h = 0;
                    for (int c = 0; c < d; c++) { const double v = w[c * nwp + q + 1] - w[c * nwp + q]; h += v * v; }
                    tot += sqrt(h);
                }
                double run = 0;
                for (int q = 0; q < ws - 1; q++) {
                    double h = 0;
                    for (int c = 0; c < d; c++) { const double v = w[c * nwp + q + 1] - w[c * nwp + q]; h += v * v; }
                    run += sqrt(h);
                    x[L.it0 + q] = run / tot * (double)(N - 3);
                }
            }
        }
    }
}

// points: [B][d][ncorr + 1]; pad: [B][ncorr][d] (box dimensions = pad + (segment length, 0, 0), the convention of
// test_sfc_trajectory_3D.py:33-38); par rows receive [R^T (d x d) | lower (d) | upper (d)] per corridor at p_sfc
__global__ void __launch_bounds__(128)
tg_sfc_boxes_kernel(int d, int B, int ncorr, const double *__restrict__ points, const double *__restrict__ pad,
                    double *__restrict__ par, long par_stride, int p_sfc, double *__restrict__ lengths)
{
    const int stride = d * d + 2 * d;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < B * ncorr; g += gridDim.x * blockDim.x) {
        const int b = g / ncorr, q = g - b * ncorr;
        const double *p = points + (size_t)b * d * (ncorr + 1);
        double p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0}, R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int c = 0; c < d; c++) { p1[c] = p[c * (ncorr + 1) + q]; p2[c] = p[c * (ncorr + 1) + q + 1]; }
        const double dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = p2[2] - p1[2];
        double len;
        if (d == 2) {
            // DS/safe_flight_corridor.py:109-121
            const double psi = atan2(dy, dx), cp = cos(psi), sp = sin(psi);
            R[0][0] = cp; R[0][1] = -sp; R[1][0] = sp; R[1][1] = cp;
            len = sqrt(dx * dx + dy * dy);
        } else {
            // DS/safe_flight_corridor.py:123-146: rotation = Ry^T Rz
            const double th = atan2(dz, dx), ct = cos(th), st = sin(th);
            const double dx2 = (ct * dx + 0.0 * dy) + st * dz, dy2 = (0.0 * dx + 1.0 * dy) + 0.0 * dz;      // (Ry @ distance)[0:2]
            const double psi = atan2(dy2, dx2), cp = cos(psi), sp = sin(psi);
            const double RyT[3][3] = {{ct, 0, -st}, {0, 1, 0}, {st, 0, ct}};
            const double Rz[3][3] = {{cp, -sp, 0}, {sp, cp, 0}, {0, 0, 1}};
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) R[i][j] = (RyT[i][0] * Rz[0][j] + RyT[i][1] * Rz[1][j]) + RyT[i][2] * Rz[2][j];
            len = sqrt((dx * dx + dy * dy) + dz * dz);
        }
        double *row = par + (size_t)b * par_stride + p_sfc + q * stride;
        // translation = rotation.T @ (point_1 + point_2) / 2  (matrix product first, then the division)
        for (int i = 0; i < d; i++) {
            double t = 0;
            for (int c = 0; c < d; c++) t += R[c][i] * (p1[c] + p2[c]);
            t = t / 2;
            const double dim = pad[((size_t)b * ncorr + q) * d + i] + (i == 0 ? len : 0.0);
            for (int c = 0; c < d; c++) row[i * d + c] = R[c][i];          // R^T, row-major
            row[d * d + i] = t - dim / 2;
            row[d * d + d + i] = t + dim / 2;
        }
        if (lengths) lengths[g] = len;
    }
}

// SFC_Data.__evaluate_intervals_per_corridor (DS/safe_flight_corridor.py:78-88): the shape of a corridor problem
// follows from its geometry -- intervals of corridor i = (int(round(length_i / shortest length)) + 1) * minimum, a
// single corridor gets 5.  ipc[b][ncorr]; key[b] packs them (4 bits each... 8 bits each) so that problems of one
// shape compare equal; numpy's round is round-half-to-even = rint.
__global__ void __launch_bounds__(128)
tg_sfc_intervals_kernel(int d, int B, int ncorr, const double *__restrict__ points, int min_per_corridor,
                        int *__restrict__ ipc, long long *__restrict__ key)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const double *p = points + (size_t)b * d * (ncorr + 1);
        double len[TG_MAX_CORRIDORS], shortest = INFINITY;
        for (int q = 0; q < ncorr; q++) {
            double s = 0;
            for (int c = 0; c < d; c++) { const double v = p[c * (ncorr + 1) + q + 1] - p[c * (ncorr + 1) + q]; s += v * v; }
            len[q] = sqrt(s);
            shortest = fmin(shortest, len[q]);
        }
        long long k = 0;
        for (int q = 0; q < ncorr; q++) {
            const int n = ncorr < 2 ? 5 : ((int)rint(len[q] / shortest) + 1) * min_per_corridor;
            ipc[(size_t)b * ncorr + q] = n;
            k = (k << 8) | (long long)(n & 0xff);
        }
        if (key) key[b] = k;
    }
}

}  // namespace

extern "C" int tg_sfc_intervals_batch(int d, int ncorr, int B, const double *points, int min_per_corridor, int *ipc,
                                      long long *key, void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if ((d != 2 && d != 3) || ncorr < 1 || ncorr > TG_MAX_CORRIDORS || !points || !ipc || min_per_corridor < 1) return 2;
    int grid = (B + 127) / 128;
    if (grid > 148 * 16) grid = 148 * 16;
    tg_sfc_intervals_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d, B, ncorr, points, min_per_corridor, ipc, key);
    tg_note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}

extern "C" int tg_initial_guess_batch(const int *spec, int B, const double *seq, int npts, const double *wseq, int nwp,
                                      double scale0, double *x0, void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if (!spec || !seq || !x0 || npts < 2) return 2;
    TgLayout L;
    tg_make_layout(spec, &L);
    if (L.niw > 0 && wseq && nwp != L.niw + 2) return 2;
    int grid = (B + 127) / 128;
    if (grid > 148 * 16) grid = 148 * 16;
    tg_initial_guess_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(L, B, seq, npts, wseq, nwp, scale0, x0);
    tg_note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}

extern "C" int tg_sfc_boxes_batch(const int *spec, int B, const double *points, const double *pad, double *par,
                                  double *lengths, void *stream)
{
    if (B <= 0) return 0;
    int rc = tg_device_check();
    if (rc) return rc;
    if (!spec || !points || !pad || !par) return 2;
    TgLayout L;
    tg_make_layout(spec, &L);
    const int ncorr = spec[TG_SP_NCORR];
    if (ncorr < 1) return 2;
    int grid = (B * ncorr + 127) / 128;
    if (grid > 148 * 16) grid = 148 * 16;
    tg_sfc_boxes_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(L.d, B, ncorr, points, pad, par, (long)L.P, L.p_sfc, lengths);
    tg_note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : 100 + (int)e;
}
