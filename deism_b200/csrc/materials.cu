// materials.cu — shape-preserving cubic (PCHIP) interpolation of wall material data onto the
// dense frequency grid, straight into the device-resident Z_S[walls, K].
//
// Replaces interpolate_functions (deism/core_deism.py:1126-1163): real and imaginary parts are
// interpolated separately with scipy's PchipInterpolator (Fritsch–Butland three-point harmonic-mean
// slopes, shape-preserving one-sided end slopes), query frequencies are clipped to the measured
// band first so that the end values are held outside it, and a single band is broadcast.
#include "common.cuh"

namespace deism {

constexpr int PCHIP_MAXB = 64;     // material bands (octave / third-octave tables have <= 31)

__device__ __forceinline__ double sgn(double v) { return (v > 0.0) - (v < 0.0); }

// one-sided three-point end slope, kept shape preserving
__device__ __forceinline__ double pchip_end(double h0, double h1, double m0, double m1) {
    double d = ((2.0 * h0 + h1) * m0 - h0 * m1) / (h0 + h1);
    if (sgn(d) != sgn(m0)) d = 0.0;
    else if (sgn(m0) != sgn(m1) && fabs(d) > 3.0 * fabs(m0)) d = 3.0 * m0;
    return d;
}

// blockIdx.y = row; one thread per dense frequency.  Slopes of the row's two curves (Re, Im) are
// rebuilt per CTA in shared memory: B is tiny.
__global__ void __launch_bounds__(256)
pchip_kernel(int rows, int B, long K, const double *__restrict__ xs, const cplx *__restrict__ ys,
             const double *__restrict__ xq, cplx *out) {
    __shared__ double x[PCHIP_MAXB], y[2][PCHIP_MAXB], d[2][PCHIP_MAXB];
    const int row = blockIdx.y;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        x[i] = xs[i];
        cplx v = ys[(long)row * B + i];
        y[0][i] = v.x; y[1][i] = v.y;
    }
    __syncthreads();
    if (B >= 2) {
        for (int e = threadIdx.x; e < 2 * B; e += blockDim.x) {
            const int c = e / B, i = e % B;
            const double *yy = y[c];
            double dv;
            if (B == 2) {
                dv = (yy[1] - yy[0]) / (x[1] - x[0]);
            } else if (i == 0) {
                dv = pchip_end(x[1] - x[0], x[2] - x[1], (yy[1] - yy[0]) / (x[1] - x[0]),
                               (yy[2] - yy[1]) / (x[2] - x[1]));
            } else if (i == B - 1) {
                dv = pchip_end(x[B - 1] - x[B - 2], x[B - 2] - x[B - 3],
                               (yy[B - 1] - yy[B - 2]) / (x[B - 1] - x[B - 2]),
                               (yy[B - 2] - yy[B - 3]) / (x[B - 2] - x[B - 3]));
            } else {
                const double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
                const double m0 = (yy[i] - yy[i - 1]) / h0, m1 = (yy[i + 1] - yy[i]) / h1;
                if (sgn(m0) != sgn(m1) || m0 == 0.0 || m1 == 0.0) {
                    dv = 0.0;
                } else {
                    const double w1 = 2.0 * h1 + h0, w2 = h1 + 2.0 * h0;
                    dv = 1.0 / ((w1 / m0 + w2 / m1) / (w1 + w2));
                }
            }
            d[c][i] = dv;
        }
    }
    __syncthreads();
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= K) return;
    if (B == 1) {
        out[(long)row * K + q] = cmake(y[0][0], y[1][0]);
        return;
    }
    double xv = xq[q];
    xv = xv < x[0] ? x[0] : (xv > x[B - 1] ? x[B - 1] : xv);
    int i = 0;
    while (i + 2 < B && xv >= x[i + 1]) ++i;        // x[i] <= xv < x[i+1]; last interval closed
    const double h = x[i + 1] - x[i], s = xv - x[i];
    double r[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        // cubic Hermite segment in the local power basis, as scipy's CubicHermiteSpline builds it
        const double slope = (y[c][i + 1] - y[c][i]) / h;
        const double t = (d[c][i] + d[c][i + 1] - 2.0 * slope) / h;
        const double c3 = t / h, c2 = (slope - d[c][i]) / h - t, c1 = d[c][i], c0 = y[c][i];
        r[c] = c0 + c1 * s + c2 * (s * s) + c3 * (s * s * s);
    }
    out[(long)row * K + q] = cmake(r[0], r[1]);
}

}  // namespace deism

using namespace deism;

extern "C" int deism_pchip_interp(int rows, int n_bands, int64_t K, const double *d_sparse_freqs,
                                  const double *d_data, const double *d_dense_freqs, double *d_out,
                                  void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    if (rows <= 0 || K <= 0) return DEISM_OK;
    if (n_bands < 1 || n_bands > PCHIP_MAXB)
        return set_error(DEISM_ERR_INVALID, "n_bands %d outside 1..%d", n_bands, PCHIP_MAXB);
    if (!d_sparse_freqs || !d_data || !d_dense_freqs || !d_out)
        return set_error(DEISM_ERR_INVALID, "null input pointer");
    dim3 grid((unsigned)((K + 255) / 256), (unsigned)rows);
    pchip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, n_bands, K, d_sparse_freqs,
                                                         (const cplx *)d_data, d_dense_freqs,
                                                         (cplx *)d_out);
    DEISM_LAUNCH_CHECK("pchip_kernel");
    return DEISM_OK;
}
