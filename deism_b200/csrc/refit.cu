// refit.cu — DEISM-ARG source directivity of every image source.
//
// Replaces cal_C_nm_s_new (deism/core_deism.py:1545-1605) with its per-image
// SHCs_from_pressure_LS (cd:1481-1509: Y_pinv @ Psh.T) and vectorize_C_nm_s_ARG (cd:1193-1217):
//
//   for every image i:   X_i = R_i @ X                       (reflected sampling directions)
//                        Y_i[s, j] = Y_{n_j}^{m_j}(az(X_i[:,s]), pi/2 - el(X_i[:,s]))
//                        f_i = pinv(Y_i) @ Psh^T             (least squares, [J, K])
//                        C_i[k, j] = f_i[j, k] / h_n^{(2)}(k r0)          -> complex64 [K, J, I]
//
// The reference pays one SVD pseudo-inverse of an [n_dir x J] matrix and one [J x n_dir] . [n_dir x K]
// product per image on the host.  Here:
//
//   * the fit is done in the REAL harmonic basis (a_m = Re Y_n^m, b_m = Im Y_n^m; the complex and
//     the real basis span the same space, so the least-squares fit is the same one) — the normal
//     matrix G_i = Yr_i^T Yr_i is real symmetric [J x J] and is Cholesky-factorised per image;
//   * refit_pinv_kernel (one CTA per image) writes W_i = G_i^-1 Yr_i^T, the real [J x n_dir]
//     pseudo-inverse, image index innermost;
//   * refit_apply_kernel contracts W with the sampled pressures (real x complex, thread = image,
//     16 frequencies per thread), maps real-basis coefficients back to (n, +-m), divides by the
//     spherical Hankel function and stores complex64 coalesced along the image axis.
//
// R_i is the float32 matrix the reflection-tree engine produced; it is applied in double, as
// numpy's float32 @ float64 product does (cd:1577).
#include "common.cuh"

namespace deism {

constexpr int RF_NMAX = 7;                        // source order supported by the shared-memory budget
constexpr int RF_JMAX = (RF_NMAX + 1) * (RF_NMAX + 1);
constexpr int RF_CH = 128;                        // samples per chunk
constexpr int RF_THREADS = 256;
constexpr int RF_KT = 16;                         // frequencies per thread in the apply kernel
constexpr int RF_SC = 64;                         // samples staged per step in the apply kernel

// Real functions are ordered pairs first — (a_m, b_m) for n = 1..N, m = 1..n at q = 2p, 2p+1 with
// p = n(n-1)/2 + m - 1 — then a_0 of n = 0..N at q = N(N+1) + n: pairs stay 16-byte aligned.
__device__ __forceinline__ void real_harmonics(int N, double phi, double ct, double *y, int stride) {
    const int npairs = N * (N + 1) / 2;
    for (int n = 0; n <= N; ++n) {
        y[(2 * npairs + n) * stride] = sph_harm_dev(0, n, phi, ct).x;
        for (int m = 1; m <= n; ++m) {
            const cplx v = sph_harm_dev(m, n, phi, ct);
            const int p = n * (n - 1) / 2 + m - 1;
            y[(2 * p) * stride] = v.x;
            y[(2 * p + 1) * stride] = v.y;
        }
    }
}

// direction of sample s seen from image i: (az, cos(pi/2 - el)) as cart2sph + the caller's
// pi/2 - el make them (utilities.py:95-102, cd:1577-1586)
__device__ __forceinline__ void reflected_direction(const double *R, const double *X, long n_dir, long s,
                                                    double *phi, double *ct) {
    const double x = X[s], y = X[n_dir + s], z = X[2 * n_dir + s];
    const double vx = R[0] * x + R[1] * y + R[2] * z;
    const double vy = R[3] * x + R[4] * y + R[5] * z;
    const double vz = R[6] * x + R[7] * y + R[8] * z;
    const double el = atan2(vz, hypot(vx, vy));
    double sn, cs;
    sincos_cw(1.5707963267948966 - el, &sn, &cs);
    *phi = atan2(vy, vx);
    *ct = cs;
}

struct RefitSmem {
    double R[9];
    int bad;
    double G[RF_JMAX][RF_JMAX + 1];          // normal matrix, then its Cholesky factor (lower)
    double y[RF_CH][RF_JMAX + 1];            // real harmonics of a chunk of samples, one row per thread
};

// W layout: pairs   W2[p][s][I_pad] double2   (p < npairs)
//           singles W1[n][s][I_pad] double    (n <= N)
__global__ void __launch_bounds__(RF_THREADS)
refit_pinv_kernel(long n_images, long I_pad, int N, long n_dir, const float *__restrict__ refl,
                  const double *__restrict__ X, double2 *W2, double *W1, int *status) {
    extern __shared__ __align__(16) unsigned char rf_smem_raw[];
    RefitSmem &sm = *reinterpret_cast<RefitSmem *>(rf_smem_raw);
    const long img = blockIdx.x;
    const int tid = threadIdx.x;
    const int J = (N + 1) * (N + 1), npairs = N * (N + 1) / 2;
    if (tid < 9) sm.R[tid] = (double)refl[img * 9 + tid];
    if (tid == 0) sm.bad = 0;
    for (int e = tid; e < J * J; e += RF_THREADS) sm.G[e / J][e % J] = 0.0;
    __syncthreads();

    // ---- pass 1: G = sum_s y(s) y(s)^T ----
    for (long s0 = 0; s0 < n_dir; s0 += RF_CH) {
        const int ns = (int)(n_dir - s0 < RF_CH ? n_dir - s0 : RF_CH);
        if (tid < ns) {
            double phi, ct;
            reflected_direction(sm.R, X, n_dir, s0 + tid, &phi, &ct);
            real_harmonics(N, phi, ct, &sm.y[tid][0], 1);
        }
        __syncthreads();
        for (int e = tid; e < J * J; e += RF_THREADS) {
            const int q = e / J, q2 = e % J;
            if (q2 > q) continue;
            double acc = 0.0;
            for (int t = 0; t < ns; ++t) acc = fma(sm.y[t][q], sm.y[t][q2], acc);
            sm.G[q][q2] += acc;
        }
        __syncthreads();
    }
    // ---- Cholesky G = L L^T (lower triangle in place), column by column ----
    for (int c = 0; c < J; ++c) {
        if (tid == 0) {
            double d = sm.G[c][c];
            for (int t = 0; t < c; ++t) d = fma(-sm.G[c][t], sm.G[c][t], d);
            if (!(d > 0.0)) { sm.bad = 1; d = 1.0; }
            sm.G[c][c] = sqrt(d);
        }
        __syncthreads();
        const double inv = 1.0 / sm.G[c][c];
        for (int r = c + 1 + tid; r < J; r += RF_THREADS) {
            double v = sm.G[r][c];
            for (int t = 0; t < c; ++t) v = fma(-sm.G[r][t], sm.G[c][t], v);
            sm.G[r][c] = v * inv;
        }
        __syncthreads();
    }
    if (sm.bad) {
        if (tid == 0) atomicExch(status, 1);      // rank-deficient sampling: reported by the host
        return;
    }
    // ---- pass 2: w(s) = G^-1 y(s), one sample per thread ----
    for (long s0 = 0; s0 < n_dir; s0 += RF_CH) {
        const int ns = (int)(n_dir - s0 < RF_CH ? n_dir - s0 : RF_CH);
        if (tid < ns) {
            double phi, ct;
            reflected_direction(sm.R, X, n_dir, s0 + tid, &phi, &ct);
            double *y = &sm.y[tid][0];
            real_harmonics(N, phi, ct, y, 1);
            for (int r = 0; r < J; ++r) {                       // L z = y
                double v = y[r];
                for (int t = 0; t < r; ++t) v = fma(-sm.G[r][t], y[t], v);
                y[r] = v / sm.G[r][r];
            }
            for (int r = J - 1; r >= 0; --r) {                  // L^T w = z
                double v = y[r];
                for (int t = r + 1; t < J; ++t) v = fma(-sm.G[t][r], y[t], v);
                y[r] = v / sm.G[r][r];
            }
            const long s = s0 + tid;
            for (int p = 0; p < npairs; ++p)
                W2[((long)p * n_dir + s) * I_pad + img] = make_double2(y[2 * p], y[2 * p + 1]);
            for (int n = 0; n <= N; ++n) W1[((long)n * n_dir + s) * I_pad + img] = y[2 * npairs + n];
        }
        __syncthreads();
    }
}

// 1 / h_n^{(2)}(k r0), n <= N: upward recurrence of (j_n, -y_n) as pb:86-118 does it
__global__ void refit_hankel_kernel(long K, int N, const double *__restrict__ k, double r0, cplx *Hinv) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const double x = k[i] * r0, ix = 1.0 / x;
    double sn, cs;
    sincos_cw(x, &sn, &cs);
    cplx hm = cmake(sn * ix, cs * ix);                                            // h_0 = (j_0, -y_0)
    cplx hc = cmake(fma(sn * ix, ix, -cs * ix), fma(cs * ix, ix, sn * ix));       // h_1
    for (int n = 0; n <= N; ++n) {
        cplx h;
        if (n == 0) h = hm;
        else if (n == 1) h = hc;
        else {
            const double cf = (double)(2 * n - 1) * ix;
            h = cmake(fma(cf, hc.x, -hm.x), fma(cf, hc.y, -hm.y));
            hm = hc; hc = h;
        }
        // stored as (j, -y): h^(2) = j - i y = h.x + i h.y
        Hinv[(long)n * K + i] = cdiv(cmake(1.0, 0.0), h);
    }
}

// blockIdx.y: group g < npairs -> pair (n, m) of real functions, else single a_0 of n = g - npairs
__global__ void __launch_bounds__(128)
refit_apply_kernel(long n_images, long I_pad, long K, int N, long n_dir, const double2 *__restrict__ W2,
                   const double *__restrict__ W1, const cplx *__restrict__ Psh,
                   const cplx *__restrict__ Hinv, float2 *out) {
    __shared__ double2 Ps[RF_SC][RF_KT];
    const long img = (long)blockIdx.x * 128 + threadIdx.x;
    const int g = blockIdx.y;
    const long k0 = (long)blockIdx.z * RF_KT;
    const int npairs = N * (N + 1) / 2, J = (N + 1) * (N + 1);
    const bool pair = g < npairs;
    cplx ca[RF_KT], cb[RF_KT];
#pragma unroll
    for (int q = 0; q < RF_KT; ++q) { ca[q] = cmake(0.0, 0.0); cb[q] = cmake(0.0, 0.0); }
    for (long s0 = 0; s0 < n_dir; s0 += RF_SC) {
        __syncthreads();
        for (int e = threadIdx.x; e < RF_SC * RF_KT; e += 128) {
            const int kk = e / RF_SC, sl = e % RF_SC;           // consecutive threads: consecutive samples
            const long kq = k0 + kk, s = s0 + sl;
            Ps[sl][kk] = (kq < K && s < n_dir) ? Psh[kq * n_dir + s] : cmake(0.0, 0.0);
        }
        __syncthreads();
        const int ns = (int)(n_dir - s0 < RF_SC ? n_dir - s0 : RF_SC);
        if (pair) {
            const double2 *w = W2 + ((long)g * n_dir + s0) * I_pad + img;
#pragma unroll 2
            for (int sl = 0; sl < ns; ++sl) {
                const double2 wv = w[(long)sl * I_pad];
#pragma unroll
                for (int q = 0; q < RF_KT; ++q) {
                    const double2 pv = Ps[sl][q];
                    ca[q].x = fma(wv.x, pv.x, ca[q].x); ca[q].y = fma(wv.x, pv.y, ca[q].y);
                    cb[q].x = fma(wv.y, pv.x, cb[q].x); cb[q].y = fma(wv.y, pv.y, cb[q].y);
                }
            }
        } else {
            const double *w = W1 + ((long)(g - npairs) * n_dir + s0) * I_pad + img;
#pragma unroll 2
            for (int sl = 0; sl < ns; ++sl) {
                const double wv = w[(long)sl * I_pad];
#pragma unroll
                for (int q = 0; q < RF_KT; ++q) {
                    const double2 pv = Ps[sl][q];
                    ca[q].x = fma(wv, pv.x, ca[q].x); ca[q].y = fma(wv, pv.y, ca[q].y);
                }
            }
        }
    }
    if (img >= n_images) return;
    int n, m;
    if (pair) {
        n = 1;
        while (n * (n + 1) / 2 <= g) ++n;
        m = g - n * (n - 1) / 2 + 1;
    } else {
        n = g - npairs; m = 0;
    }
#pragma unroll
    for (int q = 0; q < RF_KT; ++q) {
        const long kq = k0 + q;
        if (kq >= K) break;
        const cplx hi = Hinv[(long)n * K + kq];
        if (m == 0) {
            const cplx c = cmul(ca[q], hi);
            out[(kq * J + (long)(n * n + n)) * n_images + img] = make_float2((float)c.x, (float)c.y);
        } else {
            // C_m = (c_a - i c_b)/2,   C_-m = (-1)^m (c_a + i c_b)/2
            const cplx cp = cmake(0.5 * (ca[q].x + cb[q].y), 0.5 * (ca[q].y - cb[q].x));
            const double sg = (m & 1) ? -0.5 : 0.5;
            const cplx cm = cmake(sg * (ca[q].x - cb[q].y), sg * (ca[q].y + cb[q].x));
            const cplx a = cmul(cp, hi), b = cmul(cm, hi);
            out[(kq * J + (long)(n * n + n + m)) * n_images + img] = make_float2((float)a.x, (float)a.y);
            out[(kq * J + (long)(n * n + n - m)) * n_images + img] = make_float2((float)b.x, (float)b.y);
        }
    }
}

}  // namespace deism

using namespace deism;

extern "C" int deism_arg_refit(int64_t n_images, int64_t K, int N, int64_t n_dir,
                               const float *d_reflection, const double *d_coords, const double *d_Psh,
                               const double *d_k, double r0, float *d_C_vec, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_images <= 0 || K <= 0) return DEISM_OK;
    if (N < 0 || N > RF_NMAX)
        return set_error(DEISM_ERR_INVALID, "source order %d outside 0..%d", N, RF_NMAX);
    const int J = (N + 1) * (N + 1), npairs = N * (N + 1) / 2;
    if (n_dir < J)
        return set_error(DEISM_ERR_INVALID, "%ld sampling directions cannot determine %d coefficients",
                         (long)n_dir, J);
    if (!d_reflection || !d_coords || !d_Psh || !d_k || !d_C_vec)
        return set_error(DEISM_ERR_INVALID, "null input pointer");
    if (!(r0 > 0.0)) return set_error(DEISM_ERR_INVALID, "radius r0 must be positive");

    const long I_pad = (n_images + 127) / 128 * 128;
    const size_t w2_bytes = (size_t)(npairs > 0 ? npairs : 1) * n_dir * I_pad * sizeof(double2);
    const size_t w1_bytes = (size_t)(N + 1) * n_dir * I_pad * sizeof(double);
    double2 *W2 = (double2 *)workspace(WS_ORG_B, w2_bytes);
    double *W1 = (double *)workspace(WS_ORG_Y, w1_bytes);
    cplx *Hinv = (cplx *)workspace(WS_SMALL, (size_t)(N + 1) * K * sizeof(cplx));
    int *status = (int *)workspace(WS_FLAGS, 64);
    if (!W2 || !W1 || !Hinv || !status) return DEISM_ERR_NOMEM;
    // padded image columns are read by the apply kernel: keep them finite
    DEISM_CUDA_TRY(cudaMemsetAsync(W2, 0, w2_bytes, st));
    DEISM_CUDA_TRY(cudaMemsetAsync(W1, 0, w1_bytes, st));
    DEISM_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));

    DEISM_CUDA_TRY(cudaFuncSetAttribute(refit_pinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(RefitSmem)));
    prof_begin("refit_pinv", st);
    refit_pinv_kernel<<<(unsigned)n_images, RF_THREADS, sizeof(RefitSmem), st>>>(
        n_images, I_pad, N, n_dir, d_reflection, d_coords, W2, W1, status);
    prof_end("refit_pinv", st);
    DEISM_LAUNCH_CHECK("refit_pinv_kernel");
    refit_hankel_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(K, N, d_k, r0, Hinv);
    DEISM_LAUNCH_CHECK("refit_hankel_kernel");
    dim3 grid((unsigned)(I_pad / 128), (unsigned)(npairs + N + 1), (unsigned)((K + RF_KT - 1) / RF_KT));
    prof_begin("refit_apply", st);
    refit_apply_kernel<<<grid, 128, 0, st>>>(n_images, I_pad, K, N, n_dir, W2, W1, (const cplx *)d_Psh,
                                             Hinv, (float2 *)d_C_vec);
    prof_end("refit_apply", st);
    DEISM_LAUNCH_CHECK("refit_apply_kernel");
    int h_status = 0;
    DEISM_CUDA_TRY(cudaMemcpyAsync(&h_status, status, sizeof(int), cudaMemcpyDeviceToHost, st));
    DEISM_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_status)
        return set_error(DEISM_ERR_INVALID,
                         "sampling directions do not determine an order-%d fit (normal matrix not positive definite)", N);
    return DEISM_OK;
}
