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
//     pseudo-inverse, in the operand layout of the contraction below;
//   * refit_apply_kernel contracts W with the sampled pressures on the FP64 tensor cores: a real
//     [64 images x n_dir] by complex [n_dir x 32 frequencies] product per (real-function pair,
//     image tile, frequency tile) as DMMA m8n8k4 GEMMs fed by a TMA ring, then maps real-basis
//     coefficients back to (n, +-m), divides by the spherical Hankel function and stores
//     complex64 along the image axis.
//
// R_i is the float32 matrix the reflection-tree engine produced; it is applied in double, as
// numpy's float32 @ float64 product does (cd:1577).
#include "common.cuh"

namespace deism {

constexpr int RF_NMAX = 11;                       // source order the shared-memory budget admits (J = 144)
constexpr int RF_CH = 128;                        // most samples per chunk (fewer at high orders)
constexpr int RF_THREADS = 256;
constexpr int RF_TI = 64, RF_TF = 32;             // apply kernel CTA tile: images x frequencies
constexpr int RF_SC = 24;                         // samples per staged chunk (multiple of 4)
constexpr int RF_WS = RF_TI + 2;                  // W row (doubles); a sample = 2 rows = 132 = 4 mod 16
constexpr int RF_PS = RF_TF + 2;                  // P plane (doubles); a sample = 2 planes = 68 = 4 mod 16
constexpr int RF_APPLY_THREADS = 256;             // 4 (image) x 2 (frequency) warps of 16 x 16

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

// Shared memory of refit_pinv_kernel: header, then G[J][J+1] (normal matrix, then its lower Cholesky
// factor) and y[ch][J+1] (real harmonics of a chunk of `ch` samples, one row per thread), sized for
// the order of the call.
struct RefitSmemHead {
    double R[9];
    int bad, pad;
};
static size_t refit_smem_bytes(int J, int ch) {
    return sizeof(RefitSmemHead) + (size_t)(J + ch) * (J + 1) * sizeof(double);
}

// W layout: [group][image tile][sample < n_dir_pad][a|b][RF_WS] doubles, image within the tile
// fastest; group g < npairs is the pair (a_m, b_m), g >= npairs the single a_0 of n = g - npairs
// (its b row stays zero).  Pre-zeroed: padded samples, images and rows contribute nothing.
__global__ void __launch_bounds__(RF_THREADS)
refit_pinv_kernel(long n_images, long n_itiles, int N, int ch, long n_dir, long n_dir_pad,
                  const float *__restrict__ refl, const double *__restrict__ X, double *W, int *status) {
    extern __shared__ __align__(16) unsigned char rf_smem_raw[];
    RefitSmemHead &sm = *reinterpret_cast<RefitSmemHead *>(rf_smem_raw);
    const long img = blockIdx.x;
    const int tid = threadIdx.x;
    const int J = (N + 1) * (N + 1), npairs = N * (N + 1) / 2;
    const int GS = J + 1;                                        // row stride of G and y
    double *G = reinterpret_cast<double *>(rf_smem_raw + sizeof(RefitSmemHead));
    double *Ych = G + (size_t)J * GS;
    if (tid < 9) sm.R[tid] = (double)refl[img * 9 + tid];
    if (tid == 0) sm.bad = 0;
    for (int e = tid; e < J * J; e += RF_THREADS) G[(e / J) * GS + e % J] = 0.0;
    __syncthreads();

    // ---- pass 1: G = sum_s y(s) y(s)^T ----
    for (long s0 = 0; s0 < n_dir; s0 += ch) {
        const int ns = (int)(n_dir - s0 < ch ? n_dir - s0 : ch);
        if (tid < ns) {
            double phi, ct;
            reflected_direction(sm.R, X, n_dir, s0 + tid, &phi, &ct);
            real_harmonics(N, phi, ct, Ych + (size_t)tid * GS, 1);
        }
        __syncthreads();
        for (int e = tid; e < J * J; e += RF_THREADS) {
            const int q = e / J, q2 = e % J;
            if (q2 > q) continue;
            double acc = 0.0;
            for (int t = 0; t < ns; ++t) acc = fma(Ych[t * GS + q], Ych[t * GS + q2], acc);
            G[q * GS + q2] += acc;
        }
        __syncthreads();
    }
    // ---- Cholesky G = L L^T (lower triangle in place), column by column ----
    for (int c = 0; c < J; ++c) {
        if (tid == 0) {
            double d = G[(c) * GS + (c)];
            for (int t = 0; t < c; ++t) d = fma(-G[(c) * GS + (t)], G[(c) * GS + (t)], d);
            if (!(d > 0.0)) { sm.bad = 1; d = 1.0; }
            G[(c) * GS + (c)] = sqrt(d);
        }
        __syncthreads();
        const double inv = 1.0 / G[(c) * GS + (c)];
        for (int r = c + 1 + tid; r < J; r += RF_THREADS) {
            double v = G[(r) * GS + (c)];
            for (int t = 0; t < c; ++t) v = fma(-G[(r) * GS + (t)], G[(c) * GS + (t)], v);
            G[(r) * GS + (c)] = v * inv;
        }
        __syncthreads();
    }
    if (sm.bad) {
        if (tid == 0) atomicExch(status, 1);      // rank-deficient sampling: reported by the host
        return;
    }
    // ---- pass 2: w(s) = G^-1 y(s), one sample per thread ----
    for (long s0 = 0; s0 < n_dir; s0 += ch) {
        const int ns = (int)(n_dir - s0 < ch ? n_dir - s0 : ch);
        if (tid < ns) {
            double phi, ct;
            reflected_direction(sm.R, X, n_dir, s0 + tid, &phi, &ct);
            double *y = Ych + (size_t)tid * GS;
            real_harmonics(N, phi, ct, y, 1);
            for (int r = 0; r < J; ++r) {                       // L z = y
                double v = y[r];
                for (int t = 0; t < r; ++t) v = fma(-G[(r) * GS + (t)], y[t], v);
                y[r] = v / G[(r) * GS + (r)];
            }
            for (int r = J - 1; r >= 0; --r) {                  // L^T w = z
                double v = y[r];
                for (int t = r + 1; t < J; ++t) v = fma(-G[(t) * GS + (r)], y[t], v);
                y[r] = v / G[(r) * GS + (r)];
            }
            const long s = s0 + tid;
            const long itile = img / RF_TI;
            const int it = (int)(img % RF_TI);
            for (int gq = 0; gq < npairs + N + 1; ++gq) {
                double *w = W + ((((long)gq * n_itiles + itile) * n_dir_pad + s) * 2) * RF_WS + it;
                if (gq < npairs) { w[0] = y[2 * gq]; w[RF_WS] = y[2 * gq + 1]; }
                else w[0] = y[2 * npairs + (gq - npairs)];
            }
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

// sampled pressures complex128 [K][n_dir] -> planes [K_pad/32][sample < n_dir_pad][Re|Im][RF_PS]
__global__ void __launch_bounds__(256)
refit_planes_kernel(long K, long n_dir, long n_dir_pad, long n_ktiles, const cplx *__restrict__ Psh,
                    double *PB) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = n_ktiles * n_dir_pad * RF_PS;
    if (i >= total) return;
    const int col = (int)(i % RF_PS);
    const long s = (i / RF_PS) % n_dir_pad, kt = i / (RF_PS * n_dir_pad);
    const long k = kt * RF_TF + col;
    cplx v = cmake(0.0, 0.0);
    if (col < RF_TF && k < K && s < n_dir) v = Psh[k * n_dir + s];
    double *p = PB + ((kt * n_dir_pad + s) * 2) * RF_PS + col;
    p[0] = v.x;
    p[RF_PS] = v.y;
}

struct RefitStage {
    double W[RF_SC][2][RF_WS];
    double P[RF_SC][2][RF_PS];
};
struct RefitApplySmem {
    RefitStage st[2];
    unsigned long long full[2];
    int done[2];
};

// grid (image tiles, frequency tiles, groups).  c_a[img,k] = sum_s W_a[img,s] P[k,s] (and c_b):
// per k-step of 4 samples 4 + 4 fragment loads feed 16 DMMAs; operand chunks arrive by TMA bulk
// copies, the last warp out of a stage refills it (no CTA barrier in the loop).
__global__ void __launch_bounds__(RF_APPLY_THREADS, 2)
refit_apply_kernel(long n_images, long n_itiles, long K, int N, long n_dir_pad, long n_ktiles,
                   const double *__restrict__ W, const double *__restrict__ PB,
                   const cplx *__restrict__ Hinv, float2 *out) {
    extern __shared__ __align__(128) unsigned char rfa_smem_raw[];
    RefitApplySmem &sm = *reinterpret_cast<RefitApplySmem *>(rfa_smem_raw);
    constexpr int NWARPS = RF_APPLY_THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3, wm = warp >> 1, wn = warp & 1;
    const long itile = blockIdx.x, kt = blockIdx.y;
    const int grp = blockIdx.z;
    const int npairs = N * (N + 1) / 2, J = (N + 1) * (N + 1);
    const int n_chunks = (int)(n_dir_pad / RF_SC);
    const double *Wg = W + (((long)grp * n_itiles + itile) * n_dir_pad) * 2 * RF_WS;
    const double *Pg = PB + (kt * n_dir_pad) * 2 * RF_PS;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(&sm.full[b], 1); sm.done[b] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c, int stage) {
        const unsigned wb = RF_SC * 2 * RF_WS * 8, pb = RF_SC * 2 * RF_PS * 8;
        mbar_expect_tx(&sm.full[stage], wb + pb);
        bulk_g2s(&sm.st[stage].W[0][0][0], Wg + (long)c * RF_SC * 2 * RF_WS, wb, &sm.full[stage]);
        bulk_g2s(&sm.st[stage].P[0][0][0], Pg + (long)c * RF_SC * 2 * RF_PS, pb, &sm.full[stage]);
    };
    if (tid == 0) {
        issue(0, 0);
        if (n_chunks > 1) issue(1, 1);
    }
    double T[2][2][2][2][2];      // [a|b][Re|Im][m-tile][n-tile][2]
#pragma unroll
    for (int i = 0; i < 32; ++i) (&T[0][0][0][0][0])[i] = 0.0;

#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
        const int stage = c & 1;
        mbar_wait(&sm.full[stage], (unsigned)((c >> 1) & 1));
        const double *Ww = &sm.st[stage].W[t4][0][wm * 16 + g];
        const double *Pw = &sm.st[stage].P[t4][0][wn * 16 + g];
#pragma unroll
        for (int ks = 0; ks < RF_SC; ks += 4) {
            double af[2][2], bf[2][2];
#pragma unroll
            for (int ab = 0; ab < 2; ++ab)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) af[ab][mt] = Ww[(ks * 2 + ab) * RF_WS + mt * 8];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) bf[p][nt] = Pw[(ks * 2 + p) * RF_PS + nt * 8];
#pragma unroll
            for (int ab = 0; ab < 2; ++ab)
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
                            dmma884(T[ab][p][mt][nt][0], T[ab][p][mt][nt][1], af[ab][mt], bf[p][nt]);
        }
        __syncwarp();
        if (lane == 0 && atomicAdd(&sm.done[stage], 1) == NWARPS - 1) {
            sm.done[stage] = 0;
            if (c + 2 < n_chunks) issue(c + 2, stage);
        }
    }

    // ---- epilogue: real-basis coefficients -> (n, +-m), / h_n, complex64 ----
    int n, m;
    if (grp < npairs) {
        n = 1;
        while (n * (n + 1) / 2 <= grp) ++n;
        m = grp - n * (n - 1) / 2 + 1;
    } else {
        n = grp - npairs; m = 0;
    }
    const double sg = (m & 1) ? -0.5 : 0.5;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const long img = itile * RF_TI + wm * 16 + mt * 8 + g;
        if (img >= n_images) continue;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long kq = kt * RF_TF + wn * 16 + nt * 8 + 2 * t4 + e;
                if (kq >= K) continue;
                const cplx hi = Hinv[(long)n * K + kq];
                const cplx ca = cmake(T[0][0][mt][nt][e], T[0][1][mt][nt][e]);
                const cplx cb = cmake(T[1][0][mt][nt][e], T[1][1][mt][nt][e]);
                if (m == 0) {
                    const cplx v = cmul(ca, hi);
                    out[(kq * J + (long)(n * n + n)) * n_images + img] = make_float2((float)v.x, (float)v.y);
                } else {
                    // C_m = (c_a - i c_b)/2,   C_-m = (-1)^m (c_a + i c_b)/2
                    const cplx cp = cmake(0.5 * (ca.x + cb.y), 0.5 * (ca.y - cb.x));
                    const cplx cm = cmake(sg * (ca.x - cb.y), sg * (ca.y + cb.x));
                    const cplx va = cmul(cp, hi), vb = cmul(cm, hi);
                    out[(kq * J + (long)(n * n + n + m)) * n_images + img] = make_float2((float)va.x, (float)va.y);
                    out[(kq * J + (long)(n * n + n - m)) * n_images + img] = make_float2((float)vb.x, (float)vb.y);
                }
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
    CallGuard call_guard(stream);
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

    const long n_itiles = (n_images + RF_TI - 1) / RF_TI, n_ktiles = (K + RF_TF - 1) / RF_TF;
    const long n_dir_pad = (n_dir + RF_SC - 1) / RF_SC * RF_SC;
    const int n_groups = npairs + N + 1;
    const size_t w_bytes = (size_t)n_groups * n_itiles * n_dir_pad * 2 * RF_WS * sizeof(double);
    const size_t p_bytes = (size_t)n_ktiles * n_dir_pad * 2 * RF_PS * sizeof(double);
    double *W = (double *)workspace(WS_ORG_B, w_bytes);
    double *PB = (double *)workspace(WS_ORG_Y, p_bytes);
    cplx *Hinv = (cplx *)workspace(WS_SMALL, (size_t)(N + 1) * K * sizeof(cplx));
    int *status = (int *)workspace(WS_FLAGS, 64);
    if (!W || !PB || !Hinv || !status) return DEISM_ERR_NOMEM;
    // padded samples / images / the b row of the single groups must read as zero
    DEISM_CUDA_TRY(cudaMemsetAsync(W, 0, w_bytes, st));
    DEISM_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));

    // samples per chunk: as many as the shared memory left by G admits (two CTAs per SM up to order 7)
    int ch = RF_CH;
    while (ch > 32 && refit_smem_bytes(J, ch) > (size_t)smem_optin_bytes() / (N <= 7 ? 2 : 1) - 1024) ch /= 2;
    const size_t pinv_smem = refit_smem_bytes(J, ch);
    if (pinv_smem > (size_t)smem_optin_bytes())
        return set_error(DEISM_ERR_INVALID, "source order %d needs %zu B of shared memory", N, pinv_smem);
    DEISM_CUDA_TRY(cudaFuncSetAttribute(refit_pinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)pinv_smem));
    prof_begin("refit_pinv", st);
    refit_pinv_kernel<<<(unsigned)n_images, RF_THREADS, pinv_smem, st>>>(
        n_images, n_itiles, N, ch, n_dir, n_dir_pad, d_reflection, d_coords, W, status);
    prof_end("refit_pinv", st);
    DEISM_LAUNCH_CHECK("refit_pinv_kernel");
    refit_hankel_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(K, N, d_k, r0, Hinv);
    DEISM_LAUNCH_CHECK("refit_hankel_kernel");
    refit_planes_kernel<<<(unsigned)((n_ktiles * n_dir_pad * RF_PS + 255) / 256), 256, 0, st>>>(
        K, n_dir, n_dir_pad, n_ktiles, (const cplx *)d_Psh, PB);
    DEISM_LAUNCH_CHECK("refit_planes_kernel");
    if (n_ktiles > 65535 || n_groups > 65535)
        return set_error(DEISM_ERR_INVALID, "too many frequency tiles (%ld) for one launch", n_ktiles);
    dim3 grid((unsigned)n_itiles, (unsigned)n_ktiles, (unsigned)n_groups);
    DEISM_CUDA_TRY(cudaFuncSetAttribute(refit_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(RefitApplySmem)));
    prof_begin("refit_apply", st);
    refit_apply_kernel<<<grid, RF_APPLY_THREADS, sizeof(RefitApplySmem), st>>>(
        n_images, n_itiles, K, N, n_dir_pad, n_ktiles, W, PB, Hinv, (float2 *)d_C_vec);
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
