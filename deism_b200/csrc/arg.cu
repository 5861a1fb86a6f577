// arg.cu — convex-room DEISM-ARG LC accumulation.
//
// Replaces _numba_ARG_LC_batch (deism/parallel_backends.py:328-385):
//   P[k] = sum_img -atten[k,img] * 4pi/k * exp(-i k r)/(k r) * S * R
//   S = sum_j i^{-n_j} (-1)^{n_j} C_s[k,j,img] Y_{n_j}^{m_j}(theta,phi)
//   R = sum_j i^{v_j}  (-1)^{v_j} C_r[k,j]     Y_{v_j}^{u_j}(theta,phi)
// with ONE direction (phi,theta,r) = R_sI_r[:,img] for both sums and per-image source
// coefficients C_s[K,Js,I] (image index innermost, cd:1193-1217).
//
// The per-image coefficients make this a batched dot product, not a GEMM: every C_s element
// is used once, so the kernel is bound by streaming C_s (8 B per (k,j,img)) from HBM.
// Mapping: thread = image (lanes run along the innermost, contiguous image axis -> coalesced
// 128-byte loads), blockIdx.y = frequency; Y tables [J][I] stay in L2.
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace deism {

int launch_reduce_partials(const cplx *partial, long n_chunks, long K, double *d_out, int accumulate,
                           cudaStream_t st);

constexpr int ARG_JMAX = 256;
__constant__ int c_arg_n[ARG_JMAX], c_arg_m[ARG_JMAX], c_arg_v[ARG_JMAX], c_arg_u[ARG_JMAX];

struct ArgPrepArgs {
    long n_images;
    int Js, Jr;
    const float *R_sI_r;   // [3][I]
};

// Ys'[j][img] = i^{-n}(-1)^n Y = i^n Y ;  Yr'[j][img] = i^v (-1)^v Y = (-i)^v Y
__global__ void __launch_bounds__(128)
arg_prep_kernel(ArgPrepArgs a, cplx *Ys, cplx *Yr, double *r_out) {
    long img = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= a.n_images) return;
    double phi = (double)a.R_sI_r[img], theta = (double)a.R_sI_r[a.n_images + img];
    r_out[img] = (double)a.R_sI_r[2 * a.n_images + img];
    double st, ct;
    sincos_cw(theta, &st, &ct);
    for (int j = 0; j < a.Js; ++j)
        Ys[(long)j * a.n_images + img] =
            cmul_ipow(sph_harm_dev(c_arg_m[j], c_arg_n[j], phi, ct), c_arg_n[j]);
    for (int j = 0; j < a.Jr; ++j)
        Yr[(long)j * a.n_images + img] =
            cmul_ipow(sph_harm_dev(c_arg_u[j], c_arg_v[j], phi, ct), 3 * c_arg_v[j]);
}

struct ArgMainArgs {
    long n_sel, n_images, K;
    int Js, Jr;
    const float2 *Cs;     // [K][Js][I]
    const float2 *Cr;     // [K][Jr]
    const cplx *Ys, *Yr;  // [J][I]
    const double *r;      // [I]
    const float2 *atten;  // [K][I]
    const double *k;
    const long *sel;      // [n_sel] or nullptr
    cplx *partial;        // [gridDim.x][K]
};

// thread = image, blockIdx.y = a tile of ARG_KT frequencies: the image's harmonics Y_j are read
// once per tile and reused from registers for every frequency of the tile (re-reading them per
// frequency cost twice the traffic of the C_s stream itself), ARG_KT independent C_s loads are in
// flight per harmonic.
constexpr int ARG_KT = 8;

__global__ void __launch_bounds__(128, 4) arg_lc_kernel(ArgMainArgs a) {
    __shared__ double2 red[4][ARG_KT];
    __shared__ double2 cr_sm[ARG_KT][ARG_JMAX];
    const long k0 = (long)blockIdx.y * ARG_KT;
    for (int e = threadIdx.x; e < ARG_KT * a.Jr; e += 128) {
        const int q = e / a.Jr, j = e % a.Jr;
        float2 c = k0 + q < a.K ? a.Cr[(k0 + q) * a.Jr + j] : make_float2(0.f, 0.f);
        cr_sm[q][j] = cmake((double)c.x, (double)c.y);
    }
    __syncthreads();
    const long s = (long)blockIdx.x * 128 + threadIdx.x;
    cplx P[ARG_KT];
#pragma unroll
    for (int q = 0; q < ARG_KT; ++q) P[q] = cmake(0.0, 0.0);
    if (s < a.n_sel) {
        const long img = a.sel ? a.sel[s] : s;
        const double r = a.r[img];
        cplx S[ARG_KT], R[ARG_KT];
#pragma unroll
        for (int q = 0; q < ARG_KT; ++q) { S[q] = cmake(0.0, 0.0); R[q] = cmake(0.0, 0.0); }
        const long kstride = (long)a.Js * a.n_images;
        const float2 *Cs = a.Cs + k0 * kstride + img;
#pragma unroll 2
        for (int j = 0; j < a.Js; ++j) {
            const cplx y = a.Ys[(long)j * a.n_images + img];
            float2 c[ARG_KT];
#pragma unroll
            for (int q = 0; q < ARG_KT; ++q)
                c[q] = k0 + q < a.K ? __ldcs(&Cs[q * kstride + (long)j * a.n_images]) : make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < ARG_KT; ++q) cfma(S[q], cmake((double)c[q].x, (double)c[q].y), y);
        }
        for (int j = 0; j < a.Jr; ++j) {
            const cplx y = a.Yr[(long)j * a.n_images + img];
#pragma unroll
            for (int q = 0; q < ARG_KT; ++q) cfma(R[q], cr_sm[q][j], y);
        }
#pragma unroll
        for (int q = 0; q < ARG_KT; ++q) {
            if (k0 + q >= a.K) break;
            const double kk = a.k[k0 + q];
            double sn, cs;
            sincos_cw(kk * r, &sn, &cs);
            const float2 at = a.atten[(k0 + q) * a.n_images + img];
            const double w = -4.0 * 3.14159265358979323846 / (kk * kk * r);
            const cplx fac = cmul(cmake((double)at.x, (double)at.y), cmake(cs * w, -sn * w));
            P[q] = cmul(cmul(fac, S[q]), R[q]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < ARG_KT; ++q) {
        double x = P[q].x, y = P[q].y;
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_down_sync(0xffffffffu, x, o);
            y += __shfl_down_sync(0xffffffffu, y, o);
        }
        if (lane == 0) red[warp][q] = cmake(x, y);
    }
    __syncthreads();
    if (threadIdx.x < ARG_KT && k0 + threadIdx.x < a.K) {
        const int q = threadIdx.x;
        a.partial[(long)blockIdx.x * a.K + k0 + q] =
            cadd(cadd(red[0][q], red[1][q]), cadd(red[2][q], red[3][q]));
    }
}

}  // namespace deism

using namespace deism;

extern "C" int deism_lc_arg(int64_t n_images, int64_t K, int Js, int Jr, const int64_t *h_n_all,
                            const int64_t *h_m_all, const int64_t *h_v_all, const int64_t *h_u_all,
                            const float *d_C_s_ARG_vec, const float *d_C_r_vec, const float *d_R_sI_r,
                            const float *d_atten, const double *d_k, const int64_t *h_image_index,
                            int64_t n_sel, double *d_out, int accumulate, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 0) return DEISM_OK;
    if (!d_out || !d_k) return set_error(DEISM_ERR_INVALID, "null d_out/d_k");
    if (!h_image_index) n_sel = n_images;
    if (n_images <= 0 || n_sel <= 0) {
        if (!accumulate) DEISM_CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)K * 16, st));
        return DEISM_OK;
    }
    if (Js <= 0 || Jr <= 0 || Js > ARG_JMAX || Jr > ARG_JMAX)
        return set_error(DEISM_ERR_INVALID, "Js/Jr out of range (1..%d)", ARG_JMAX);
    if (!h_n_all || !h_m_all || !h_v_all || !h_u_all || !d_C_s_ARG_vec || !d_C_r_vec || !d_R_sI_r ||
        !d_atten)
        return set_error(DEISM_ERR_INVALID, "null input pointer");
    int hn[ARG_JMAX], hm[ARG_JMAX], hv[ARG_JMAX], hu[ARG_JMAX];
    for (int j = 0; j < Js; ++j) {
        hn[j] = (int)h_n_all[j]; hm[j] = (int)h_m_all[j];
        if (hn[j] < 0 || hn[j] > 64 || abs(hm[j]) > hn[j])
            return set_error(DEISM_ERR_INVALID, "bad (n,m) at j=%d", j);
    }
    for (int j = 0; j < Jr; ++j) {
        hv[j] = (int)h_v_all[j]; hu[j] = (int)h_u_all[j];
        if (hv[j] < 0 || hv[j] > 64 || abs(hu[j]) > hv[j])
            return set_error(DEISM_ERR_INVALID, "bad (v,u) at j=%d", j);
    }
    bool tables_uploaded = false;
    {   // (n,m,v,u) live in constant memory: upload only when they change (no host sync otherwise)
        static std::mutex mu;
        static int last[4][ARG_JMAX], last_js = -1, last_jr = -1, last_dev = -1;
        int dev = -1;
        DEISM_CUDA_TRY(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (dev != last_dev || Js != last_js || Jr != last_jr || memcmp(last[0], hn, Js * sizeof(int)) ||
            memcmp(last[1], hm, Js * sizeof(int)) || memcmp(last[2], hv, Jr * sizeof(int)) ||
            memcmp(last[3], hu, Jr * sizeof(int))) {
            last_dev = -1;
            bump_generation();
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_arg_n, hn, Js * sizeof(int), 0, cudaMemcpyHostToDevice, st));
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_arg_m, hm, Js * sizeof(int), 0, cudaMemcpyHostToDevice, st));
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_arg_v, hv, Jr * sizeof(int), 0, cudaMemcpyHostToDevice, st));
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_arg_u, hu, Jr * sizeof(int), 0, cudaMemcpyHostToDevice, st));
            memcpy(last[0], hn, Js * sizeof(int)); memcpy(last[1], hm, Js * sizeof(int));
            memcpy(last[2], hv, Jr * sizeof(int)); memcpy(last[3], hu, Jr * sizeof(int));
            last_js = Js; last_jr = Jr; last_dev = dev;
            tables_uploaded = true;
        }
    }

    cplx *Ys = (cplx *)workspace(WS_LC_YS, (size_t)n_images * Js * sizeof(cplx));
    cplx *Yr = (cplx *)workspace(WS_LC_YR, (size_t)n_images * Jr * sizeof(cplx));
    double *r = (double *)workspace(WS_LC_IMG, (size_t)n_images * sizeof(double));
    long nblk = (n_sel + 127) / 128;
    cplx *partial = (cplx *)workspace(WS_PARTIAL, (size_t)nblk * K * sizeof(cplx));
    long *sel = nullptr;
    if (h_image_index) {
        for (int64_t i = 0; i < n_sel; ++i)
            if (h_image_index[i] < 0 || h_image_index[i] >= n_images)
                return set_error(DEISM_ERR_INVALID, "image index %lld out of range",
                                 (long long)h_image_index[i]);
        sel = (long *)workspace(WS_ARG_IDX, (size_t)n_sel * sizeof(long));
        if (!sel) return DEISM_ERR_NOMEM;
        DEISM_CUDA_TRY(cudaMemcpyAsync(sel, h_image_index, (size_t)n_sel * sizeof(long),
                                       cudaMemcpyHostToDevice, st));
    }
    if (tables_uploaded || sel)
        DEISM_CUDA_TRY(cudaStreamSynchronize(st));   // stack / caller host buffers consumed
    if (!Ys || !Yr || !r || !partial) return DEISM_ERR_NOMEM;

    ArgPrepArgs pa;
    pa.n_images = n_images; pa.Js = Js; pa.Jr = Jr; pa.R_sI_r = d_R_sI_r;
    arg_prep_kernel<<<(unsigned)((n_images + 127) / 128), 128, 0, st>>>(pa, Ys, Yr, r);
    DEISM_LAUNCH_CHECK("arg_prep_kernel");
    ArgMainArgs ma;
    ma.n_sel = n_sel; ma.n_images = n_images; ma.K = K; ma.Js = Js; ma.Jr = Jr;
    ma.Cs = (const float2 *)d_C_s_ARG_vec; ma.Cr = (const float2 *)d_C_r_vec; ma.Ys = Ys; ma.Yr = Yr;
    ma.r = r; ma.atten = (const float2 *)d_atten; ma.k = d_k; ma.sel = sel; ma.partial = partial;
    if ((K + ARG_KT - 1) / ARG_KT > 65535)
        return set_error(DEISM_ERR_INVALID, "ARG LC supports K <= %d", 65535 * ARG_KT);
    dim3 grid((unsigned)nblk, (unsigned)((K + ARG_KT - 1) / ARG_KT));
    prof_begin("arg_lc", st);
    arg_lc_kernel<<<grid, 128, 0, st>>>(ma);
    prof_end("arg_lc", st);
    DEISM_LAUNCH_CHECK("arg_lc_kernel");
    return launch_reduce_partials(partial, nblk, K, d_out, accumulate, st);
}
