/*
 * oracle/deism_oracle.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C (C11 + OpenMP) restatement of the CPU algorithm behind
 * deism.core_deism.DEISM.run_DEISM() of audiolabs/DEISM 2.2.1.14, i.e. the Numba kernels of
 * deism/parallel_backends.py and the Numba shoebox image generator of deism/core_deism.py.
 * It is the CHECKER for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference legs).  The product (deism_b200/) never links or calls it.
 *
 * Parity pinning: the reference holds no golden vectors for this path, so this
 * file is pinned against outputs of the reference itself, run in the build container from
 * /root/reference by oracle/make_golden.py and frozen under tests/golden/ (see
 * tests/test_oracle_golden.py: image lists / f32 geometry / c64 attenuation bit-equal,
 * complex128 RTF to <= 1e-12 relative).
 *
 * Numerics rules followed:
 *   - every arithmetic statement keeps the reference's operation order; compiled with
 *     -ffp-contract=off because Numba/LLVM does not fuse a*b+c;
 *   - complex division follows the Smith-type algorithm Numba borrows from CPython's
 *     _Py_c_quot (branch on |re(b)| >= |im(b)|);
 *   - libm (glibc) sin/cos/exp/sqrt/hypot/atan2 — the same library Numba's LLVM lowers to;
 *   - parallel loop over images, per-image rows written to P_all, then a SERIAL sum over
 *     images in index order (parallel_backends.py:254-261), exactly as the reference.
 *
 * All arrays are C-contiguous; complex128 = interleaved (re, im) doubles, complex64 =
 * interleaved floats.  "pb" = deism/parallel_backends.py, "cd" = deism/core_deism.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

typedef struct { double re, im; } cplx;
typedef struct { float re, im; } cplxf;

static inline cplx c_make(double re, double im) { cplx z = { re, im }; return z; }
static inline cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
static inline cplx c_mul(cplx a, cplx b) {
    return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
static inline cplx c_scale(cplx a, double s) { return c_make(a.re * s, a.im * s); }
static inline cplx c_conj(cplx a) { return c_make(a.re, -a.im); }

/* Numba lowers complex/complex to CPython's _Py_c_quot scheme. */
static inline cplx c_div(cplx a, cplx b) {
    double abs_br = fabs(b.re), abs_bi = fabs(b.im);
    if (abs_br >= abs_bi) {
        if (abs_br == 0.0) return c_make(NAN, NAN);
        double ratio = b.im / b.re;
        double denom = b.re + b.im * ratio;
        return c_make((a.re + a.im * ratio) / denom, (a.im - a.re * ratio) / denom);
    } else {
        double ratio = b.re / b.im;
        double denom = b.re * ratio + b.im;
        return c_make((a.re * ratio + a.im) / denom, (a.im * ratio - a.re) / denom);
    }
}

/* i^n for any integer n (pb:234 "(1j) ** l", pb:283, pb:347) */
static inline cplx c_ipow(long n) {
    switch (((n % 4) + 4) % 4) {
    case 0: return c_make(1.0, 0.0);
    case 1: return c_make(0.0, 1.0);
    case 2: return c_make(-1.0, 0.0);
    default: return c_make(0.0, -1.0);
    }
}
static inline double minus_one_pow(long n) { return (n % 2 == 0) ? 1.0 : -1.0; }

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------------------------------
 * pb:40-83  _sph_harm_numba(m, n, phi, theta): scipy-convention Y_n^m, theta = polar angle.
 * ------------------------------------------------------------------------------------- */
static cplx sph_harm(long m, long n, double phi, double theta) {
    double x = cos(theta);
    long am = m < 0 ? -m : m;
    double pmm = 1.0;
    if (am > 0) {
        double somx2 = sqrt(1.0 - x * x);
        double fact = 1.0;
        for (long i = 1; i <= am; ++i) {
            pmm *= -fact * somx2;
            fact += 2.0;
        }
    }
    double p_val;
    if (n == am) {
        p_val = pmm;
    } else if (n == am + 1) {
        p_val = x * (double)(2 * am + 1) * pmm;
    } else {
        double pmm1 = x * (double)(2 * am + 1) * pmm;
        for (long ll = am + 2; ll <= n; ++ll) {
            double pll = (x * (double)(2 * ll - 1) * pmm1 - (double)(ll + am - 1) * pmm)
                         / (double)(ll - am);
            pmm = pmm1;
            pmm1 = pll;
        }
        p_val = pmm1;
    }
    double ratio = 1.0;
    for (long i = n - am + 1; i <= n + am; ++i) ratio *= (double)i;
    ratio = 1.0 / ratio;
    double norm = sqrt((double)(2 * n + 1) / (4.0 * M_PI) * ratio);
    double ang = (double)am * phi;
    double np_ = norm * p_val;
    cplx pos = c_make(np_ * cos(ang), np_ * sin(ang));
    if (m < 0) {
        double sign = (am % 2 == 0) ? 1.0 : -1.0;
        return c_scale(c_conj(pos), sign);
    }
    return pos;
}

ORACLE_API void oracle_sph_harm(long m, long n, double phi, double theta, double *out2) {
    cplx y = sph_harm(m, n, phi, theta);
    out2[0] = y.re; out2[1] = y.im;
}

/* ---------------------------------------------------------------------------------------
 * pb:86-118  _sphankel2_numba(n, kr) = j_n(kr) - i y_n(kr), upward recurrences.
 * ------------------------------------------------------------------------------------- */
static cplx sphankel2(long n, double kr) {
    if (kr == 0.0) return c_make(NAN, NAN);
    double s = sin(kr), c = cos(kr);
    double jn, yn;
    if (n == 0) {
        jn = s / kr;
        yn = -c / kr;
    } else if (n == 1) {
        jn = s / (kr * kr) - c / kr;
        yn = -c / (kr * kr) - s / kr;
    } else {
        double j0 = s / kr, j1 = s / (kr * kr) - c / kr;
        for (long ll = 2; ll <= n; ++ll) {
            double jnew = (double)(2 * ll - 1) / kr * j1 - j0;
            j0 = j1; j1 = jnew;
        }
        jn = j1;
        double y0 = -c / kr, y1 = -c / (kr * kr) - s / kr;
        for (long ll = 2; ll <= n; ++ll) {
            double ynew = (double)(2 * ll - 1) / kr * y1 - y0;
            y0 = y1; y1 = ynew;
        }
        yn = y1;
    }
    return c_make(jn, -yn);
}

ORACLE_API void oracle_sphankel2(long n, double kr, double *out2) {
    cplx h = sphankel2(n, kr);
    out2[0] = h.re; out2[1] = h.im;
}

/* ---------------------------------------------------------------------------------------
 * Shoebox wall reflection factor and integer power
 * pb:121-129 / cd:2634-2646:  beta = (zeta*cos - 1)/(zeta*cos + 1);  base**e by repeated
 * multiplication starting at 1+0j.
 * ------------------------------------------------------------------------------------- */
static inline cplx ref_coef(double cos_theta, cplx zeta) {
    cplx zc = c_make(zeta.re * cos_theta, zeta.im * cos_theta);
    return c_div(c_make(zc.re - 1.0, zc.im), c_make(zc.re + 1.0, zc.im));
}
static inline cplx cpow_int(cplx base, long e) {
    cplx out = c_make(1.0, 0.0);
    for (long i = 0; i < e; ++i) out = c_mul(out, base);
    return out;
}
static inline cplx shoebox_atten(const cplx *Z, long K, long ki, double cx, double cy,
                                 double cz, const long e[6]) {
    cplx bx1 = ref_coef(cx, Z[0 * K + ki]);
    cplx bx2 = ref_coef(cx, Z[1 * K + ki]);
    cplx by1 = ref_coef(cy, Z[2 * K + ki]);
    cplx by2 = ref_coef(cy, Z[3 * K + ki]);
    cplx bz1 = ref_coef(cz, Z[4 * K + ki]);
    cplx bz2 = ref_coef(cz, Z[5 * K + ki]);
    cplx a = cpow_int(bx1, e[0]);
    a = c_mul(a, cpow_int(bx2, e[1]));
    a = c_mul(a, cpow_int(by1, e[2]));
    a = c_mul(a, cpow_int(by2, e[3]));
    a = c_mul(a, cpow_int(bz1, e[4]));
    a = c_mul(a, cpow_int(bz2, e[5]));
    return a;
}
static inline void shoebox_exponents(long qx, long qy, long qz, long px, long py, long pz,
                                     long e[6]) {
    e[0] = labs(qx - px); e[1] = labs(qx);
    e[2] = labs(qy - py); e[3] = labs(qy);
    e[4] = labs(qz - pz); e[5] = labs(qz);
}

/* pb:140-186  _numba_build_shoebox_attenuation_batch ("compact" image storage):
 * direction cosines from the (f32-rounded, up-cast) angles of R_sI_r (pb:131-137). */
ORACLE_API void oracle_build_shoebox_attenuation_batch(
    const int64_t *A, const double *R_sI_r, const double *Z_S, int ang_dep_flag,
    long n_images, long K, double *atten_out)
{
    const cplx *Z = (const cplx *)Z_S;
    cplx *out = (cplx *)atten_out;
#pragma omp parallel for schedule(static)
    for (long img = 0; img < n_images; ++img) {
        long e[6];
        shoebox_exponents(A[img * 6 + 0], A[img * 6 + 1], A[img * 6 + 2],
                          A[img * 6 + 3], A[img * 6 + 4], A[img * 6 + 5], e);
        double cx = 1.0, cy = 1.0, cz = 1.0;
        if (ang_dep_flag == 1) {
            double phi = R_sI_r[img * 3 + 0], theta = R_sI_r[img * 3 + 1];
            double st = sin(theta);
            cx = fabs(st * cos(phi));
            cy = fabs(st * sin(phi));
            cz = fabs(cos(theta));
        }
        for (long ki = 0; ki < K; ++ki)
            out[img * K + ki] = shoebox_atten(Z, K, ki, cx, cy, cz, e);
    }
}

/* ---------------------------------------------------------------------------------------
 * pb:189-261  _numba_ORG_batch  (shoebox DEISM-ORG, direct quintuple sum)
 * W_2 is indexed [n, v, l, m', u] with Python negative-index wrap on the last two axes
 * (pb:228), C_nm_s[k, n, m] and C_vu_r[k, v, -u] likewise (pb:249-250).
 * ------------------------------------------------------------------------------------- */
static inline long wrap(long i, long n) { return i < 0 ? i + n : i; }

static void org_one_image(
    long N, long V, const cplx *C_s, const cplx *C_r, const cplx *W1, const cplx *W2,
    const double *k, long K, double phi, double theta, double r,
    long px, long py, long pz, int shoebox_mirror,
    const cplx *atten, long atten_stride,          /* atten[ki*atten_stride] */
    const cplx *C_s_img, long cs_k_stride, long cs_n_stride, long cs_m_stride,
    cplx *P_img, cplx *sphan2, cplx *local_sum)
{
    const long Lmax = N + V + 1, M2 = 2 * N + 1, U2 = 2 * V + 1;
    (void)C_s;
    for (long l = 0; l < Lmax; ++l)
        for (long ki = 0; ki < K; ++ki)
            sphan2[l * K + ki] = sphankel2(l, k[ki] * r);
    for (long ki = 0; ki < K; ++ki) P_img[ki] = c_make(0.0, 0.0);

    for (long n = 0; n <= N; ++n)
    for (long m = -n; m <= n; ++m) {
        double mirror = 1.0;
        long m_mod = m;
        if (shoebox_mirror) {
            mirror = minus_one_pow((py + pz) * m + pz * n);
            m_mod = ((px + py) % 2 == 0) ? m : -m;
        }
        for (long v = 0; v <= V; ++v)
        for (long u = -v; u <= v; ++u) {
            for (long ki = 0; ki < K; ++ki) local_sum[ki] = c_make(0.0, 0.0);
            long lo = labs(n - v);
            for (long l = lo; l <= n + v; ++l) {
                if (labs(u - m_mod) > l) continue;
                cplx w1 = W1[(n * (V + 1) + v) * Lmax + l];
                cplx w2 = W2[((((n * (V + 1) + v) * Lmax + l) * M2) + wrap(m_mod, M2)) * U2
                             + wrap(u, U2)];
                if ((w1.re == 0.0 && w1.im == 0.0) || (w2.re == 0.0 && w2.im == 0.0)) continue;
                double Xi = sqrt((double)((2 * n + 1) * (2 * v + 1) * (2 * l + 1)) / (4.0 * M_PI));
                cplx Ylm = sph_harm(m_mod - u, l, phi, theta);
                cplx il = c_ipow(l);
                for (long ki = 0; ki < K; ++ki) {
                    /* il * h * Ylm * w1 * w2 * Xi, left to right (pb:236-238) */
                    cplx t = c_mul(il, sphan2[l * K + ki]);
                    t = c_mul(t, Ylm);
                    t = c_mul(t, w1);
                    t = c_mul(t, w2);
                    t = c_scale(t, Xi);
                    local_sum[ki] = c_add(local_sum[ki], t);
                }
            }
            /* 4*pi * i^(v-n) * (-1)^m'  (pb:240) */
            cplx Sf = c_scale(c_scale(c_ipow(v - n), 4.0 * M_PI), minus_one_pow(m_mod));
            double sign_u = minus_one_pow(u);
            for (long ki = 0; ki < K; ++ki) {
                cplx S = c_mul(Sf, local_sum[ki]);
                cplx cs = C_s_img[ki * cs_k_stride + n * cs_n_stride + wrap(m, M2) * cs_m_stride];
                cplx cr = C_r[(ki * (V + 1) + v) * U2 + wrap(-u, U2)];
                /* mirror * atten * C_s * S * C_r * 1j / k * sign_u (pb:247-252) */
                cplx t = c_scale(atten[ki * atten_stride], mirror);
                t = c_mul(t, cs);
                t = c_mul(t, S);
                t = c_mul(t, cr);
                t = c_mul(t, c_make(0.0, 1.0));
                t = c_make(t.re / k[ki], t.im / k[ki]);
                t = c_scale(t, sign_u);
                P_img[ki] = c_add(P_img[ki], t);
            }
        }
    }
}

ORACLE_API int oracle_ORG_batch(
    long N, long V, const double *C_nm_s, const double *C_vu_r, const int64_t *A,
    const double *atten, const double *x0, const double *W_1, const double *W_2,
    const double *k, long n_images, long K, double *result)
{
    const long Lmax = N + V + 1, M2 = 2 * N + 1;
    cplx *P_all = (cplx *)calloc((size_t)n_images * K, sizeof(cplx));
    if (!P_all) return -1;
    int fail = 0;
#pragma omp parallel
    {
        cplx *sphan2 = (cplx *)malloc((size_t)Lmax * K * sizeof(cplx));
        cplx *local_sum = (cplx *)malloc((size_t)K * sizeof(cplx));
        if (!sphan2 || !local_sum) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp barrier
        if (!fail) {
#pragma omp for schedule(dynamic, 1)
            for (long img = 0; img < n_images; ++img) {
                org_one_image(N, V, NULL, (const cplx *)C_vu_r, (const cplx *)W_1,
                              (const cplx *)W_2, k, K,
                              x0[img * 3 + 0], x0[img * 3 + 1], x0[img * 3 + 2],
                              A[img * 6 + 3], A[img * 6 + 4], A[img * 6 + 5], 1,
                              (const cplx *)atten + img * K, 1,
                              (const cplx *)C_nm_s, (N + 1) * M2, M2, 1,
                              P_all + img * K, sphan2, local_sum);
            }
        }
        free(sphan2); free(local_sum);
    }
    if (!fail) {
        cplx *res = (cplx *)result;
        for (long ki = 0; ki < K; ++ki) res[ki] = c_make(0.0, 0.0);
        for (long img = 0; img < n_images; ++img)
            for (long ki = 0; ki < K; ++ki) res[ki] = c_add(res[ki], P_all[img * K + ki]);
    }
    free(P_all);
    return fail ? -1 : 0;
}

/* pb:388-460  _numba_ARG_ORG_batch: no mirror / m', C_nm_s_ARG[k, n, m, img],
 * atten[k, img], R_sI_r[3, img]. */
ORACLE_API int oracle_ARG_ORG_batch(
    long N, long V, const double *C_nm_s_ARG, const double *C_vu_r, const double *atten,
    const double *R_sI_r, const double *W_1, const double *W_2, const double *k,
    long n_images, long K, double *result)
{
    const long Lmax = N + V + 1, M2 = 2 * N + 1;
    cplx *P_all = (cplx *)calloc((size_t)n_images * K, sizeof(cplx));
    if (!P_all) return -1;
    int fail = 0;
#pragma omp parallel
    {
        cplx *sphan2 = (cplx *)malloc((size_t)Lmax * K * sizeof(cplx));
        cplx *local_sum = (cplx *)malloc((size_t)K * sizeof(cplx));
        if (!sphan2 || !local_sum) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp barrier
        if (!fail) {
#pragma omp for schedule(dynamic, 1)
            for (long img = 0; img < n_images; ++img) {
                org_one_image(N, V, NULL, (const cplx *)C_vu_r, (const cplx *)W_1,
                              (const cplx *)W_2, k, K,
                              R_sI_r[0 * n_images + img], R_sI_r[1 * n_images + img],
                              R_sI_r[2 * n_images + img], 0, 0, 0, 0,
                              (const cplx *)atten + img, n_images,
                              (const cplx *)C_nm_s_ARG + img,
                              (N + 1) * M2 * n_images, M2 * n_images, n_images,
                              P_all + img * K, sphan2, local_sum);
            }
        }
        free(sphan2); free(local_sum);
    }
    if (!fail) {
        cplx *res = (cplx *)result;
        for (long ki = 0; ki < K; ++ki) res[ki] = c_make(0.0, 0.0);
        for (long img = 0; img < n_images; ++img)
            for (long ki = 0; ki < K; ++ki) res[ki] = c_add(res[ki], P_all[img * K + ki]);
    }
    free(P_all);
    return fail ? -1 : 0;
}

/* ---------------------------------------------------------------------------------------
 * pb:264-325  _numba_LC_matrix_batch  (shoebox DEISM-LC).
 * C_*_vec stay complex64 and are promoted per use (pb:619-620 passes them un-cast).
 * ------------------------------------------------------------------------------------- */
static inline cplx lc_factor(cplx atten, double k, double r) {
    /* -1.0 * atten * 4.0 * pi / k * exp(-1j*k*r) / k / r   (pb:314-317), left to right */
    cplx t = c_make(-1.0 * atten.re, -1.0 * atten.im);
    t = c_scale(t, 4.0);
    t = c_scale(t, M_PI);
    t = c_make(t.re / k, t.im / k);
    double ang = -(k * r);
    t = c_mul(t, c_make(cos(ang), sin(ang)));
    t = c_make(t.re / k, t.im / k);
    t = c_make(t.re / r, t.im / r);
    return t;
}

ORACLE_API int oracle_LC_matrix_batch(
    const int64_t *n_all, const int64_t *m_all, const int64_t *v_all, const int64_t *u_all,
    long Js, long Jr, const float *C_nm_s_vec, const float *C_vu_r_vec,
    const double *R_s_rI, const double *R_r_sI, const double *atten, const double *k,
    long n_images, long K, double *result)
{
    const cplxf *Cs = (const cplxf *)C_nm_s_vec;
    const cplxf *Cr = (const cplxf *)C_vu_r_vec;
    cplx *P_all = (cplx *)calloc((size_t)n_images * K, sizeof(cplx));
    cplx *phase_s = (cplx *)malloc((size_t)(Js > 0 ? Js : 1) * sizeof(cplx));
    cplx *phase_r = (cplx *)malloc((size_t)(Jr > 0 ? Jr : 1) * sizeof(cplx));
    if (!P_all || !phase_s || !phase_r) { free(P_all); free(phase_s); free(phase_r); return -1; }
    for (long j = 0; j < Js; ++j) phase_s[j] = c_ipow(n_all[j]);
    for (long j = 0; j < Jr; ++j) phase_r[j] = c_ipow(v_all[j]);
#pragma omp parallel
    {
        cplx *Y_s = (cplx *)malloc((size_t)(Js > 0 ? Js : 1) * sizeof(cplx));
        cplx *Y_r = (cplx *)malloc((size_t)(Jr > 0 ? Jr : 1) * sizeof(cplx));
#pragma omp for schedule(static)
        for (long img = 0; img < n_images; ++img) {
            double phi_s = R_s_rI[img * 3 + 0], theta_s = R_s_rI[img * 3 + 1];
            double r_s = R_s_rI[img * 3 + 2];
            double phi_r = R_r_sI[img * 3 + 0], theta_r = R_r_sI[img * 3 + 1];
            for (long j = 0; j < Js; ++j) Y_s[j] = sph_harm(m_all[j], n_all[j], phi_s, theta_s);
            for (long j = 0; j < Jr; ++j) Y_r[j] = sph_harm(u_all[j], v_all[j], phi_r, theta_r);
            for (long ki = 0; ki < K; ++ki) {
                cplx src = c_make(0.0, 0.0), rec = c_make(0.0, 0.0);
                for (long j = 0; j < Js; ++j) {
                    cplxf c = Cs[ki * Js + j];
                    cplx t = c_mul(c_mul(phase_s[j], c_make((double)c.re, (double)c.im)), Y_s[j]);
                    src = c_add(src, t);
                }
                for (long j = 0; j < Jr; ++j) {
                    cplxf c = Cr[ki * Jr + j];
                    cplx t = c_mul(c_mul(phase_r[j], c_make((double)c.re, (double)c.im)), Y_r[j]);
                    rec = c_add(rec, t);
                }
                cplx f = lc_factor(((const cplx *)atten)[img * K + ki], k[ki], r_s);
                P_all[img * K + ki] = c_mul(c_mul(f, src), rec);
            }
        }
        free(Y_s); free(Y_r);
    }
    cplx *res = (cplx *)result;
    for (long ki = 0; ki < K; ++ki) res[ki] = c_make(0.0, 0.0);
    for (long img = 0; img < n_images; ++img)
        for (long ki = 0; ki < K; ++ki) res[ki] = c_add(res[ki], P_all[img * K + ki]);
    free(P_all); free(phase_s); free(phase_r);
    return 0;
}

/* pb:328-385  _numba_ARG_LC_batch: one direction R_sI_r[3, img] for both sums; phases
 * i^(-n) (-1)^n and i^v (-1)^v; C_s[k, j, img] complex128; atten[k, img]. */
ORACLE_API int oracle_ARG_LC_batch(
    const int64_t *n_all, const int64_t *m_all, const int64_t *v_all, const int64_t *u_all,
    long Js, long Jr, const double *C_nm_s_ARG_vec, const double *C_vu_r_vec,
    const double *R_sI_r, const double *atten, const double *k,
    long n_images, long K, double *result)
{
    const cplx *Cs = (const cplx *)C_nm_s_ARG_vec;
    const cplx *Cr = (const cplx *)C_vu_r_vec;
    cplx *P_all = (cplx *)calloc((size_t)n_images * K, sizeof(cplx));
    cplx *phase_s = (cplx *)malloc((size_t)(Js > 0 ? Js : 1) * sizeof(cplx));
    cplx *phase_r = (cplx *)malloc((size_t)(Jr > 0 ? Jr : 1) * sizeof(cplx));
    if (!P_all || !phase_s || !phase_r) { free(P_all); free(phase_s); free(phase_r); return -1; }
    for (long j = 0; j < Js; ++j)
        phase_s[j] = c_scale(c_ipow(-n_all[j]), minus_one_pow(n_all[j]));
    for (long j = 0; j < Jr; ++j)
        phase_r[j] = c_scale(c_ipow(v_all[j]), minus_one_pow(v_all[j]));
#pragma omp parallel
    {
        cplx *Y_s = (cplx *)malloc((size_t)(Js > 0 ? Js : 1) * sizeof(cplx));
        cplx *Y_r = (cplx *)malloc((size_t)(Jr > 0 ? Jr : 1) * sizeof(cplx));
#pragma omp for schedule(static)
        for (long img = 0; img < n_images; ++img) {
            double phi = R_sI_r[0 * n_images + img], theta = R_sI_r[1 * n_images + img];
            double r = R_sI_r[2 * n_images + img];
            for (long j = 0; j < Js; ++j) Y_s[j] = sph_harm(m_all[j], n_all[j], phi, theta);
            for (long j = 0; j < Jr; ++j) Y_r[j] = sph_harm(u_all[j], v_all[j], phi, theta);
            for (long ki = 0; ki < K; ++ki) {
                cplx src = c_make(0.0, 0.0), rec = c_make(0.0, 0.0);
                for (long j = 0; j < Js; ++j)
                    src = c_add(src, c_mul(c_mul(phase_s[j], Cs[(ki * Js + j) * n_images + img]),
                                           Y_s[j]));
                for (long j = 0; j < Jr; ++j)
                    rec = c_add(rec, c_mul(c_mul(phase_r[j], Cr[ki * Jr + j]), Y_r[j]));
                cplx f = lc_factor(((const cplx *)atten)[ki * n_images + img], k[ki], r);
                P_all[img * K + ki] = c_mul(c_mul(f, src), rec);
            }
        }
        free(Y_s); free(Y_r);
    }
    cplx *res = (cplx *)result;
    for (long ki = 0; ki < K; ++ki) res[ki] = c_make(0.0, 0.0);
    for (long img = 0; img < n_images; ++img)
        for (long ki = 0; ki < K; ++ki) res[ki] = c_add(res[ki], P_all[img * K + ki]);
    free(P_all); free(phase_s); free(phase_r);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * Shoebox image lattice: cd:2649-2728 (count) and cd:2731-2928 (fill).
 * task = parity_idx*(N_o+1) + ref_order, parity_idx = 4 p_x + 2 p_y + p_z (cd:2596-2620).
 * ------------------------------------------------------------------------------------- */
static inline void cart2sph_polar(double x, double y, double z, double *az, double *th,
                                  double *r) {
    /* cd:2623-2631 */
    double h = hypot(x, y);
    *r = hypot(h, z);
    double el = atan2(z, h);
    *az = atan2(y, x);
    *th = M_PI / 2 - el;
}

typedef struct {
    long qx, qy, qz;
    double Isx, Isy, Isz, dist2;
} lattice_pt;

/* Visit the images of one task in the reference order; cb returns nothing. */
#define LATTICE_LOOP_BEGIN(ref_order, px, py, pz)                                          \
    for (long i_abs = 0; i_abs <= (ref_order); ++i_abs)                                    \
    for (long j_abs = 0; j_abs <= (ref_order) - i_abs; ++j_abs) {                          \
        long k_abs = (ref_order) - i_abs - j_abs;                                          \
        int i_cnt = i_abs == 0 ? 1 : 2, j_cnt = j_abs == 0 ? 1 : 2, k_cnt = k_abs == 0 ? 1 : 2; \
        for (int ii = 0; ii < i_cnt; ++ii) {                                               \
            long i = i_abs == 0 ? 0 : (ii == 0 ? -i_abs : i_abs);                          \
            for (int jj = 0; jj < j_cnt; ++jj) {                                           \
                long j = j_abs == 0 ? 0 : (jj == 0 ? -j_abs : j_abs);                      \
                for (int kk = 0; kk < k_cnt; ++kk) {                                       \
                    long kq = k_abs == 0 ? 0 : (kk == 0 ? -k_abs : k_abs);                 \
                    /* Python % is non-negative for a positive modulus (cd:2697-2702) */   \
                    if ((((i + (px)) % 2) + 2) % 2 != 0 || (((j + (py)) % 2) + 2) % 2 != 0 \
                        || (((kq + (pz)) % 2) + 2) % 2 != 0) continue;                     \
                    /* (i+p) is even here, so floor division is exact */                   \
                    long q_x = (i + (px)) / 2, q_y = (j + (py)) / 2, q_z = (kq + (pz)) / 2;
#define LATTICE_LOOP_END }}}}

ORACLE_API void oracle_count_shoebox_images(
    const double *LL, const double *x_s, const double *x_r, double max_dist2,
    long N_o, long N_o_ORG, int64_t *early_counts, int64_t *late_counts)
{
    long n_ref = N_o + 1, n_tasks = 8 * n_ref;
#pragma omp parallel for schedule(dynamic, 1)
    for (long t = 0; t < n_tasks; ++t) {
        long par = t / n_ref, ord = t % n_ref;
        long px = par / 4, py = (par / 2) % 2, pz = par % 2;
        double sox = x_s[0] - 2 * px * x_s[0];
        double soy = x_s[1] - 2 * py * x_s[1];
        double soz = x_s[2] - 2 * pz * x_s[2];
        int64_t ec = 0, lc = 0;
        LATTICE_LOOP_BEGIN(ord, px, py, pz)
            double Isx = 2 * q_x * LL[0] + sox;
            double Isy = 2 * q_y * LL[1] + soy;
            double Isz = 2 * q_z * LL[2] + soz;
            double dx = Isx - x_r[0], dy = Isy - x_r[1], dz = Isz - x_r[2];
            double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 > max_dist2) continue;
            if (ord <= N_o_ORG) ++ec; else ++lc;
        LATTICE_LOOP_END
        early_counts[t] = ec;
        late_counts[t] = lc;
    }
}

ORACLE_API void oracle_fill_shoebox_images(
    const double *LL, const double *x_s, const double *x_r, const double *Z_S, long K,
    double max_dist2, int angdep_flag, long N_o, long N_o_ORG, int store_attenuation,
    const int64_t *early_offsets, const int64_t *late_offsets,
    float *R_sI_r_e, float *R_s_rI_e, float *R_r_sI_e, float *atten_e, int32_t *A_e,
    float *R_sI_r_l, float *R_s_rI_l, float *R_r_sI_l, float *atten_l, int32_t *A_l)
{
    const cplx *Z = (const cplx *)Z_S;
    long n_ref = N_o + 1, n_tasks = 8 * n_ref;
#pragma omp parallel for schedule(dynamic, 1)
    for (long t = 0; t < n_tasks; ++t) {
        long par = t / n_ref, ord = t % n_ref;
        long px = par / 4, py = (par / 2) % 2, pz = par % 2;
        double sox = x_s[0] - 2 * px * x_s[0];
        double soy = x_s[1] - 2 * py * x_s[1];
        double soz = x_s[2] - 2 * pz * x_s[2];
        long e_idx = early_offsets[t], l_idx = late_offsets[t];
        LATTICE_LOOP_BEGIN(ord, px, py, pz)
            double Isx = 2 * q_x * LL[0] + sox;
            double Isy = 2 * q_y * LL[1] + soy;
            double Isz = 2 * q_z * LL[2] + soz;
            double dx = Isx - x_r[0], dy = Isy - x_r[1], dz = Isz - x_r[2];
            double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 > max_dist2) continue;

            double Irx = (double)(1 - 2 * px) * (x_r[0] - 2 * q_x * LL[0]);
            double Iry = (double)(1 - 2 * py) * (x_r[1] - 2 * q_y * LL[1]);
            double Irz = (double)(1 - 2 * pz) * (x_r[2] - 2 * q_z * LL[2]);

            double ax = x_r[0] - Isx, ay = x_r[1] - Isy, az_ = x_r[2] - Isz;   /* R_sI_r */
            double bx = Irx - x_s[0], by = Iry - x_s[1], bz = Irz - x_s[2];     /* R_s_rI */
            double cx_ = Isx - x_r[0], cy_ = Isy - x_r[1], cz_ = Isz - x_r[2];  /* R_r_sI */
            double a0, a1, a2, b0, b1, b2, c0, c1, c2;
            cart2sph_polar(ax, ay, az_, &a0, &a1, &a2);
            cart2sph_polar(bx, by, bz, &b0, &b1, &b2);
            cart2sph_polar(cx_, cy_, cz_, &c0, &c1, &c2);

            double cosx = 1.0, cosy = 1.0, cosz = 1.0;
            if (angdep_flag == 1) {
                double rn = sqrt(d2);
                cosx = fabs(ax) / rn; cosy = fabs(ay) / rn; cosz = fabs(az_) / rn;
            }
            long e[6];
            shoebox_exponents(q_x, q_y, q_z, px, py, pz, e);

            long idx; int32_t *A; float *Ra, *Rb, *Rc; cplxf *at;
            if (ord <= N_o_ORG) {
                idx = e_idx++; A = A_e; Ra = R_sI_r_e; Rb = R_s_rI_e; Rc = R_r_sI_e;
                at = (cplxf *)atten_e;
            } else {
                idx = l_idx++; A = A_l; Ra = R_sI_r_l; Rb = R_s_rI_l; Rc = R_r_sI_l;
                at = (cplxf *)atten_l;
            }
            A[idx * 6 + 0] = (int32_t)q_x; A[idx * 6 + 1] = (int32_t)q_y;
            A[idx * 6 + 2] = (int32_t)q_z; A[idx * 6 + 3] = (int32_t)px;
            A[idx * 6 + 4] = (int32_t)py;  A[idx * 6 + 5] = (int32_t)pz;
            Ra[idx * 3 + 0] = (float)a0; Ra[idx * 3 + 1] = (float)a1; Ra[idx * 3 + 2] = (float)a2;
            Rb[idx * 3 + 0] = (float)b0; Rb[idx * 3 + 1] = (float)b1; Rb[idx * 3 + 2] = (float)b2;
            Rc[idx * 3 + 0] = (float)c0; Rc[idx * 3 + 1] = (float)c1; Rc[idx * 3 + 2] = (float)c2;
            if (store_attenuation == 1) {
                for (long fi = 0; fi < K; ++fi) {
                    cplx a = shoebox_atten(Z, K, fi, cosx, cosy, cosz, e);
                    at[idx * K + fi].re = (float)a.re;
                    at[idx * K + fi].im = (float)a.im;
                }
            }
        LATTICE_LOOP_END
    }
}
