// common.cuh — shared device/host helpers of libdeism_b200 (sm_100a only).
//
// Device math here restates the scalar helpers of the reference's Numba backend
// (deism/parallel_backends.py, "pb"): _sph_harm_numba pb:40-83, _sphankel2_numba pb:86-118,
// _shoebox_ref_coef_from_cos_numba pb:121, _shoebox_complex_pow_int_numba pb:125.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deism_b200.h"

namespace deism {

// ---------------------------------------------------------------------------------------
// host-side error / bookkeeping (api.cu)
// ---------------------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);
// Image-chunk schedule with chunks of decreasing size (api.cu): filled in and true when a simulation of
// the block scheduler's greedy hand-out says it beats `uni_chunks` equal chunks of `uni_tpc` tiles
// (from the SECOND call with a given shape on: see api.cu).
constexpr int DEISM_MAXCHUNKS = 96;
struct GuidedPlan {
    int n;
    int lo[DEISM_MAXCHUNKS + 1];
};
bool guided_chunks(long n_img_tiles, long n_ftiles, long slots, long uni_chunks, long uni_tpc,
                   long max_tiles_per_chunk, GuidedPlan &out);
// whether a later call with this shape may come back with a plan (of up to DEISM_MAXCHUNKS chunks):
// scratch that depends on the chunk count is sized for that from the first call on
bool guided_chunks_possible(long n_img_tiles, long n_ftiles, long uni_chunks);
void bump_generation();                   // scratch pointers or __constant__ tables changed
int check_device();                       // DEISM_OK iff current device is sm_100
void *workspace(int slot, size_t bytes);  // cached device scratch; nullptr on failure
int sm_count();
int smem_optin_bytes();   // largest dynamic shared memory a CTA may opt into
// optional CUDA-event bracket around a named kernel (no-op unless deism_profile_enable(1))
void prof_begin(const char *name, cudaStream_t st);
void prof_end(const char *name, cudaStream_t st);

// Serialises the entry points per device and orders calls issued on different streams (api.cu).
class CallGuard {
public:
    explicit CallGuard(void *stream);
    ~CallGuard();
    CallGuard(const CallGuard &) = delete;
    CallGuard &operator=(const CallGuard &) = delete;
private:
    void *st_;
    int dev_ = -1;
    bool capturing_ = false;
};

enum WsSlot {
    WS_LC_IMG = 0, WS_LC_YS, WS_LC_YR, WS_PARTIAL, WS_FLAGS, WS_SMALL, WS_ORG_B, WS_ORG_IMG,
    WS_ORG_Y, WS_ORG_LIST, WS_ARG_IDX, WS_CONVEX_A, WS_CONVEX_B, WS_CONVEX_C, WS_CONVEX_D,
    WS_CONVEX_E, WS_BENCH,
    // the ORG entry points own their scratch: the ORG and LC parts of a MIX run may execute
    // concurrently on two streams
    WS_ORG_SMALL, WS_ORG_PARTIAL, WS_ORG_FLAGS, WS_NSLOTS
};

#define DEISM_CUDA_TRY(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return deism::set_error(DEISM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

#define DEISM_LAUNCH_CHECK(name)                                                              \
    do {                                                                                      \
        deism::count_launch();                                                                \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return deism::set_error(DEISM_ERR_CUDA, "launch of %s failed: %s", name,          \
                                    cudaGetErrorString(_e));                                  \
    } while (0)

// ---------------------------------------------------------------------------------------
// complex128 as double2
// ---------------------------------------------------------------------------------------
typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmake(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
// acc += a*b  (4 DFMA)
__device__ __forceinline__ void cfma(cplx &acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    double inv = 1.0 / fma(b.x, b.x, b.y * b.y);
    return cmake(fma(a.x, b.x, a.y * b.y) * inv, fma(a.y, b.x, -a.x * b.y) * inv);
}
// multiply by i^n
__device__ __forceinline__ cplx cmul_ipow(cplx a, int n) {
    switch (n & 3) {
    case 0: return a;
    case 1: return cmake(-a.y, a.x);
    case 2: return cmake(-a.x, -a.y);
    default: return cmake(a.y, -a.x);
    }
}
// round to complex64 and back (the reference's Q2 quantisation point, cd:2921-2928)
__device__ __forceinline__ cplx round_c64(cplx a) {
    return cmake((double)__double2float_rn(a.x), (double)__double2float_rn(a.y));
}

// ---------------------------------------------------------------------------------------
// sincos with a 3-term Cody-Waite reduction (exactly-rounded differences via FMA) and the
// fdlibm minimax kernels; |error| < 2 ulp for |x| < 2^30.  CUDA's own sincos() drops into a
// slow Payne-Hanek path above |x| = 105615, which k*r reaches (k up to 440 rad/m, r to 350 m).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_cw(double x, double *sp, double *cp) {
    double q = rint(x * 0.6366197723675814);
    int n = (int)q;
    double r = fma(-q, 1.5707963267948966, x);
    r = fma(-q, 6.123233995736766e-17, r);
    r = fma(-q, -1.4973849048591698e-33, r);
    double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    double s = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    if (n & 1) { double t = s; s = c; c = -t; }
    if (n & 2) { s = -s; c = -c; }
    *sp = s;
    *cp = c;
}

// ---------------------------------------------------------------------------------------
// pb:40-83  Y_n^m(theta = polar, phi = azimuth), scipy convention.
// ct = cos(theta); (sphi_m, cphi_m) supplied by the caller would save work, but n <= ~10 and
// this runs once per image, never per term.
// ---------------------------------------------------------------------------------------
__device__ inline cplx sph_harm_dev(int m, int n, double phi, double ct) {
    int am = m < 0 ? -m : m;
    double pmm = 1.0;
    if (am > 0) {
        double somx2 = sqrt(fma(-ct, ct, 1.0));
        double fact = 1.0;
        for (int i = 1; i <= am; ++i) {
            pmm *= -fact * somx2;
            fact += 2.0;
        }
    }
    double p_val;
    if (n == am) {
        p_val = pmm;
    } else {
        double pmm1 = ct * (double)(2 * am + 1) * pmm;
        for (int ll = am + 2; ll <= n; ++ll) {
            double pll = (ct * (double)(2 * ll - 1) * pmm1 - (double)(ll + am - 1) * pmm) /
                         (double)(ll - am);
            pmm = pmm1;
            pmm1 = pll;
        }
        p_val = pmm1;
    }
    double ratio = 1.0;
    for (int i = n - am + 1; i <= n + am; ++i) ratio *= (double)i;
    double norm = sqrt((double)(2 * n + 1) / (4.0 * 3.14159265358979323846) / ratio);
    double s, c;
    sincos_cw((double)am * phi, &s, &c);
    double np_ = norm * p_val;
    cplx pos = cmake(np_ * c, np_ * s);
    if (m < 0) {
        double sign = (am & 1) ? -1.0 : 1.0;
        return cmake(sign * pos.x, -sign * pos.y);
    }
    return pos;
}

// ---------------------------------------------------------------------------------------
// Shoebox wall reflection factor (pb:121 / cd:2635) and its integer power (pb:125 / cd:2641;
// the reference multiplies `e` times, here by squaring: identical to ~e*1e-16 relative, far
// below the 6e-8 complex64 rounding step that follows — see DESIGN.md "quantisation").
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ cplx ref_coef(double cos_t, cplx zeta) {
    double zr = zeta.x * cos_t, zi = zeta.y * cos_t;
    return cdiv(cmake(zr - 1.0, zi), cmake(zr + 1.0, zi));
}
__device__ __forceinline__ cplx cpow_uint(cplx base, unsigned e) {
    cplx out = cmake(1.0, 0.0);
    while (e) {
        if (e & 1u) out = cmul(out, base);
        e >>= 1;
        if (e) base = cmul(base, base);
    }
    return out;
}
// atten = bx1^e0 * bx2^e1 * by1^e2 * by2^e3 * bz1^e4 * bz2^e5   (cd:2921-2928)
// Z points at Z_S[0][k]; zstride = K (row stride in complex elements).
__device__ __forceinline__ bool csame(cplx a, cplx b) { return a.x == b.x && a.y == b.y; }
// one wall pair: b1^e1 * b2^e2; when both walls carry the same impedance (the common case) the
// two factors share one division and one power, b^(e1+e2)
__device__ __forceinline__ cplx wall_pair(cplx z1, cplx z2, double c, unsigned e1, unsigned e2) {
    if (csame(z1, z2)) return cpow_uint(ref_coef(c, z1), e1 + e2);
    return cmul(cpow_uint(ref_coef(c, z1), e1), cpow_uint(ref_coef(c, z2), e2));
}
__device__ __forceinline__ cplx shoebox_atten(const cplx *__restrict__ Z, long zstride, double cx,
                                              double cy, double cz, const unsigned char *e) {
    cplx a = wall_pair(Z[0], Z[zstride], cx, e[0], e[1]);
    a = cmul(a, wall_pair(Z[2 * zstride], Z[3 * zstride], cy, e[2], e[3]));
    a = cmul(a, wall_pair(Z[4 * zstride], Z[5 * zstride], cz, e[4], e[5]));
    return a;
}

// The same product for NQ frequencies of ONE image at once: the exponents (and with them the
// control flow of the powers) are shared, so the NQ chains advance in one loop with NQ-way
// instruction-level parallelism.  Element for element the operations are those of shoebox_atten.
template <int NQ>
__device__ __forceinline__ void cpow_uint_n(cplx (&base)[NQ], unsigned e, cplx (&out)[NQ]) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = cmake(1.0, 0.0);
    while (e) {
        if (e & 1u) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) out[q] = cmul(out[q], base[q]);
        }
        e >>= 1;
        if (e) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) base[q] = cmul(base[q], base[q]);
        }
    }
}
// Real base, integer power, NQ chains at once (one DMUL per step instead of a complex product).
template <int NQ>
__device__ __forceinline__ void rpow_uint_n(double (&base)[NQ], unsigned e, double (&out)[NQ]) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = 1.0;
    while (e) {
        if (e & 1u) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) out[q] *= base[q];
        }
        e >>= 1;
        if (e) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) base[q] *= base[q];
        }
    }
}
// 1/x for the per-term chains: hardware seed + two Newton steps (<= 1 ulp).
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    return fma(fma(-x, r, 1.0), r, r);
}
// ref_coef with that reciprocal (the per-image / materialising callers keep the IEEE division)
__device__ __forceinline__ cplx ref_coef_fast(double cos_t, cplx zeta) {
    const double zr = zeta.x * cos_t, zi = zeta.y * cos_t;
    const double ax = zr - 1.0, bx = zr + 1.0;
    const double inv = rcp_newton(fma(bx, bx, zi * zi));
    return cmake(fma(ax, bx, zi * zi) * inv, fma(zi, bx, -ax * zi) * inv);
}
// An impedance whose imaginary part is below 2 ulp of its real part is treated as real: the
// reference's material conversions return Re + 1e-16j (cd:807-880), for which the complex factor
// (zeta c - 1)/(zeta c + 1) differs from the real one by ~1e-17 absolutely — two orders below the
// rounding of zeta c - 1 itself.  The real chain is one real division and real powers per wall.
__device__ __forceinline__ bool z_is_real(cplx z) { return fabs(z.y) <= 4.0e-16 * fabs(z.x); }
// (t - 1)/(t + 1) = 1 - 2/(t + 1) with the hardware reciprocal seed and two Newton steps (<= 1 ulp
// on 1/(t+1)): a third of the instructions of the IEEE division sequence.
__device__ __forceinline__ double real_ref_coef(double t) {
    return fma(-2.0, rcp_newton(t + 1.0), 1.0);
}
template <int NQ>
__device__ __forceinline__ void shoebox_atten_n(const cplx *__restrict__ Z, long zstride, const int (&fq)[NQ],
                                                double cx, double cy, double cz, const unsigned char *e,
                                                cplx (&out)[NQ]) {
    const double c3[3] = {cx, cy, cz};
#pragma unroll
    for (int w = 0; w < 3; ++w) {
        cplx z1[NQ], z2[NQ];
        bool same = true, real = true;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            z1[q] = Z[(2 * w) * zstride + fq[q]];
            z2[q] = Z[(2 * w + 1) * zstride + fq[q]];
            same = same && csame(z1[q], z2[q]);
            real = real && z_is_real(z1[q]) && z_is_real(z2[q]);
        }
        const unsigned e1 = e[2 * w], e2 = e[2 * w + 1];
        if (real) {
            double r1[NQ], rp[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                r1[q] = real_ref_coef(z1[q].x * c3[w]);
            }
            if (same) {
                rpow_uint_n<NQ>(r1, e1 + e2, rp);
            } else {
                double r2[NQ], rp2[NQ];
                rpow_uint_n<NQ>(r1, e1, rp);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    r2[q] = real_ref_coef(z2[q].x * c3[w]);
                }
                rpow_uint_n<NQ>(r2, e2, rp2);
#pragma unroll
                for (int q = 0; q < NQ; ++q) rp[q] *= rp2[q];
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                out[q] = w == 0 ? cmake(rp[q], 0.0) : cmake(out[q].x * rp[q], out[q].y * rp[q]);
            continue;
        }
        cplx b1[NQ], p[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) b1[q] = ref_coef_fast(c3[w], z1[q]);
        if (same) {
            cpow_uint_n<NQ>(b1, e1 + e2, p);
        } else {
            cplx b2[NQ], p2[NQ];
            cpow_uint_n<NQ>(b1, e1, p);
#pragma unroll
            for (int q = 0; q < NQ; ++q) b2[q] = ref_coef_fast(c3[w], z2[q]);
            cpow_uint_n<NQ>(b2, e2, p2);
#pragma unroll
            for (int q = 0; q < NQ; ++q) p[q] = cmul(p[q], p2[q]);
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) out[q] = w == 0 ? p[q] : cmul(out[q], p[q]);
    }
}

// ---------------------------------------------------------------------------------------
// Real impedances on every wall and frequency of a tile (decided once per CTA): the attenuation is
// REAL, and the whole product  prod_w beta_w^{e_w}  is formed by ONE simultaneous square-and-multiply
// over the (up to six) exponents — one squaring per bit position instead of one per wall — from
// wall factors 1 - 2/(zeta c + 1).  29 FP64 instructions per term instead of the 66 of the general
// chain above (which carries a complex accumulator through the three wall pairs).
//   NEWTON = 1: one Newton step on the hardware reciprocal seed (2^-23 -> 2^-46).  Enough wherever
//               the attenuation is rounded to complex64 right after (DEISM_ATTEN_FUSED_ROOM, the
//               reference's default storage, cd:2921-2928): a value moves to the neighbouring float
//               only if it lies within 2^-46 of a rounding boundary (~3e-7 of the values, one
//               float ulp of ONE term each).
//   NEWTON = 2: <= 1 ulp of FP64 (compact storage, not rounded).
// ---------------------------------------------------------------------------------------
template <int NEWTON>
__device__ __forceinline__ double real_ref_coef_n(double zeta, double c) {
    const double x = fma(zeta, c, 1.0);          // zeta c + 1 in one rounding
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    if (NEWTON > 1) r = fma(fma(-x, r, 1.0), r, r);
    return fma(-2.0, r, 1.0);
}
// PAIRS = true: the caller has established (per CTA, over the whole frequency tile) that the two
// walls of every axis carry the same impedance, so an axis is ONE base with the summed exponent:
// three reciprocals and 1 + 3 multiplications per bit position.  The selects the compiler turns the
// bit tests into execute for every base, so bases that cannot occur must not be in the loop at all.
template <int NQ, int NEWTON, bool PAIRS>
__device__ __forceinline__ void shoebox_atten_real_n(const cplx *__restrict__ Z, long zstride, const int (&fq)[NQ],
                                                     double cx, double cy, double cz, const unsigned char *e,
                                                     double (&out)[NQ]) {
    const double c3[3] = {cx, cy, cz};
    constexpr int NB = PAIRS ? 3 : 6;
    double b[NB][NQ];
    unsigned ex[NB];
    if constexpr (PAIRS) {
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            ex[w] = (unsigned)e[2 * w] + (unsigned)e[2 * w + 1];
#pragma unroll
            for (int q = 0; q < NQ; ++q) b[w][q] = real_ref_coef_n<NEWTON>(Z[(2 * w) * zstride + fq[q]].x, c3[w]);
        }
    } else {
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            bool same = true;
            double z1[NQ], z2[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                z1[q] = Z[(2 * w) * zstride + fq[q]].x;
                z2[q] = Z[(2 * w + 1) * zstride + fq[q]].x;
                same = same && z1[q] == z2[q];
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) b[2 * w][q] = real_ref_coef_n<NEWTON>(z1[q], c3[w]);
            if (same) {
                ex[2 * w] = (unsigned)e[2 * w] + (unsigned)e[2 * w + 1];
                ex[2 * w + 1] = 0u;
#pragma unroll
                for (int q = 0; q < NQ; ++q) b[2 * w + 1][q] = 1.0;
            } else {
                ex[2 * w] = e[2 * w];
                ex[2 * w + 1] = e[2 * w + 1];
#pragma unroll
                for (int q = 0; q < NQ; ++q) b[2 * w + 1][q] = real_ref_coef_n<NEWTON>(z2[q], c3[w]);
            }
        }
    }
    unsigned all = 0u;
#pragma unroll
    for (int s_ = 0; s_ < NB; ++s_) all |= ex[s_];
    if (all == 0u) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) out[q] = 1.0;
        return;
    }
    // Left to right over the bit positions of the largest exponent.  The multiplications by a base
    // must stay PREDICATED on the exponent bit (the compiler's @P DMUL): a lane whose bit is clear
    // then costs the FP64 pipe nothing, whereas multiplying by a selected 1.0 always pays (measured:
    // 18.8 ms predicated against 19.4 ms with selects on C5).  The top position has nothing to square.
    int bit = 31 - __clz(all);
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = 1.0;
#pragma unroll
    for (int s_ = 0; s_ < NB; ++s_) {
        if ((ex[s_] >> bit) & 1u) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) out[q] *= b[s_][q];
        }
    }
    for (--bit; bit >= 0; --bit) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) out[q] *= out[q];
#pragma unroll
        for (int s_ = 0; s_ < NB; ++s_) {
            if ((ex[s_] >> bit) & 1u) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) out[q] *= b[s_][q];
            }
        }
    }
}

// one (image, frequency) term inside a main loop: the chains above with NQ = 1
__device__ __forceinline__ cplx shoebox_atten_term(const cplx *__restrict__ Z, long zstride, double cx, double cy,
                                                   double cz, const unsigned char *e) {
    const int fq[1] = {0};
    cplx out[1];
    shoebox_atten_n<1>(Z, zstride, fq, cx, cy, cz, e, out);
    return out[0];
}

// Per-image attenuation geometry shared by images.cu / lc.cu / org.cu
struct AttenGeom {
    double cx, cy, cz;
    unsigned char e[8];  // e[0..5] exponents; e[6], e[7] padding
};

// Direction cosines exactly as the reference fill computes them (cd:2804-2851): FP64, no FMA
// contraction (Numba/LLVM does not contract), from A and the room geometry.
__device__ __forceinline__ void atten_geom_room(const int32_t *A6, const double *LL, const double *xs,
                                                const double *xr, int angdep, AttenGeom &g) {
    int qx = A6[0], qy = A6[1], qz = A6[2], px = A6[3], py = A6[4], pz = A6[5];
    g.e[0] = (unsigned char)abs(qx - px); g.e[1] = (unsigned char)abs(qx);
    g.e[2] = (unsigned char)abs(qy - py); g.e[3] = (unsigned char)abs(qy);
    g.e[4] = (unsigned char)abs(qz - pz); g.e[5] = (unsigned char)abs(qz);
    g.e[6] = g.e[7] = 0;
    if (angdep != 1) { g.cx = g.cy = g.cz = 1.0; return; }
    double sox = __dsub_rn(xs[0], __dmul_rn((double)(2 * px), xs[0]));
    double soy = __dsub_rn(xs[1], __dmul_rn((double)(2 * py), xs[1]));
    double soz = __dsub_rn(xs[2], __dmul_rn((double)(2 * pz), xs[2]));
    double Isx = __dadd_rn(__dmul_rn((double)(2 * qx), LL[0]), sox);
    double Isy = __dadd_rn(__dmul_rn((double)(2 * qy), LL[1]), soy);
    double Isz = __dadd_rn(__dmul_rn((double)(2 * qz), LL[2]), soz);
    double dx = __dsub_rn(Isx, xr[0]), dy = __dsub_rn(Isy, xr[1]), dz = __dsub_rn(Isz, xr[2]);
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    double rn = __dsqrt_rn(d2);
    g.cx = __ddiv_rn(fabs(__dsub_rn(xr[0], Isx)), rn);
    g.cy = __ddiv_rn(fabs(__dsub_rn(xr[1], Isy)), rn);
    g.cz = __ddiv_rn(fabs(__dsub_rn(xr[2], Isz)), rn);
}
// Compact-storage variant (pb:131-137): cosines from the f32-rounded angles.
__device__ __forceinline__ void atten_geom_angles(const int32_t *A6, const float *R3, int angdep,
                                                  AttenGeom &g) {
    int qx = A6[0], qy = A6[1], qz = A6[2], px = A6[3], py = A6[4], pz = A6[5];
    g.e[0] = (unsigned char)abs(qx - px); g.e[1] = (unsigned char)abs(qx);
    g.e[2] = (unsigned char)abs(qy - py); g.e[3] = (unsigned char)abs(qy);
    g.e[4] = (unsigned char)abs(qz - pz); g.e[5] = (unsigned char)abs(qz);
    g.e[6] = g.e[7] = 0;
    if (angdep != 1) { g.cx = g.cy = g.cz = 1.0; return; }
    double sp, cp, st, ct;
    sincos_cw((double)R3[0], &sp, &cp);
    sincos_cw((double)R3[1], &st, &ct);
    g.cx = fabs(st * cp);
    g.cy = fabs(st * sp);
    g.cz = fabs(ct);
}

// ---- mbarrier / bulk-copy (TMA) / DMMA primitives ---------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
#if defined(DEISM_MBAR_TEST)
    // spin on the non-suspending test
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#elif defined(DEISM_MBAR_HINT)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"((unsigned)DEISM_MBAR_HINT)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy (TMA, SASS UBLKCP); bytes must be a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only) and the mbarrier arrival that fires
// when all of the executing thread's earlier cp.async have landed (.noinc: it counts against the
// barrier's expected arrivals)
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D(8x8) += A(8x4) * B(4x8), FP64 tensor core
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Device copy of the room part of deism_shoebox_atten (passed by value to kernels)
struct RoomGeom {
    double LL[3], xs[3], xr[3];
};

}  // namespace deism
