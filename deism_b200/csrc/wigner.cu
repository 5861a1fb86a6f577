// wigner.cu — exact Wigner 3j tables (host code; setup, not the hot path).
//
// Replaces pre_calc_Wigner (deism/core_deism.py:1842-1913), which calls
// sympy.physics.wigner.wigner_3j once per (n,m,v,u,l) — 169 s at N=V=10.
//
//   W_1[n,v,l]      = (n v l; 0 0 0)
//   W_2[n,v,l,m,u]  = (n v l; -m u m-u)      for |u-m| <= l, |n-v| <= l <= n+v, else 0
//
// Method: Racah's single-sum formula rewritten with binomials so that the alternating sum is
// an exact integer (__int128), and the square-root prefactor is a ratio of factorials kept as
// a prime-exponent vector; the value is assembled in long double (64-bit mantissa) and rounded
// double -> float, the reference's rounding chain (sympy exact -> float64 -> complex64,
// cd:1896-1904).  Valid for n+v+l+1 <= 170 and binomial arguments <= 62.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace deism {

static const int kPrimes[] = {2,  3,  5,  7,  11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61,
                              67, 71, 73, 79, 83, 89, 97, 101, 103, 107, 109, 113, 127, 131, 137,
                              139, 149, 151, 157, 163, 167};
static const int kNPrimes = sizeof(kPrimes) / sizeof(int);

// exps += sign * prime exponents of n!
static void add_factorial(int *exps, int n, int sign) {
    for (int i = 0; i < kNPrimes; ++i) {
        int p = kPrimes[i], e = 0;
        for (long q = p; q <= n; q *= p) e += n / (int)q;
        exps[i] += sign * e;
    }
}

static __int128 binom(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;   // exact at every step
    return r;
}

// (j1 j2 j3; m1 m2 m3), integer arguments
static double wigner3j(int j1, int j2, int j3, int m1, int m2, int m3) {
    if (m1 + m2 + m3 != 0) return 0.0;
    if (abs(m1) > j1 || abs(m2) > j2 || abs(m3) > j3) return 0.0;
    if (j3 < abs(j1 - j2) || j3 > j1 + j2) return 0.0;
    const int a = j1 + j2 - j3, b = j1 - j2 + j3, c = -j1 + j2 + j3, J = j1 + j2 + j3;
    // integer alternating sum  B = sum_t (-1)^t C(a,t) C(b, j1-m1-t) C(c, j2+m2-t)
    int tmin = 0;
    if (j2 - j3 - m1 > tmin) tmin = j2 - j3 - m1;   // j3-j2+t+m1 >= 0
    if (j1 - j3 + m2 > tmin) tmin = j1 - j3 + m2;   // j3-j1+t-m2 >= 0
    int tmax = a;
    if (j1 - m1 < tmax) tmax = j1 - m1;
    if (j2 + m2 < tmax) tmax = j2 + m2;
    __int128 B = 0;
    for (int t = tmin; t <= tmax; ++t) {
        __int128 term = binom(a, t) * binom(b, j1 - m1 - t) * binom(c, j2 + m2 - t);
        B += (t & 1) ? -term : term;
    }
    if (B == 0) return 0.0;
    // value = (-1)^(j1-j2-m3) * B * sqrt( prod (j_i +- m_i)! / (a! b! c! (J+1)!) )
    int exps[kNPrimes];
    memset(exps, 0, sizeof(exps));
    add_factorial(exps, j1 + m1, 1); add_factorial(exps, j1 - m1, 1);
    add_factorial(exps, j2 + m2, 1); add_factorial(exps, j2 - m2, 1);
    add_factorial(exps, j3 + m3, 1); add_factorial(exps, j3 - m3, 1);
    add_factorial(exps, a, -1); add_factorial(exps, b, -1); add_factorial(exps, c, -1);
    add_factorial(exps, J + 1, -1);
    long double sq_num = 1.0L, sq_den = 1.0L, free_num = 1.0L, free_den = 1.0L;
    for (int i = 0; i < kNPrimes; ++i) {
        int e = exps[i];
        if (e == 0) continue;
        long double p = (long double)kPrimes[i];
        int ae = e < 0 ? -e : e;
        long double half = powl(p, (long double)(ae / 2));
        if (e > 0) { sq_num *= half; if (ae & 1) free_num *= p; }
        else       { sq_den *= half; if (ae & 1) free_den *= p; }
    }
    long double mag = (long double)B * (sq_num / sq_den) * sqrtl(free_num / free_den);
    int ph = j1 - j2 - m3;
    if (ph & 1) mag = -mag;
    return (double)mag;
}

}  // namespace deism

using namespace deism;

extern "C" int deism_wigner_tables(int N, int V, float *h_W1, float *h_W2) {
    if (N < 0 || V < 0 || N > 20 || V > 20)
        return set_error(DEISM_ERR_INVALID, "Wigner tables support orders 0..20 (got N=%d V=%d)", N, V);
    if (!h_W1 || !h_W2) return set_error(DEISM_ERR_INVALID, "null table pointer");
    const long L = N + V + 1, M2 = 2 * N + 1, U2 = 2 * V + 1;
    memset(h_W1, 0, sizeof(float) * 2 * (size_t)((N + 1) * (V + 1) * L));
    memset(h_W2, 0, sizeof(float) * 2 * (size_t)((N + 1) * (V + 1) * L * M2 * U2));
    for (int n = 0; n <= N; ++n)
        for (int m = -n; m <= n; ++m)
            for (int v = 0; v <= V; ++v)
                for (int u = -v; u <= v; ++u)
                    for (int l = abs(n - v); l <= n + v; ++l) {
                        if (abs(u - m) > l) continue;   // cd:1895
                        long i1 = ((long)n * (V + 1) + v) * L + l;
                        h_W1[2 * i1] = (float)wigner3j(n, v, l, 0, 0, 0);
                        long mi = m < 0 ? m + M2 : m, ui = u < 0 ? u + U2 : u;
                        long i2 = ((i1 * M2) + mi) * U2 + ui;
                        h_W2[2 * i2] = (float)wigner3j(n, v, l, -m, u, m - u);
                    }
    return DEISM_OK;
}
