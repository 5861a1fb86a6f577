/*
 * oracle/convex_oracle.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the convex-room (DEISM-ARG) image enumeration of the reference's
 * native extension libroom_deism:
 *     Room_deism<3>::image_source_model / image_sources_dfs / is_visible_dfs /
 *     get_image_attenuation / fill_sources              deism/libroom_src/room.cpp:271-511
 *     Wall_deism<3>::reflect / side / intersection / get_attenuation
 *                                                        deism/libroom_src/wall.cpp:443-552,633-635
 *     intersection_3d_segment_plane / ccw3p / is_left / on_segment / is_inside_2d_polygon
 *                                                        deism/libroom_src/geometry.cpp:54-80,175-323
 *
 * Everything is float32, as in the reference (Eigen float matrices).  Evaluation order of the
 * 3-element inner products follows Eigen 3.3's unrolled redux, x0 + (x1 + x2); no FMA (the
 * reference is built without -march, setup.py:72).  acosf / cosf are transcriptions of the
 * glibc 2.39 x86-64 routines the reference binary calls on the build machine (fdlibm-derived
 * __ieee754_acosf; the FMA variant of the ARM-optimized-routines cosf), so that attenuations
 * can be compared bit for bit.
 *
 * Pinned against the compiled reference (oracle/_ref/libroom_deism*.so) by
 * tests/test_oracle_convex.py: primitives on random inputs and whole image lists
 * (sources, orders, gen_walls, attenuations, reflection matrices) bit-equal.
 *
 * The obstruction test (room.cpp:549-598) is only active for walls outside the convex hull;
 * DEISM-ARG supports convex rooms only (obstructing_walls is empty), so it is not restated.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))
#define EPS 1e-5f   /* libroom_eps, wall.hpp:42 */

/* ------------------------------------------------------------------------------------- */
/* glibc 2.39 __ieee754_acosf (sysdeps/ieee754/flt-32/e_acosf.c), SSE scalar, no FMA        */
/* ------------------------------------------------------------------------------------- */
static const float
    a_one = 1.0f, a_pi = 3.1415925026e+00f, a_pio2_hi = 1.5707962513e+00f,
    a_pio2_lo = 7.5497894159e-08f, pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f,
    pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f,
    pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
    qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

ORACLE_API float oracle_acosf(float x) {
    int32_t hx = (int32_t)f2u(x);
    int32_t ix = hx & 0x7fffffff;
    float z, p, q, r, w, s, c, df;
    if (ix == 0x3f800000) {
        if (hx > 0) return 0.0f;
        return a_pi + 2.0f * a_pio2_lo;
    } else if (ix > 0x3f800000) {
        return (x - x) / (x - x);
    }
    if (ix < 0x3f000000) {
        if (ix <= 0x32800000) return a_pio2_hi + a_pio2_lo;
        z = x * x;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = a_one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        return a_pio2_hi - (x - (a_pio2_lo - x * r));
    } else if (hx < 0) {
        z = (a_one + x) * 0.5f;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = a_one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        s = sqrtf(z);
        r = p / q;
        w = r * s - a_pio2_lo;
        return a_pi - 2.0f * (s + w);
    } else {
        z = (a_one - x) * 0.5f;
        s = sqrtf(z);
        df = u2f(f2u(s) & 0xfffff000u);
        c = (z - df * df) / (s + df);
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = a_one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        w = r * s + c;
        return 2.0f * (df + w);
    }
}

/* ------------------------------------------------------------------------------------- */
/* glibc 2.39 __cosf_fma (sysdeps/ieee754/flt-32/s_cosf.c + sincosf.h, built with -mfma),  */
/* |x| < 120 paths only (theta = acosf(.) lies in [0, pi])                                 */
/* ------------------------------------------------------------------------------------- */
static const double
    t_hpi_inv = 0x1.45f306dc9c883p+23, t_hpi = 0x1.921fb54442d18p+0,
    t_c0 = 1.0, t_c1 = -0x1.ffffffd0c621cp-2, t_c2 = 0x1.55553e1068f19p-5,
    t_c3 = -0x1.6c087e89a359dp-10, t_c4 = 0x1.99343027bf8c3p-16,
    t_s1 = -0x1.555545995a603p-3, t_s2 = 0x1.1107605230bc4p-7, t_s3 = -0x1.994eb3774cf24p-13;
static const double t_sign[4] = { 1.0, -1.0, -1.0, 1.0 };

/* sgn = +1 selects table 0, -1 the negated-cosine table 1 (sincosf_data.c) */
static inline float cos_poly(double x2, double sgn) {
    double x4 = x2 * x2;
    double c1 = fma(x2, sgn * t_c1, sgn * t_c0);
    double c2 = fma(x2, sgn * t_c4, sgn * t_c3);
    double x6 = x2 * x4;
    double c = fma(x4, sgn * t_c2, c1);
    return (float)fma(c2, x6, c);
}
static inline float sin_poly(double xs, double x2) {
    double s1 = fma(x2, t_s3, t_s2);
    double x3 = x2 * xs;
    double x5 = x2 * x3;
    double s = fma(x3, t_s1, xs);
    return (float)fma(s1, x5, s);
}

ORACLE_API float oracle_cosf(float y) {
    double x = (double)y;
    uint32_t top = (f2u(y) >> 20) & 0x7ff;
    if (top <= 0x3f3) {
        double x2 = x * x;
        if (top <= 0x397) return 1.0f;
        return cos_poly(x2, 1.0);
    }
    if (top <= 0x42e) {
        double r = x * t_hpi_inv;
        int n = ((int32_t)r + 0x800000) >> 24;
        double xr = fma(-(double)n, t_hpi, x);
        double x2 = xr * xr;
        if (n & 1) return sin_poly(xr * t_sign[n & 3], x2);
        return cos_poly(x2, (n & 2) ? -1.0 : 1.0);
    }
    return cosf(y);   /* not reached for theta in [0, pi] */
}

/* ------------------------------------------------------------------------------------- */
/* walls                                                                                  */
/* ------------------------------------------------------------------------------------- */
typedef struct {
    int n_walls, max_corners;
    const float *normal;        /* [W][3] */
    const float *origin;        /* [W][3] */
    const float *basis;         /* [W][3][2] row-major */
    const float *flat_corners;  /* [W][2][max_corners] */
    const int32_t *n_corners;   /* [W] */
    const float *reflection;    /* [W][3][3] row-major */
    const float *impedance;     /* [W][K] */
} walls_t;

static inline float dot3(const float *a, const float *b) {
    return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);   /* Eigen redux order */
}

/* wall.cpp:509-536 */
ORACLE_API int oracle_wall_reflect(const float *normal, const float *origin, const float *p,
                                   float *p_reflected) {
    float d3[3] = { origin[0] - p[0], origin[1] - p[1], origin[2] - p[2] };
    float dist = dot3(normal, d3);
    float two_d = 2 * dist;
    for (int i = 0; i < 3; ++i) p_reflected[i] = p[i] + two_d * normal[i];
    if (dist > EPS) return 1;
    if (dist < -EPS) return -1;
    return 0;
}

/* wall.cpp:541-552 */
ORACLE_API int oracle_wall_side(const float *normal, const float *origin, const float *p) {
    float d3[3] = { p[0] - origin[0], p[1] - origin[1], p[2] - origin[2] };
    float ip = dot3(d3, normal);
    if (ip > EPS) return 1;
    if (ip < -EPS) return -1;
    return 0;
}

/* geometry.cpp:54-80 */
static int ccw3p(const float *p1, const float *p2, const float *p3) {
    float d = (p2[0] - p1[0]) * (p3[1] - p1[1]) - (p3[0] - p1[0]) * (p2[1] - p1[1]);
    if (d < EPS && d > -EPS) return 0;
    return d > 0.f ? 1 : -1;
}
/* geometry.cpp:237-247 */
static int on_segment(const float *c1, const float *c2, const float *p) {
    float xd = fminf(c1[0], c2[0]), xu = fmaxf(c1[0], c2[0]);
    float yd = fminf(c1[1], c2[1]), yu = fmaxf(c1[1], c2[1]);
    return xd <= p[0] && p[0] <= xu && yd <= p[1] && p[1] <= yu;
}
/* geometry.cpp:249-269 */
static int is_left(const float *P0, const float *P1, const float *P2) {
    float t = (P1[0] - P0[0]) * (P2[1] - P0[1]) - (P2[0] - P0[0]) * (P1[1] - P0[1]);
    if (fabsf(t) < EPS) return 0;
    return t > 0 ? 1 : -1;
}
/* geometry.cpp:271-323; corners = [2][stride] */
static int inside_polygon(const float *pt, const float *corners, int n, int stride) {
    int wn = 0;
    for (int i = 0, j = n - 1; i < n; j = i++) {
        float cj[2] = { corners[j], corners[stride + j] };
        float ci[2] = { corners[i], corners[stride + i] };
        if (ccw3p(cj, ci, pt) == 0 && on_segment(cj, ci, pt)) return 1;
        if (cj[1] <= pt[1]) {
            if (ci[1] > pt[1])
                if (is_left(cj, ci, pt) > 0) ++wn;
        } else {
            if (ci[1] <= pt[1])
                if (is_left(cj, ci, pt) < 0) --wn;
        }
    }
    return wn == 0 ? -1 : 0;
}
/* geometry.cpp:175-228 */
static int seg_plane(const float *a1, const float *a2, const float *p, const float *normal,
                     float *inter) {
    float u[3] = { a2[0] - a1[0], a2[1] - a1[1], a2[2] - a1[2] };
    float denom = dot3(normal, u);
    if (fabsf(denom) > EPS) {
        float w[3] = { a1[0] - p[0], a1[1] - p[1], a1[2] - p[2] };
        float num = -dot3(normal, w);
        float s = num / denom;
        if (0 - EPS <= s && s <= 1 + EPS) {
            for (int i = 0; i < 3; ++i) inter[i] = s * u[i] + a1[i];
            if (fabsf(s) < EPS || fabsf(s - 1) < EPS) return 1;
            return 0;
        }
    }
    return -1;
}
/* wall.cpp:443-497 */
static int wall_intersection(const walls_t *W, int w, const float *p1, const float *p2,
                             float *inter) {
    const float *origin = W->origin + 3 * w, *normal = W->normal + 3 * w;
    const float *basis = W->basis + 6 * w;
    int ret = 0;
    int ret1 = seg_plane(p1, p2, origin, normal, inter);
    if (ret1 == -1) return -1;
    if (ret1 == 1) ret = 1;
    float d3[3] = { inter[0] - origin[0], inter[1] - origin[1], inter[2] - origin[2] };
    float b0[3] = { basis[0], basis[2], basis[4] }, b1[3] = { basis[1], basis[3], basis[5] };
    float flat[2] = { dot3(b0, d3), dot3(b1, d3) };
    int ret2 = inside_polygon(flat, W->flat_corners + (size_t)2 * W->max_corners * w,
                              W->n_corners[w], W->max_corners);
    if (ret2 < 0) return -1;
    if (ret2 == 1) ret |= 2;
    return ret;
}
ORACLE_API int oracle_wall_intersection(int n_walls, int max_corners, const float *normal,
                                        const float *origin, const float *basis,
                                        const float *flat_corners, const int32_t *n_corners, int w,
                                        const float *p1, const float *p2, float *inter) {
    walls_t W = { n_walls, max_corners, normal, origin, basis, flat_corners, n_corners, 0, 0 };
    return wall_intersection(&W, w, p1, p2, inter);
}

/* wall.cpp:633-635 for one band */
ORACLE_API float oracle_wall_attenuation(float z, float theta) {
    float c = oracle_cosf(theta);
    return (z * c - 1) / (z * c + 1);
}

/* ------------------------------------------------------------------------------------- */
/* DFS (room.cpp:351-420), visibility (room.cpp:426-481), attenuation (room.cpp:484-511)   */
/* ------------------------------------------------------------------------------------- */
#define MAX_ORDER 64
typedef struct {
    float loc[3];
    int order, gen_wall, parent;   /* parent = index on the DFS stack */
} node_t;

typedef struct {
    const walls_t *W;
    const float *mic;
    int K;
    node_t stack[MAX_ORDER + 2];
    /* outputs (grown) */
    long n, cap;
    float *sources;       /* [n][3] */
    int32_t *orders, *gen_walls;
    float *cos_theta;     /* [n][MAX_ORDER]: acosf() argument per path step, top first */
    int32_t *path;        /* [n][MAX_ORDER] wall per step, top (generating wall) first */
    long n_nodes;         /* tree nodes visited (statistics) */
} dfs_t;

static int grow(dfs_t *D) {
    if (D->n < D->cap) return 0;
    long cap = D->cap ? 2 * D->cap : 1024;
    D->sources = (float *)realloc(D->sources, sizeof(float) * 3 * cap);
    D->orders = (int32_t *)realloc(D->orders, sizeof(int32_t) * cap);
    D->gen_walls = (int32_t *)realloc(D->gen_walls, sizeof(int32_t) * cap);
    D->cos_theta = (float *)realloc(D->cos_theta, sizeof(float) * MAX_ORDER * cap);
    D->path = (int32_t *)realloc(D->path, sizeof(int32_t) * MAX_ORDER * cap);
    D->cap = cap;
    return (D->sources && D->orders && D->gen_walls && D->cos_theta && D->path) ? 0 : -1;
}

/* returns 1 if the image at stack[depth] is visible from the mic; fills the per-step acos
 * arguments v.n/|v| (v = image - hit point, room.cpp:452 + 498-500) */
static int visible(dfs_t *D, int depth, float *cos_out, int32_t *path_out) {
    float p[3] = { D->mic[0], D->mic[1], D->mic[2] };
    int step = 0;
    for (int d = depth; d > 0; --d, ++step) {
        const node_t *is = &D->stack[d];
        float hit[3];
        int ret = wall_intersection(D->W, is->gen_wall, p, is->loc, hit);
        if (ret < 0) return 0;
        float v[3] = { is->loc[0] - hit[0], is->loc[1] - hit[1], is->loc[2] - hit[2] };
        const float *nrm = D->W->normal + 3 * is->gen_wall;
        float nv = sqrtf(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]));
        cos_out[step] = dot3(v, nrm) / nv;
        path_out[step] = is->gen_wall;
        p[0] = hit[0]; p[1] = hit[1]; p[2] = hit[2];
    }
    return 1;
}

static int dfs(dfs_t *D, int depth, int max_order) {
    node_t *is = &D->stack[depth];
    D->n_nodes++;
    float cs[MAX_ORDER];
    int32_t path[MAX_ORDER];
    if (visible(D, depth, cs, path)) {
        if (grow(D)) return -1;
        long i = D->n++;
        memcpy(D->sources + 3 * i, is->loc, sizeof(float) * 3);
        D->orders[i] = is->order;
        D->gen_walls[i] = is->gen_wall;
        memcpy(D->cos_theta + MAX_ORDER * i, cs, sizeof(float) * (size_t)depth);
        memcpy(D->path + MAX_ORDER * i, path, sizeof(int32_t) * (size_t)depth);
    }
    if (max_order == 0) return 0;
    for (int w = 0; w < D->W->n_walls; ++w) {
        node_t *child = &D->stack[depth + 1];
        int dir = oracle_wall_reflect(D->W->normal + 3 * w, D->W->origin + 3 * w, is->loc,
                                      child->loc);
        if (dir <= 0) continue;
        child->order = is->order + 1;
        child->gen_wall = w;
        child->parent = depth;
        if (dfs(D, depth + 1, max_order - 1)) return -1;
    }
    return 0;
}

/* Opaque handle API: run -> sizes -> export -> free */
ORACLE_API void *oracle_convex_run(int n_walls, int max_corners, const float *normal,
                                   const float *origin, const float *basis,
                                   const float *flat_corners, const int32_t *n_corners,
                                   const float *reflection, const float *impedance, int K,
                                   const float *source, const float *mic, int ism_order,
                                   long *n_images, long *n_nodes) {
    if (ism_order > MAX_ORDER - 1) return NULL;
    walls_t *W = (walls_t *)malloc(sizeof(walls_t));
    dfs_t *D = (dfs_t *)calloc(1, sizeof(dfs_t));
    if (!W || !D) return NULL;
    W->n_walls = n_walls; W->max_corners = max_corners; W->normal = normal; W->origin = origin;
    W->basis = basis; W->flat_corners = flat_corners; W->n_corners = n_corners;
    W->reflection = reflection; W->impedance = impedance;
    D->W = W; D->mic = mic; D->K = K;
    memcpy(D->stack[0].loc, source, sizeof(float) * 3);
    D->stack[0].order = 0; D->stack[0].gen_wall = -1; D->stack[0].parent = -1;
    if (dfs(D, 0, ism_order)) return NULL;
    *n_images = D->n;
    if (n_nodes) *n_nodes = D->n_nodes;
    return D;
}

/* sources [3][n], orders [n], gen_walls [n], attenuations [K][n], reflection [n][3][3] */
ORACLE_API void oracle_convex_export(void *handle, float *sources, int32_t *orders,
                                     int32_t *gen_walls, float *attenuations, float *reflection) {
    dfs_t *D = (dfs_t *)handle;
    const walls_t *W = D->W;
    long n = D->n;
    int K = D->K;
    for (long i = 0; i < n; ++i) {
        for (int c = 0; c < 3; ++c) sources[c * n + i] = D->sources[3 * i + c];
        orders[i] = D->orders[i];
        gen_walls[i] = D->gen_walls[i];
        int depth = D->orders[i];
        /* attenuation = a_1 * (a_2 * (... * (a_d * 1))), room.cpp:484-511 */
        float theta[MAX_ORDER];
        for (int s = 0; s < depth; ++s) theta[s] = oracle_acosf(D->cos_theta[MAX_ORDER * i + s]);
        for (int k = 0; k < K; ++k) {
            float t = 1.0f;
            for (int s = depth - 1; s >= 0; --s) {
                float z = W->impedance[(size_t)D->path[MAX_ORDER * i + s] * K + k];
                t = oracle_wall_attenuation(z, theta[s]) * t;
            }
            attenuations[(size_t)k * n + i] = t;
        }
        /* reflection matrix = R_{w_top} * ( ... * (R_{w_first} * I)), room.cpp:412 */
        float M[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        for (int s = depth - 1; s >= 0; --s) {
            const float *R = W->reflection + 9 * D->path[MAX_ORDER * i + s];
            float T[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    float a[3] = { R[3 * r], R[3 * r + 1], R[3 * r + 2] };
                    float b[3] = { M[c], M[3 + c], M[6 + c] };
                    T[3 * r + c] = dot3(a, b);
                }
            memcpy(M, T, sizeof(M));
        }
        memcpy(reflection + 9 * i, M, sizeof(M));
    }
}

ORACLE_API void oracle_convex_free(void *handle) {
    dfs_t *D = (dfs_t *)handle;
    if (!D) return;
    free((void *)D->W);
    free(D->sources); free(D->orders); free(D->gen_walls); free(D->cos_theta); free(D->path);
    free(D);
}
