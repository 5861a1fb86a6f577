// convex.cu — convex-room (DEISM-ARG) reflection-tree enumeration on the GPU.
//
// Replaces Room_deism<3>::image_source_model and its helpers in the reference's native
// extension libroom_deism (pybind11 + Eigen float32):
//   image_sources_dfs room.cpp:351-420, is_visible_dfs room.cpp:426-481,
//   get_image_attenuation room.cpp:484-511, fill_sources room.cpp:300-345,
//   Wall_deism<3>::reflect/side/intersection/get_attenuation wall.cpp:443-552,633-635,
//   intersection_3d_segment_plane / ccw3p / is_left / is_inside_2d_polygon geometry.cpp:54-323.
//
// The reference walks the tree depth-first on one CPU thread.  Here the tree is expanded
// level by level (one thread per node, one bit per candidate wall, block prefix sums + a
// scan of block totals compact the children into the next level), every node runs the
// visibility walk independently, and the reference's DFS PRE-ORDER output position of each
// visible image is recovered from subtree visible-counts (bottom-up integer sums, top-down
// sibling prefix), so sources/orders/gen_walls come out in exactly the reference's order.
//
// Numerics: everything is float32 with the reference's evaluation order — Eigen's unrolled
// 3-element redux x0 + (x1 + x2), no FMA contraction (explicit _rn intrinsics) — and
// acosf/cosf are transcriptions of the glibc 2.39 x86-64 routines the reference binary calls
// (fdlibm __ieee754_acosf; FMA build of the ARM-optimized-routines cosf), so positions, orders,
// generating walls, reflection matrices AND float32 attenuations are bit-identical.
#include <string.h>

#include "common.cuh"

namespace deism {

constexpr int CVX_MAX_WALLS = 64;
constexpr int CVX_MAX_CORNERS = 16;
constexpr int CVX_MAX_ORDER = 48;
constexpr int CVX_THREADS = 256;
#define CVX_EPS 1e-5f

struct CvxWalls {
    int n_walls;
    float normal[CVX_MAX_WALLS][3];
    float origin[CVX_MAX_WALLS][3];
    float basis[CVX_MAX_WALLS][6];          // row-major 3x2
    float flat[CVX_MAX_WALLS][2][CVX_MAX_CORNERS];
    int n_corners[CVX_MAX_WALLS];
    float refl[CVX_MAX_WALLS][9];
};
__constant__ CvxWalls c_walls;
__constant__ float c_mic[3];

// ---- float32 helpers with the reference's rounding points --------------------------------
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __fadd_rn(__fmul_rn(a0, b0), __fadd_rn(__fmul_rn(a1, b1), __fmul_rn(a2, b2)));
}

// glibc 2.39 __ieee754_acosf
__device__ float acosf_glibc(float x) {
    const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f,
                pio2_lo = 7.5497894159e-08f, pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f,
                pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f,
                pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
                qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    int hx = __float_as_int(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : __fadd_rn(pi, __fmul_rn(2.0f, pio2_lo));
    if (ix > 0x3f800000) return __int_as_float(0x7fc00000);
    float z;
    if (ix < 0x3f000000) {
        if (ix <= 0x32800000) return __fadd_rn(pio2_hi, pio2_lo);
        z = __fmul_rn(x, x);
    } else if (hx < 0) {
        z = __fmul_rn(__fadd_rn(one, x), 0.5f);
    } else {
        z = __fmul_rn(__fsub_rn(one, x), 0.5f);
    }
    float p = __fadd_rn(pS4, __fmul_rn(z, pS5));
    p = __fadd_rn(pS3, __fmul_rn(z, p));
    p = __fadd_rn(pS2, __fmul_rn(z, p));
    p = __fadd_rn(pS1, __fmul_rn(z, p));
    p = __fadd_rn(pS0, __fmul_rn(z, p));
    p = __fmul_rn(z, p);
    float q = __fadd_rn(qS3, __fmul_rn(z, qS4));
    q = __fadd_rn(qS2, __fmul_rn(z, q));
    q = __fadd_rn(qS1, __fmul_rn(z, q));
    q = __fadd_rn(one, __fmul_rn(z, q));
    float r = __fdiv_rn(p, q);
    if (ix < 0x3f000000)
        return __fsub_rn(pio2_hi, __fsub_rn(x, __fsub_rn(pio2_lo, __fmul_rn(x, r))));
    float s = __fsqrt_rn(z);
    if (hx < 0) {
        float w = __fsub_rn(__fmul_rn(r, s), pio2_lo);
        return __fsub_rn(pi, __fmul_rn(2.0f, __fadd_rn(s, w)));
    }
    float df = __int_as_float(__float_as_int(s) & 0xfffff000);
    float c = __fdiv_rn(__fsub_rn(z, __fmul_rn(df, df)), __fadd_rn(s, df));
    float w = __fadd_rn(__fmul_rn(r, s), c);
    return __fmul_rn(2.0f, __fadd_rn(df, w));
}

// glibc 2.39 __cosf_fma, |x| < 120 (theta = acosf(.) is in [0, pi])
__device__ float cosf_glibc(float y) {
    const double hpi_inv = 0x1.45f306dc9c883p+23, hpi = 0x1.921fb54442d18p+0;
    const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5,
                 c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    double x = (double)y;
    unsigned top = ((unsigned)__float_as_int(y) >> 20) & 0x7ffu;
    double sgn = 1.0, xr = x;
    int n = 0;
    if (top <= 0x3f3u) {
        if (top <= 0x397u) return 1.0f;
    } else {
        double r = __dmul_rn(x, hpi_inv);
        n = (__double2int_rz(r) + 0x800000) >> 24;
        xr = __fma_rn(-(double)n, hpi, x);
        if (n & 2) sgn = -1.0;
    }
    double x2 = __dmul_rn(xr, xr);
    if (n & 1) {
        const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        double xs = __dmul_rn(xr, sg);
        double t1 = __fma_rn(x2, s3, s2);
        double x3 = __dmul_rn(x2, xs);
        double x5 = __dmul_rn(x2, x3);
        double s = __fma_rn(x3, s1, xs);
        return __double2float_rn(__fma_rn(t1, x5, s));
    }
    double x4 = __dmul_rn(x2, x2);
    double a1 = __fma_rn(x2, sgn * c1, sgn * c0);
    double a2 = __fma_rn(x2, sgn * c4, sgn * c3);
    double x6 = __dmul_rn(x2, x4);
    double c = __fma_rn(x4, sgn * c2, a1);
    return __double2float_rn(__fma_rn(a2, x6, c));
}

// wall.cpp:509-536 — returns dir; child location in out
__device__ __forceinline__ int wall_reflect(int w, float px, float py, float pz, float *out) {
    const float *n = c_walls.normal[w], *o = c_walls.origin[w];
    float d = dot3(n[0], n[1], n[2], __fsub_rn(o[0], px), __fsub_rn(o[1], py), __fsub_rn(o[2], pz));
    float two_d = __fmul_rn(2.0f, d);
    out[0] = __fadd_rn(px, __fmul_rn(two_d, n[0]));
    out[1] = __fadd_rn(py, __fmul_rn(two_d, n[1]));
    out[2] = __fadd_rn(pz, __fmul_rn(two_d, n[2]));
    return d > CVX_EPS ? 1 : (d < -CVX_EPS ? -1 : 0);
}

__device__ __forceinline__ float cross2(float ax, float ay, float bx, float by, float cx, float cy) {
    // (b.x-a.x)*(c.y-a.y) - (c.x-a.x)*(b.y-a.y)   (geometry.cpp:71-72, 260-263)
    return __fsub_rn(__fmul_rn(__fsub_rn(bx, ax), __fsub_rn(cy, ay)),
                     __fmul_rn(__fsub_rn(cx, ax), __fsub_rn(by, ay)));
}

// geometry.cpp:271-323
__device__ int inside_polygon(int w, float px, float py) {
    const int n = c_walls.n_corners[w];
    const float *X = c_walls.flat[w][0], *Y = c_walls.flat[w][1];
    int wn = 0;
    for (int i = 0, j = n - 1; i < n; j = i++) {
        float cjx = X[j], cjy = Y[j], cix = X[i], ciy = Y[i];
        float d = cross2(cjx, cjy, cix, ciy, px, py);
        bool collinear = d < CVX_EPS && d > -CVX_EPS;
        if (collinear) {
            float xd = fminf(cjx, cix), xu = fmaxf(cjx, cix), yd = fminf(cjy, ciy), yu = fmaxf(cjy, ciy);
            if (xd <= px && px <= xu && yd <= py && py <= yu) return 1;
        }
        int left = fabsf(d) < CVX_EPS ? 0 : (d > 0.f ? 1 : -1);
        if (cjy <= py) {
            if (ciy > py && left > 0) ++wn;
        } else {
            if (ciy <= py && left < 0) --wn;
        }
    }
    return wn == 0 ? -1 : 0;
}

// wall.cpp:443-497 + geometry.cpp:175-228; p1 -> p2 segment against wall w
__device__ int wall_intersection(int w, const float *p1, const float *p2, float *inter) {
    const float *nrm = c_walls.normal[w], *org = c_walls.origin[w], *B = c_walls.basis[w];
    float ux = __fsub_rn(p2[0], p1[0]), uy = __fsub_rn(p2[1], p1[1]), uz = __fsub_rn(p2[2], p1[2]);
    float denom = dot3(nrm[0], nrm[1], nrm[2], ux, uy, uz);
    if (!(fabsf(denom) > CVX_EPS)) return -1;
    float num = -dot3(nrm[0], nrm[1], nrm[2], __fsub_rn(p1[0], org[0]), __fsub_rn(p1[1], org[1]),
                      __fsub_rn(p1[2], org[2]));
    float s = __fdiv_rn(num, denom);
    if (!(__fsub_rn(0.f, CVX_EPS) <= s && s <= __fadd_rn(1.f, CVX_EPS))) return -1;
    inter[0] = __fadd_rn(__fmul_rn(s, ux), p1[0]);
    inter[1] = __fadd_rn(__fmul_rn(s, uy), p1[1]);
    inter[2] = __fadd_rn(__fmul_rn(s, uz), p1[2]);
    int ret = (fabsf(s) < CVX_EPS || fabsf(__fsub_rn(s, 1.f)) < CVX_EPS) ? 1 : 0;
    float dx = __fsub_rn(inter[0], org[0]), dy = __fsub_rn(inter[1], org[1]), dz = __fsub_rn(inter[2], org[2]);
    float fx = dot3(B[0], B[2], B[4], dx, dy, dz), fy = dot3(B[1], B[3], B[5], dx, dy, dz);
    int r2 = inside_polygon(w, fx, fy);
    if (r2 < 0) return -1;
    if (r2 == 1) ret |= 2;
    return ret;
}

// ---- node storage -------------------------------------------------------------------------
struct CvxNodes {
    float4 *loc_par;       // x, y, z, parent index (int bits)
    int *meta;             // gen_wall (low 8 bits, 255 = none) | order << 8
    int *child_start;      // global index of the first child
    unsigned long long *mask;   // candidate walls with dir > 0
    int *cnt;              // visible nodes in the subtree
    int *pos;              // pre-order output position of the first visible node of the subtree
    unsigned char *vis;
};

// level expansion, pass 1: candidate mask + block totals
__global__ void __launch_bounds__(CVX_THREADS)
cvx_count_kernel(CvxNodes N, long lo, long hi, int *block_sums) {
    __shared__ int wsum[CVX_THREADS / 32];
    long i = lo + (long)blockIdx.x * CVX_THREADS + threadIdx.x;
    int c = 0;
    if (i < hi) {
        float4 lp = N.loc_par[i];
        unsigned long long m = 0ull;
        float tmp[3];
        for (int w = 0; w < c_walls.n_walls; ++w)
            if (wall_reflect(w, lp.x, lp.y, lp.z, tmp) > 0) m |= 1ull << w;
        N.mask[i] = m;
        c = __popcll(m);
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CVX_THREADS / 32; ++w) t += wsum[w];
        block_sums[blockIdx.x] = t;
    }
}

// exclusive scan of block totals (one CTA); total -> out_total
__global__ void __launch_bounds__(1024) cvx_scan_blocks_kernel(int *block_sums, long n, long *out_total) {
    __shared__ long warp_tot[32];
    __shared__ long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long base = 0; base < n; base += 1024) {
        long i = base + threadIdx.x;
        long v = i < n ? block_sums[i] : 0;
        long x = v;
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        long woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        long excl = carry + woff + x - v;
        __syncthreads();
        if (i < n) block_sums[i] = (int)excl;
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_total = carry;
}

// level expansion, pass 2: write the children (node order, then wall order)
__global__ void __launch_bounds__(CVX_THREADS)
cvx_fill_kernel(CvxNodes N, long lo, long hi, const int *block_offsets, long next_lo) {
    __shared__ int wsum[CVX_THREADS / 32];
    long i = lo + (long)blockIdx.x * CVX_THREADS + threadIdx.x;
    unsigned long long m = i < hi ? N.mask[i] : 0ull;
    int c = __popcll(m);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = c;
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    if (i >= hi) return;
    long start = next_lo + block_offsets[blockIdx.x] + woff + x - c;
    N.child_start[i] = (int)start;
    float4 lp = N.loc_par[i];
    int order = (N.meta[i] >> 8) + 1;
    long k = start;
    while (m) {
        int w = __ffsll((long long)m) - 1;
        m &= m - 1;
        float out[3];
        wall_reflect(w, lp.x, lp.y, lp.z, out);
        N.loc_par[k] = make_float4(out[0], out[1], out[2], __int_as_float((int)i));
        N.meta[k] = w | (order << 8);
        ++k;
    }
}

// visibility walk (room.cpp:426-481, convex rooms: no obstructing walls)
__device__ bool cvx_visible(const CvxNodes &N, long i, float *cos_out, unsigned char *path_out) {
    float p[3] = {c_mic[0], c_mic[1], c_mic[2]};
    int step = 0;
    long cur = i;
    while (true) {
        int meta = N.meta[cur];
        int gw = meta & 0xff;
        if (gw == 0xff) return true;   // the real source
        float4 lp = N.loc_par[cur];
        float loc[3] = {lp.x, lp.y, lp.z}, hit[3];
        if (wall_intersection(gw, p, loc, hit) < 0) return false;
        if (cos_out) {
            float vx = __fsub_rn(loc[0], hit[0]), vy = __fsub_rn(loc[1], hit[1]), vz = __fsub_rn(loc[2], hit[2]);
            const float *nrm = c_walls.normal[gw];
            float nv = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fadd_rn(__fmul_rn(vy, vy), __fmul_rn(vz, vz))));
            cos_out[step] = __fdiv_rn(dot3(vx, vy, vz, nrm[0], nrm[1], nrm[2]), nv);
            path_out[step] = (unsigned char)gw;
        }
        p[0] = hit[0]; p[1] = hit[1]; p[2] = hit[2];
        cur = __float_as_int(lp.w);
        ++step;
    }
}

__global__ void __launch_bounds__(CVX_THREADS) cvx_visible_kernel(CvxNodes N, long n_nodes) {
    long i = (long)blockIdx.x * CVX_THREADS + threadIdx.x;
    if (i >= n_nodes) return;
    bool v = cvx_visible(N, i, nullptr, nullptr);
    N.vis[i] = v ? 1 : 0;
    N.cnt[i] = v ? 1 : 0;
}

// bottom-up: add the subtree count of every node of a level to its parent
__global__ void __launch_bounds__(CVX_THREADS) cvx_up_kernel(CvxNodes N, long lo, long hi) {
    long i = lo + (long)blockIdx.x * CVX_THREADS + threadIdx.x;
    if (i >= hi) return;
    int c = N.cnt[i];
    if (c) atomicAdd(&N.cnt[__float_as_int(N.loc_par[i].w)], c);
}

// top-down: children positions from the parent's (siblings in wall order)
__global__ void __launch_bounds__(CVX_THREADS) cvx_down_kernel(CvxNodes N, long lo, long hi) {
    long i = lo + (long)blockIdx.x * CVX_THREADS + threadIdx.x;
    if (i >= hi) return;
    int nchild = __popcll(N.mask[i]);
    int p = N.pos[i] + N.vis[i];
    int cs = N.child_start[i];
    for (int c = 0; c < nchild; ++c) {
        N.pos[cs + c] = p;
        p += N.cnt[cs + c];
    }
}

// export of the visible images
struct CvxOut {
    long n_images;
    float *sources;        // [3][I]
    int32_t *orders, *gen_walls;
    float *reflection;     // [I][3][3]
    float *cosv;           // [I][CVX_MAX_ORDER]  cosf(acosf(v.n/|v|)) per path step
    unsigned char *path;   // [I][CVX_MAX_ORDER]
};

__global__ void __launch_bounds__(CVX_THREADS) cvx_export_kernel(CvxNodes N, long n_nodes, CvxOut o) {
    long i = (long)blockIdx.x * CVX_THREADS + threadIdx.x;
    if (i >= n_nodes || !N.vis[i]) return;
    const long idx = N.pos[i];
    const int meta = N.meta[i];
    const int order = meta >> 8, gw = meta & 0xff;
    float4 lp = N.loc_par[i];
    if (o.sources) {
        o.sources[idx] = lp.x; o.sources[o.n_images + idx] = lp.y; o.sources[2 * o.n_images + idx] = lp.z;
    }
    if (o.orders) o.orders[idx] = order;
    if (o.gen_walls) o.gen_walls[idx] = gw == 0xff ? -1 : gw;
    float cs[CVX_MAX_ORDER];
    unsigned char path[CVX_MAX_ORDER];
    cvx_visible(N, i, cs, path);
    for (int s = 0; s < order; ++s) {
        o.cosv[idx * CVX_MAX_ORDER + s] = cosf_glibc(acosf_glibc(cs[s]));
        o.path[idx * CVX_MAX_ORDER + s] = path[s];
    }
    if (o.reflection) {
        // R_top * ( ... * (R_first * I)), entries with Eigen's x0 + (x1 + x2) order (room.cpp:412)
        float M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int s = order - 1; s >= 0; --s) {
            const float *R = c_walls.refl[path[s]];
            float T[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    T[3 * r + c] = dot3(R[3 * r], R[3 * r + 1], R[3 * r + 2], M[c], M[3 + c], M[6 + c]);
#pragma unroll
            for (int q = 0; q < 9; ++q) M[q] = T[q];
        }
        for (int q = 0; q < 9; ++q) o.reflection[idx * 9 + q] = M[q];
    }
}

// attenuations[k][img] = a_1 * (a_2 * (... * (a_d * 1)))   (room.cpp:484-511, wall.cpp:633-635)
__global__ void __launch_bounds__(CVX_THREADS)
cvx_atten_kernel(long n_images, long K, const int32_t *orders, const float *cosv,
                 const unsigned char *path, const float *imp /* [W][K] */, float *out /* [K][I] */) {
    long img = (long)blockIdx.x * CVX_THREADS + threadIdx.x;
    if (img >= n_images) return;
    int depth = orders[img];
    for (long k = blockIdx.y; k < K; k += gridDim.y) {
        float t = 1.0f;
        for (int s = depth - 1; s >= 0; --s) {
            float zc = __fmul_rn(imp[(long)path[img * CVX_MAX_ORDER + s] * K + k], cosv[img * CVX_MAX_ORDER + s]);
            float a = __fdiv_rn(__fsub_rn(zc, 1.0f), __fadd_rn(zc, 1.0f));
            t = __fmul_rn(a, t);
        }
        out[k * n_images + img] = t;
    }
}

// ---- growable device buffers that keep their contents --------------------------------------
struct CvxState {
    CvxNodes N = {};
    long cap = 0, n_nodes = 0, n_images = 0;
    int n_levels = 0;
    long level_start[CVX_MAX_ORDER + 2] = {};
    int *block_sums = nullptr;
    long block_cap = 0;
    long *d_total = nullptr;
    float *cosv = nullptr;
    unsigned char *path = nullptr;
    int32_t *orders_tmp = nullptr;
    long out_cap = 0;
    float *imp = nullptr;
    long imp_cap = 0;
    bool valid = false;
};
// one state per device (the pointers belong to the device that was current when they were
// allocated); the per-device CallGuard serialises access
static CvxState g_cvx_dev[64];
static CvxState *cvx_state() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    return &g_cvx_dev[dev];
}

template <typename T> static int grow_keep(T *&ptr, long old_n, long new_n) {
    T *np_ = nullptr;
    if (cudaMalloc(&np_, (size_t)new_n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        return set_error(DEISM_ERR_NOMEM, "cudaMalloc of %ld elements failed (convex tree)", new_n);
    }
    if (ptr && old_n > 0) cudaMemcpy(np_, ptr, (size_t)old_n * sizeof(T), cudaMemcpyDeviceToDevice);
    if (ptr) cudaFree(ptr);
    ptr = np_;
    return DEISM_OK;
}

static int ensure_nodes(CvxState &S, long need) {
    if (need <= S.cap) return DEISM_OK;
    long nc = S.cap ? S.cap : (1L << 16);
    while (nc < need) nc *= 2;
    long keep = S.n_nodes;
    int rc;
    if ((rc = grow_keep(S.N.loc_par, keep, nc))) return rc;
    if ((rc = grow_keep(S.N.meta, keep, nc))) return rc;
    if ((rc = grow_keep(S.N.child_start, keep, nc))) return rc;
    if ((rc = grow_keep(S.N.mask, keep, nc))) return rc;
    if ((rc = grow_keep(S.N.cnt, keep, nc))) return rc;
    if ((rc = grow_keep(S.N.pos, keep, nc))) return rc;
    if ((rc = grow_keep(S.N.vis, keep, nc))) return rc;
    S.cap = nc;
    return DEISM_OK;
}

static int upload_walls(const deism_convex_walls *w, cudaStream_t st) {
    if (!w || !w->h_normal || !w->h_origin || !w->h_basis || !w->h_flat_corners || !w->h_n_corners ||
        !w->h_reflection)
        return set_error(DEISM_ERR_INVALID, "null wall table");
    if (w->n_walls < 4 || w->n_walls > CVX_MAX_WALLS)
        return set_error(DEISM_ERR_INVALID, "n_walls must be 4..%d (room.cpp:262-264)", CVX_MAX_WALLS);
    static CvxWalls h;   // 30 KB: keep it off the stack
    memset(&h, 0, sizeof(h));
    h.n_walls = w->n_walls;
    for (int i = 0; i < w->n_walls; ++i) {
        int nc = w->h_n_corners[i];
        if (nc < 3 || nc > CVX_MAX_CORNERS || nc > w->max_corners)
            return set_error(DEISM_ERR_INVALID, "wall %d has %d corners (3..%d supported)", i, nc,
                             CVX_MAX_CORNERS);
        for (int c = 0; c < 3; ++c) { h.normal[i][c] = w->h_normal[3 * i + c]; h.origin[i][c] = w->h_origin[3 * i + c]; }
        for (int c = 0; c < 6; ++c) h.basis[i][c] = w->h_basis[6 * i + c];
        for (int c = 0; c < 9; ++c) h.refl[i][c] = w->h_reflection[9 * i + c];
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < nc; ++c)
                h.flat[i][r][c] = w->h_flat_corners[((size_t)i * 2 + r) * w->max_corners + c];
        h.n_corners[i] = nc;
    }
    // on the caller's stream (non-blocking streams are not ordered against the legacy stream, and
    // the previous fill's kernels may still be reading c_walls); the caller synchronises `st`
    // before `h` can be rewritten
    bump_generation();
    DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_walls, &h, sizeof(h), 0, cudaMemcpyHostToDevice, st));
    return DEISM_OK;
}

}  // namespace deism

using namespace deism;

extern "C" int deism_convex_images_count(const deism_convex_walls *walls, const float h_source[3],
                                         const float h_mic[3], int ism_order, int64_t *h_n_images,
                                         void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (!h_source || !h_mic || !h_n_images) return set_error(DEISM_ERR_INVALID, "null pointer");
    if (ism_order < 0 || ism_order >= CVX_MAX_ORDER)
        return set_error(DEISM_ERR_INVALID, "ism_order must be 0..%d", CVX_MAX_ORDER - 1);
    rc = upload_walls(walls, st);
    if (rc) return rc;
    bump_generation();
    DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_mic, h_mic, 3 * sizeof(float), 0, cudaMemcpyHostToDevice, st));
    CvxState *Sp = cvx_state();
    if (!Sp) return set_error(DEISM_ERR_CUDA, "no current device");
    CvxState &S = *Sp;
    S.valid = false;
    S.n_nodes = 0;
    rc = ensure_nodes(S, 1L << 16);
    if (rc) return rc;
    if (!S.d_total) DEISM_CUDA_TRY(cudaMalloc(&S.d_total, sizeof(long)));
    // root = the real source
    float minus_one_bits;
    {
        int m1 = -1;
        memcpy(&minus_one_bits, &m1, sizeof(float));
    }
    float4 root = make_float4(h_source[0], h_source[1], h_source[2], minus_one_bits);
    int root_meta = 0xff;
    DEISM_CUDA_TRY(cudaMemcpyAsync(S.N.loc_par, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    DEISM_CUDA_TRY(cudaMemcpyAsync(S.N.meta, &root_meta, sizeof(int), cudaMemcpyHostToDevice, st));
    DEISM_CUDA_TRY(cudaStreamSynchronize(st));
    S.n_nodes = 1;
    S.level_start[0] = 0;
    S.level_start[1] = 1;
    S.n_levels = 1;
    prof_begin("convex_expand", st);
    for (int d = 0; d < ism_order; ++d) {
        long lo = S.level_start[d], hi = S.level_start[d + 1];
        long n = hi - lo;
        if (n == 0) break;
        long nb = (n + CVX_THREADS - 1) / CVX_THREADS;
        if (nb > S.block_cap) {
            if (S.block_sums) cudaFree(S.block_sums);
            S.block_sums = nullptr;
            S.block_cap = 0;
            DEISM_CUDA_TRY(cudaMalloc(&S.block_sums, (size_t)nb * 2 * sizeof(int)));
            S.block_cap = nb * 2;
        }
        cvx_count_kernel<<<(unsigned)nb, CVX_THREADS, 0, st>>>(S.N, lo, hi, S.block_sums);
        DEISM_LAUNCH_CHECK("cvx_count_kernel");
        cvx_scan_blocks_kernel<<<1, 1024, 0, st>>>(S.block_sums, nb, S.d_total);
        DEISM_LAUNCH_CHECK("cvx_scan_blocks_kernel");
        long total = 0;
        DEISM_CUDA_TRY(cudaMemcpyAsync(&total, S.d_total, sizeof(long), cudaMemcpyDeviceToHost, st));
        DEISM_CUDA_TRY(cudaStreamSynchronize(st));
        if (hi + total > 0x7fffffffL)
            return set_error(DEISM_ERR_INVALID, "reflection tree exceeds 2^31 nodes at order %d", d + 1);
        rc = ensure_nodes(S, hi + total);
        if (rc) return rc;
        cvx_fill_kernel<<<(unsigned)nb, CVX_THREADS, 0, st>>>(S.N, lo, hi, S.block_sums, hi);
        DEISM_LAUNCH_CHECK("cvx_fill_kernel");
        S.n_nodes = hi + total;
        S.level_start[d + 2] = S.n_nodes;
        S.n_levels = d + 2;
    }
    prof_end("convex_expand", st);
    // leaves of the last level have no children
    {
        long lo = S.level_start[S.n_levels - 1], hi = S.level_start[S.n_levels];
        if (hi > lo) DEISM_CUDA_TRY(cudaMemsetAsync(S.N.mask + lo, 0, (size_t)(hi - lo) * 8, st));
    }
    long nb_all = (S.n_nodes + CVX_THREADS - 1) / CVX_THREADS;
    prof_begin("convex_visible", st);
    cvx_visible_kernel<<<(unsigned)nb_all, CVX_THREADS, 0, st>>>(S.N, S.n_nodes);
    prof_end("convex_visible", st);
    DEISM_LAUNCH_CHECK("cvx_visible_kernel");
    for (int d = S.n_levels - 1; d >= 1; --d) {
        long lo = S.level_start[d], hi = S.level_start[d + 1];
        if (hi == lo) continue;
        cvx_up_kernel<<<(unsigned)((hi - lo + CVX_THREADS - 1) / CVX_THREADS), CVX_THREADS, 0, st>>>(S.N, lo, hi);
        DEISM_LAUNCH_CHECK("cvx_up_kernel");
    }
    DEISM_CUDA_TRY(cudaMemsetAsync(S.N.pos, 0, sizeof(int), st));
    for (int d = 0; d < S.n_levels - 1; ++d) {
        long lo = S.level_start[d], hi = S.level_start[d + 1];
        if (hi == lo) continue;
        cvx_down_kernel<<<(unsigned)((hi - lo + CVX_THREADS - 1) / CVX_THREADS), CVX_THREADS, 0, st>>>(S.N, lo, hi);
        DEISM_LAUNCH_CHECK("cvx_down_kernel");
    }
    int n_vis = 0;
    DEISM_CUDA_TRY(cudaMemcpyAsync(&n_vis, S.N.cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    DEISM_CUDA_TRY(cudaStreamSynchronize(st));
    S.n_images = n_vis;
    S.valid = true;
    *h_n_images = n_vis;
    return DEISM_OK;
}

extern "C" int deism_convex_images_fill(const deism_convex_walls *walls, int64_t K, float *d_sources,
                                        int32_t *d_orders, int32_t *d_gen_walls, float *d_attenuations,
                                        float *d_reflection, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    cudaStream_t st = (cudaStream_t)stream;
    CvxState *Sp = cvx_state();
    if (!Sp) return set_error(DEISM_ERR_CUDA, "no current device");
    CvxState &S = *Sp;
    if (!S.valid) return set_error(DEISM_ERR_INVALID, "call deism_convex_images_count first");
    if (!walls) return set_error(DEISM_ERR_INVALID, "null walls");
    const long I = S.n_images;
    if (I == 0) return DEISM_OK;
    if (I > S.out_cap) {
        if (S.cosv) cudaFree(S.cosv);
        if (S.path) cudaFree(S.path);
        if (S.orders_tmp) cudaFree(S.orders_tmp);
        S.cosv = nullptr; S.path = nullptr; S.orders_tmp = nullptr;
        S.out_cap = 0;              // stays 0 (and the pointers null) if an allocation below fails
        const long want = I + I / 2;
        DEISM_CUDA_TRY(cudaMalloc(&S.cosv, (size_t)want * CVX_MAX_ORDER * sizeof(float)));
        DEISM_CUDA_TRY(cudaMalloc(&S.path, (size_t)want * CVX_MAX_ORDER));
        DEISM_CUDA_TRY(cudaMalloc(&S.orders_tmp, (size_t)want * sizeof(int32_t)));
        S.out_cap = want;
    }
    CvxOut o;
    o.n_images = I; o.sources = d_sources; o.orders = S.orders_tmp; o.gen_walls = d_gen_walls;
    o.reflection = d_reflection; o.cosv = S.cosv; o.path = S.path;
    long nb_all = (S.n_nodes + CVX_THREADS - 1) / CVX_THREADS;
    cvx_export_kernel<<<(unsigned)nb_all, CVX_THREADS, 0, st>>>(S.N, S.n_nodes, o);
    DEISM_LAUNCH_CHECK("cvx_export_kernel");
    if (d_orders)
        DEISM_CUDA_TRY(cudaMemcpyAsync(d_orders, S.orders_tmp, (size_t)I * sizeof(int32_t),
                                       cudaMemcpyDeviceToDevice, st));
    if (d_attenuations) {
        if (K <= 0) return set_error(DEISM_ERR_INVALID, "K must be positive");
        if (!walls->h_impedance) return set_error(DEISM_ERR_INVALID, "null h_impedance");
        long need = (long)walls->n_walls * K;
        if (need > S.imp_cap) {
            if (S.imp) cudaFree(S.imp);
            S.imp = nullptr;
            S.imp_cap = 0;
            DEISM_CUDA_TRY(cudaMalloc(&S.imp, (size_t)need * sizeof(float)));
            S.imp_cap = need;
        }
        DEISM_CUDA_TRY(cudaMemcpyAsync(S.imp, walls->h_impedance, (size_t)need * sizeof(float),
                                       cudaMemcpyHostToDevice, st));
        DEISM_CUDA_TRY(cudaStreamSynchronize(st));
        dim3 grid((unsigned)((I + CVX_THREADS - 1) / CVX_THREADS), (unsigned)(K < 65535 ? K : 65535));
        cvx_atten_kernel<<<grid, CVX_THREADS, 0, st>>>(I, K, S.orders_tmp, S.cosv, S.path, S.imp,
                                                      d_attenuations);
        DEISM_LAUNCH_CHECK("cvx_atten_kernel");
    }
    return DEISM_OK;
}
