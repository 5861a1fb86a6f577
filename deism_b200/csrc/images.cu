// images.cu — shoebox image-lattice enumeration and wall-attenuation materialisation.
//
// Replaces the reference's Numba generator (deism/core_deism.py, "cd"):
//   _count_shoebox_images_nofs_v2_numba  cd:2649-2728
//   _fill_shoebox_images_nofs_v2_numba   cd:2731-2928
//   prefix sums                          cd:3390-3394
//   direct-path removal                  cd:3442-3451
// and deism/count_reflections.cpp (buffer pre-sizing for the legacy generators).
//
// Mapping: one CTA per task t = parity*(N_o+1) + order (the reference's prange unit).  The
// candidates of a task are linearised in the reference's visiting order
//     c = pair(i_abs, j_abs) * 8 + (i_sign*4 + j_sign*2 + k_sign)
// (i_abs ascending, j_abs ascending, negative sign first), validity = parity test + distance
// cut; kept candidates are compacted with warp ballots + a block prefix sum so that the
// output order is bit-for-bit the reference's.  Geometry is FP64 with explicit _rn intrinsics
// (no FMA contraction, as Numba) and rounded once to float32 (Q1 of SURVEY §8).
#include "common.cuh"

namespace deism {

struct LatticeParams {
    double LL[3], xs[3], xr[3];
    double max_d2;
    int N_o, N_o_ORG, remove_direct;
};

struct Candidate {
    int qx, qy, qz;
    double Isx, Isy, Isz, d2;
};

// Decode candidate c of task (px,py,pz,ord); returns true if it is an image the reference keeps.
__device__ __forceinline__ bool decode_candidate(const LatticeParams &P, int px, int py, int pz,
                                                 int ord, int c, Candidate &out) {
    int pair = c >> 3, sub = c & 7;
    // invert pair -> (i_abs, j_abs): rows of length ord+1, ord, ..., 1
    int i_abs = 0, rem = pair;
    while (rem >= ord - i_abs + 1) { rem -= ord - i_abs + 1; ++i_abs; }
    int j_abs = rem, k_abs = ord - i_abs - j_abs;
    int ii = sub >> 2, jj = (sub >> 1) & 1, kk = sub & 1;
    if ((ii && i_abs == 0) || (jj && j_abs == 0) || (kk && k_abs == 0)) return false;
    int i = i_abs == 0 ? 0 : (ii == 0 ? -i_abs : i_abs);
    int j = j_abs == 0 ? 0 : (jj == 0 ? -j_abs : j_abs);
    int k = k_abs == 0 ? 0 : (kk == 0 ? -k_abs : k_abs);
    if (((i + px) & 1) || ((j + py) & 1) || ((k + pz) & 1)) return false;
    if (P.remove_direct && ord == 0 && (px | py | pz) == 0) return false;
    // (i+p) even -> arithmetic shift is the floor division of cd:2704-2706
    out.qx = (i + px) >> 1; out.qy = (j + py) >> 1; out.qz = (k + pz) >> 1;
    double sox = __dsub_rn(P.xs[0], __dmul_rn((double)(2 * px), P.xs[0]));
    double soy = __dsub_rn(P.xs[1], __dmul_rn((double)(2 * py), P.xs[1]));
    double soz = __dsub_rn(P.xs[2], __dmul_rn((double)(2 * pz), P.xs[2]));
    out.Isx = __dadd_rn(__dmul_rn((double)(2 * out.qx), P.LL[0]), sox);
    out.Isy = __dadd_rn(__dmul_rn((double)(2 * out.qy), P.LL[1]), soy);
    out.Isz = __dadd_rn(__dmul_rn((double)(2 * out.qz), P.LL[2]), soz);
    double dx = __dsub_rn(out.Isx, P.xr[0]), dy = __dsub_rn(out.Isy, P.xr[1]);
    double dz = __dsub_rn(out.Isz, P.xr[2]);
    out.d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return !(out.d2 > P.max_d2);
}

__device__ __forceinline__ void task_decode(int t, int N_o, int &px, int &py, int &pz, int &ord) {
    int n_ref = N_o + 1;
    int par = t / n_ref;
    ord = t % n_ref;
    px = par >> 2; py = (par >> 1) & 1; pz = par & 1;
}

constexpr int IMG_THREADS = 256;

// Block-wide exclusive scan of a 0/1 flag; returns this thread's rank and the block total.
__device__ __forceinline__ int block_rank(bool flag, int *warp_sums, int &total) {
    unsigned ball = __ballot_sync(0xffffffffu, flag);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int rank = __popc(ball & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[warp] = __popc(ball);
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < IMG_THREADS / 32; ++w) {
        int s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + rank;
}

__global__ void __launch_bounds__(IMG_THREADS)
shoebox_count_kernel(LatticeParams P, int64_t *early_counts, int64_t *late_counts) {
    __shared__ int warp_sums[IMG_THREADS / 32];
    int t = blockIdx.x, px, py, pz, ord;
    task_decode(t, P.N_o, px, py, pz, ord);
    int n_cand = ((ord + 1) * (ord + 2) / 2) * 8;
    int count = 0;
    for (int c0 = 0; c0 < n_cand; c0 += IMG_THREADS) {
        int c = c0 + threadIdx.x;
        Candidate cand;
        bool keep = c < n_cand && decode_candidate(P, px, py, pz, ord, c, cand);
        int total;
        block_rank(keep, warp_sums, total);
        count += total;
    }
    if (threadIdx.x == 0) {
        bool early = ord <= P.N_o_ORG;
        early_counts[t] = early ? count : 0;
        late_counts[t] = early ? 0 : count;
    }
}

// cd:2623-2631  (az, polar, r) of one vector, FP64 -> float32
__device__ __forceinline__ void cart2sph_polar_f32(double x, double y, double z, float *out3) {
    double h = hypot(x, y);
    double r = hypot(h, z);
    double el = atan2(z, h);
    double az = atan2(y, x);
    double th = __dsub_rn(3.14159265358979323846 / 2, el);
    out3[0] = __double2float_rn(az);
    out3[1] = __double2float_rn(th);
    out3[2] = __double2float_rn(r);
}

__global__ void __launch_bounds__(IMG_THREADS)
shoebox_fill_kernel(LatticeParams P, const int64_t *__restrict__ early_counts,
                    const int64_t *__restrict__ late_counts, int32_t *A_e, float *Ra_e, float *Rb_e,
                    float *Rc_e, int32_t *A_l, float *Ra_l, float *Rb_l, float *Rc_l) {
    __shared__ int warp_sums[IMG_THREADS / 32];
    __shared__ long long red[IMG_THREADS];
    int t = blockIdx.x, px, py, pz, ord;
    task_decode(t, P.N_o, px, py, pz, ord);
    bool early = ord <= P.N_o_ORG;
    const int64_t *counts = early ? early_counts : late_counts;
    // exclusive prefix over preceding tasks (cd:3390-3394)
    long long part = 0;
    for (int u = threadIdx.x; u < t; u += IMG_THREADS) part += counts[u];
    red[threadIdx.x] = part;
    __syncthreads();
    for (int s = IMG_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    long long base = red[0];
    __syncthreads();

    int32_t *A = early ? A_e : A_l;
    float *Ra = early ? Ra_e : Ra_l, *Rb = early ? Rb_e : Rb_l, *Rc = early ? Rc_e : Rc_l;
    int n_cand = ((ord + 1) * (ord + 2) / 2) * 8;
    for (int c0 = 0; c0 < n_cand; c0 += IMG_THREADS) {
        int c = c0 + threadIdx.x;
        Candidate cd;
        bool keep = c < n_cand && decode_candidate(P, px, py, pz, ord, c, cd);
        int total;
        int rank = block_rank(keep, warp_sums, total);
        if (keep) {
            long long idx = base + rank;
            if (A) {
                int32_t *a = A + idx * 6;
                a[0] = cd.qx; a[1] = cd.qy; a[2] = cd.qz; a[3] = px; a[4] = py; a[5] = pz;
            }
            // receiver image (cd:2816-2818)
            double Irx = __dmul_rn((double)(1 - 2 * px),
                                   __dsub_rn(P.xr[0], __dmul_rn((double)(2 * cd.qx), P.LL[0])));
            double Iry = __dmul_rn((double)(1 - 2 * py),
                                   __dsub_rn(P.xr[1], __dmul_rn((double)(2 * cd.qy), P.LL[1])));
            double Irz = __dmul_rn((double)(1 - 2 * pz),
                                   __dsub_rn(P.xr[2], __dmul_rn((double)(2 * cd.qz), P.LL[2])));
            if (Ra)
                cart2sph_polar_f32(__dsub_rn(P.xr[0], cd.Isx), __dsub_rn(P.xr[1], cd.Isy),
                                   __dsub_rn(P.xr[2], cd.Isz), Ra + idx * 3);
            if (Rb)
                cart2sph_polar_f32(__dsub_rn(Irx, P.xs[0]), __dsub_rn(Iry, P.xs[1]),
                                   __dsub_rn(Irz, P.xs[2]), Rb + idx * 3);
            if (Rc)
                cart2sph_polar_f32(__dsub_rn(cd.Isx, P.xr[0]), __dsub_rn(cd.Isy, P.xr[1]),
                                   __dsub_rn(cd.Isz, P.xr[2]), Rc + idx * 3);
        }
        base += total;
    }
}

// atten[I,K] materialisation: one thread per (image, frequency); K fastest for coalescing.
template <bool OUT_C128>
__global__ void __launch_bounds__(256)
shoebox_atten_kernel(long n_images, long K, int mode, int angdep, const int32_t *__restrict__ A,
                     const float *__restrict__ R_sI_r, const cplx *__restrict__ Z, RoomGeom room,
                     void *out) {
    __shared__ AttenGeom g_sm;
    long img = blockIdx.y;
    if (threadIdx.x == 0) {
        AttenGeom g;
        if (mode == DEISM_ATTEN_FUSED_ROOM)
            atten_geom_room(A + img * 6, room.LL, room.xs, room.xr, angdep, g);
        else
            atten_geom_angles(A + img * 6, R_sI_r + img * 3, angdep, g);
        g_sm = g;
    }
    __syncthreads();
    AttenGeom g = g_sm;
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < K;
         k += (long)gridDim.x * blockDim.x) {
        cplx a = shoebox_atten(Z + k, K, g.cx, g.cy, g.cz, g.e);
        if (OUT_C128) {
            ((cplx *)out)[img * K + k] = a;
        } else {
            ((float2 *)out)[img * K + k] = make_float2(__double2float_rn(a.x), __double2float_rn(a.y));
        }
    }
}

static int fill_lattice(LatticeParams &P, const double *LL, const double *xs, const double *xr,
                        double max_d2, int N_o, int N_o_ORG, int remove_direct) {
    if (!LL || !xs || !xr) return set_error(DEISM_ERR_INVALID, "null geometry pointer");
    if (N_o < 0 || N_o > 250) return set_error(DEISM_ERR_INVALID, "maxReflOrder %d out of range", N_o);
    for (int i = 0; i < 3; ++i) { P.LL[i] = LL[i]; P.xs[i] = xs[i]; P.xr[i] = xr[i]; }
    P.max_d2 = max_d2;
    P.N_o = N_o;
    P.N_o_ORG = N_o_ORG > N_o ? N_o : N_o_ORG;  // cd:3360-3361
    P.remove_direct = remove_direct;
    return DEISM_OK;
}

}  // namespace deism

using namespace deism;

extern "C" int deism_shoebox_count_images(const double h_room_size[3], const double h_pos_source[3],
                                          const double h_pos_receiver[3], double max_dist_sq, int N_o,
                                          int N_o_ORG, int remove_direct, int64_t *d_early_counts,
                                          int64_t *d_late_counts, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    LatticeParams P;
    rc = fill_lattice(P, h_room_size, h_pos_source, h_pos_receiver, max_dist_sq, N_o, N_o_ORG,
                      remove_direct);
    if (rc) return rc;
    if (!d_early_counts || !d_late_counts) return set_error(DEISM_ERR_INVALID, "null count buffer");
    shoebox_count_kernel<<<8 * (N_o + 1), IMG_THREADS, 0, (cudaStream_t)stream>>>(P, d_early_counts,
                                                                                 d_late_counts);
    DEISM_LAUNCH_CHECK("shoebox_count_kernel");
    return DEISM_OK;
}

extern "C" int deism_shoebox_fill_images(const double h_room_size[3], const double h_pos_source[3],
                                         const double h_pos_receiver[3], double max_dist_sq, int N_o,
                                         int N_o_ORG, int remove_direct, const int64_t *d_early_counts,
                                         const int64_t *d_late_counts, int32_t *d_A_early,
                                         float *d_R_sI_r_early, float *d_R_s_rI_early,
                                         float *d_R_r_sI_early, int32_t *d_A_late, float *d_R_sI_r_late,
                                         float *d_R_s_rI_late, float *d_R_r_sI_late, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    LatticeParams P;
    rc = fill_lattice(P, h_room_size, h_pos_source, h_pos_receiver, max_dist_sq, N_o, N_o_ORG,
                      remove_direct);
    if (rc) return rc;
    if (!d_early_counts || !d_late_counts) return set_error(DEISM_ERR_INVALID, "null count buffer");
    prof_begin("shoebox_fill", (cudaStream_t)stream);
    shoebox_fill_kernel<<<8 * (N_o + 1), IMG_THREADS, 0, (cudaStream_t)stream>>>(
        P, d_early_counts, d_late_counts, d_A_early, d_R_sI_r_early, d_R_s_rI_early, d_R_r_sI_early,
        d_A_late, d_R_sI_r_late, d_R_s_rI_late, d_R_r_sI_late);
    prof_end("shoebox_fill", (cudaStream_t)stream);
    DEISM_LAUNCH_CHECK("shoebox_fill_kernel");
    return DEISM_OK;
}

extern "C" int deism_shoebox_attenuation(int64_t n_images, int64_t K, const deism_shoebox_atten *spec,
                                         void *d_out, int out_c128, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    if (!spec || !d_out) return set_error(DEISM_ERR_INVALID, "null spec/out");
    if (spec->mode != DEISM_ATTEN_FUSED_ROOM && spec->mode != DEISM_ATTEN_FUSED_ANGLES)
        return set_error(DEISM_ERR_INVALID, "deism_shoebox_attenuation needs a FUSED_* spec");
    if (!spec->d_A || !spec->d_Z_S) return set_error(DEISM_ERR_INVALID, "spec needs d_A and d_Z_S");
    if (spec->mode == DEISM_ATTEN_FUSED_ANGLES && !spec->d_R_sI_r)
        return set_error(DEISM_ERR_INVALID, "FUSED_ANGLES needs d_R_sI_r");
    if (n_images <= 0 || K <= 0) return DEISM_OK;
    if (n_images > 65535LL * 65535LL) return set_error(DEISM_ERR_INVALID, "too many images");
    RoomGeom room;
    for (int i = 0; i < 3; ++i) {
        room.LL[i] = spec->room_size[i]; room.xs[i] = spec->pos_source[i];
        room.xr[i] = spec->pos_receiver[i];
    }
    // grid.y is limited to 65535: loop over slabs of images
    const long slab = 65535;
    unsigned gx = (unsigned)((K + 255) / 256);
    if (gx > 64) gx = 64;
    for (long i0 = 0; i0 < n_images; i0 += slab) {
        long n = n_images - i0 < slab ? n_images - i0 : slab;
        dim3 grid(gx, (unsigned)n);
        const int32_t *A = spec->d_A + i0 * 6;
        const float *R = spec->d_R_sI_r ? spec->d_R_sI_r + i0 * 3 : nullptr;
        if (out_c128)
            shoebox_atten_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
                n, K, spec->mode, spec->ang_dep_flag, A, R, (const cplx *)spec->d_Z_S, room,
                (void *)((cplx *)d_out + i0 * K));
        else
            shoebox_atten_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
                n, K, spec->mode, spec->ang_dep_flag, A, R, (const cplx *)spec->d_Z_S, room,
                (void *)((float2 *)d_out + i0 * K));
        DEISM_LAUNCH_CHECK("shoebox_atten_kernel");
    }
    return DEISM_OK;
}
