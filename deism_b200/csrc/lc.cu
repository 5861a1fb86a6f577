// lc.cu — shoebox DEISM-LC accumulation with the wall reflection factors fused in.
//
// Replaces _numba_LC_matrix_batch (deism/parallel_backends.py:264-325), its batching driver
// (pb:594-654) and, in FUSED modes, the attenuation producers cd:2900-2928 / pb:140-186.
//
//   P[k] = sum_img  -atten[img,k] * 4pi/k * exp(-i k r_s)/(k r_s) * S[img,k] * R[img,k]
//   S[img,k] = sum_j i^{n_j} C_s[k,j] Y_{n_j}^{m_j}(theta_s,phi_s)       (source->receiver image)
//   R[img,k] = sum_j i^{v_j} C_r[k,j] Y_{v_j}^{u_j}(theta_r,phi_r)       (receiver->source image)
//
// S and R are dense contractions [I x J] . [J x K] whose [I x K] result must never reach memory
// (22 GB for the order-40 / 16k-frequency config), so they are computed tile by tile on the FP64
// tensor cores and consumed in registers.  In the real-harmonic basis (LcBasis below) Y is REAL,
// so each pass is two real GEMMs (Re and Im of the coefficients), issued as mma.sync m8n8k4 f64.
//
//   lc_prep_kernel   one thread per image: real harmonic rows y_q(img), r, 1/r, atten/r,
//                    rotation steps exp(-i dk r), direction cosines + wall exponents.
//   lc_coef_kernel   C[k,j] complex64 -> real-basis FP64 planes (Re | Im), i^n and (source pass)
//                    -4pi/k^2 folded in.
//   lc_main_kernel   <4,true>: CTA tile 64 images x 64 frequencies, 16 warps of 16x16; the
//                    frequency tile's coefficient planes are resident in shared memory, Y rows
//                    stream as whole-pass chunks through a TMA (cp.async.bulk -> UBLKCP) ring on
//                    mbarriers, refilled by the last warp out of a stage; <2,false>: 64 x 32,
//                    8 warps, C and Y both stream.  Operand rows are laid out in global memory as
//                    the kernel wants them in shared memory (row stride = 4 mod 16 doubles ->
//                    conflict-free fragment loads).  Per tile: fac = (atten/r) exp(-ikr) by one
//                    sincos per image and a rotation recurrence (computed in the shadow of the
//                    source pass's DMMAs), S' = fac*S parked thread-privately in shared memory,
//                    P[k] += S'*R accumulated in registers over the CTA's image chunk.
//   reduce_partials  deterministic sum of the per-chunk partials into out[K].
#include <cstring>
#include <mutex>
#include <vector>
#include <cmath>
#include <algorithm>

#include "common.cuh"

namespace deism {

constexpr int LC_TI = 64;        // images per tile
// most harmonics per staged chunk (multiple of 4): a whole order-5 pass when only Y streams
#ifndef LC_STREAM_JC
#define LC_STREAM_JC 12
#endif
#ifndef LC_STREAM_NST
#define LC_STREAM_NST 4
#endif
__host__ __device__ constexpr int lc_jcmax(bool cres_on) { return cres_on ? 36 : LC_STREAM_JC; }
#ifndef LC_NT
#define LC_NT 2                  // n-tiles (8 freqs each) per warp: warp tile = 16 images x 8*LC_NT freqs
#endif
constexpr int LC_YS = LC_TI + 4; // Y row stride (doubles): 68 = 4 mod 16 -> conflict-free fragments
constexpr int LC_MAXST = 16;     // most ring stages
constexpr int LC_MAXCHUNKS = DEISM_MAXCHUNKS; // most image chunks (grid.y) of a guided schedule
// A CTA is 4 (image) x WN (frequency) warps: tile = 64 images x TF = 16 WN frequencies.
//   WN = 4 (512 threads, 1 CTA/SM): the ftile's coefficient planes stay RESIDENT in shared memory
//          for the CTA's lifetime and only the Y chunks stream through a deep ring;
//   WN = 2 (256 threads, 2 CTAs/SM): coefficients and harmonics both stream (any number of
//          harmonics fits).
// C plane stride (doubles) = TF + 2, so a harmonic row (2 planes) is 4 mod 16 doubles: conflict-free.
constexpr int LC_WN_RES = 8 / LC_NT;     // warps along frequency: resident shape (64 frequencies)
constexpr int LC_WN_STR = 4 / LC_NT;     //                            streaming shape (32 frequencies)
__host__ __device__ constexpr int lc_tf(int WN) { return 8 * LC_NT * WN; }
__host__ __device__ constexpr int lc_cs(int WN) { return 8 * LC_NT * WN + 2; }

struct LcImage {        // 96 bytes per image (a tile of 64 records is one 6 KB bulk copy)
    double r, inv_r;
    double cx, cy, cz;
    unsigned char e[8];
    double2 a_flat;      // frequency-flat impedance: c64-rounded (mode ROOM) attenuation, times 1/r
    double2 rot;         // exp(-i dk r): one frequency step on a uniform wave-number grid
    double2 rot8;        // exp(-i 8 dk r)
};
static_assert(sizeof(LcImage) == 96, "LcImage must stay a multiple of 16 bytes");

struct LcPrepArgs {
    long n_images, n_pad, K;
    int Js, Jr, Js_pad, Jr_pad;
    const float *R_s_rI, *R_r_sI;
    int mode, angdep;
    const int32_t *A;
    const float *R_sI_r;
    const cplx *Z;
    RoomGeom room;
    const int *flags;    // flags[1]: uniform k grid, *(double *)(flags + 2) = dk
};

// Real-harmonic basis of one pass (source: index 0, receiver: index 1), in constant memory.
//
// Y_n^{-m} = (-1)^m conj(Y_n^m), so with Y_n^{|m|} = a + i b the sum over the caller's (n_j, m_j)
//     sum_j i^{n_j} C_j Y_{n_j}^{m_j}  =  sum_q C'_q y_q
// runs over REAL functions y_q in {a, b} of every distinct (n, |m|) (b only for |m| > 0), with
//     m = 0 :  C'_a += i^n C_j
//     m > 0 :  C'_a += i^n C_j            C'_b += i^{n+1} C_j
//     m < 0 :  C'_a += i^{n+2|m|} C_j     C'_b += i^{n+2|m|+3} C_j
// All weights are powers of i: forming C' costs one exact sign/swap and at most one rounding per
// entry, and the complex [I x J].[J x K] contraction becomes a REAL-by-complex one: two real
// GEMMs instead of the three of the 3-multiplication complex form.  A complete order-N set
// (m = -n..n) keeps its size, J' = J.
constexpr int LC_JMAX = 256;            // caller's harmonics per pass (order <= 15)
constexpr int LC_QMAX = 2 * LC_JMAX;    // real functions (J' <= 2 J, = J for complete sets)
constexpr int LC_MCOLS = 66;            // distinct |m| (orders up to 64 are accepted)
struct LcBasis {
    int Jq;
    unsigned char qn[LC_QMAX], qm[LC_QMAX], qim[LC_QMAX];   // y_q = (qim ? Im : Re) Y_qn^qm
    short cptr[LC_QMAX + 1];                                // contributions of q: cptr[q]..cptr[q+1]
    short cj[2 * LC_JMAX];                                  //   caller's column j
    unsigned char cp[2 * LC_JMAX];                          //   weight i^cp
    // evaluation program of the per-image harmonic rows: the distinct (n, |m|) of the basis grouped
    // into columns of equal |m| (ascending), ascending n inside a column — one sincos(|m| phi) and
    // one upward Legendre pass per column instead of one of each per (n, |m|)
    int ncol;
    unsigned char col_m[LC_MCOLS];
    short col_ptr[LC_MCOLS + 1];
    unsigned char ent_n[LC_JMAX];
    short ent_q[LC_JMAX];                                   // q of the a function (b = q + 1 when |m| > 0)
    double ent_norm[LC_JMAX];                               // sqrt((2n+1)/(4 pi) (n-|m|)!/(n+|m|)!), host-made
};
__constant__ LcBasis c_lc_basis[2];

// flags[0] = 1 iff Z_S[w,k] == Z_S[w,0] for all w,k (frequency-flat impedance)
__global__ void z_flat_kernel(const cplx *__restrict__ Z, long K, int *flags) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long n = 6 * K;
    bool diff = false;
    for (; i < n; i += (long)gridDim.x * blockDim.x) {
        long w = i / K;
        cplx a = Z[i], b = Z[w * K];
        if (a.x != b.x || a.y != b.y) diff = true;
    }
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicExch(flags, 0);
}

// flags[1] = 1 iff k is an arithmetic progression closely enough for the rotation recurrence of
// the factor phase: every step k[i+1]-k[i] within 1e-12 rad/m (and 1e-6 relative) of the mean step
// dk, so that the first-order correction for the residual d = step - dk leaves (d r)^2/2 < 1e-16
// for r up to 1e4 m.  np.arange grids deviate by a few ulp(k) ~ 1e-13.  *(double *)(flags+2) = dk.
__global__ void __launch_bounds__(256) k_uniform_kernel(const double *__restrict__ k, long K, int *flags) {
    const double dk = K > 1 ? (k[K - 1] - k[0]) / (double)(K - 1) : 0.0;
    bool bad = K < 2 || !(dk > 0.0);
    for (long i = threadIdx.x; i + 1 < K; i += blockDim.x) {
        double d = (k[i + 1] - k[i]) - dk;
        if (!(fabs(d) <= 1e-12 && fabs(d) <= 1e-6 * dk)) bad = true;
    }
    bool any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        flags[1] = any_bad ? 0 : 1;
        *reinterpret_cast<double *>(flags + 2) = dk;
    }
}

// Y planes: [tile][q < J_pad][LC_YS] doubles, the real functions y_q(image); image index within
// the tile fastest.  Rows >= n_images, functions >= J' and the 4 pad doubles are zero, so the main
// kernel's bulk copies never need a bounds check.
__device__ __forceinline__ void fill_y(double *Y, long base, int pass, int J_pad, double phi, double ct,
                                       bool live) {
    const LcBasis &B = c_lc_basis[pass];
    if (!live) {
        for (int q = 0; q < J_pad; ++q) Y[base + (long)q * LC_YS] = 0.0;
        return;
    }
    for (int q = B.Jq; q < J_pad; ++q) Y[base + (long)q * LC_YS] = 0.0;
    // Column by column (|m| fixed): P_m^m, then P_n^m upwards — the recurrence, the operations and
    // their order are those of sph_harm_dev (pb:40-83), which restarts them for every (n, m); the
    // normalisation comes from the host-made table (same operations, IEEE division and square root).
    const double somx2 = sqrt(fma(-ct, ct, 1.0));
    for (int c = 0; c < B.ncol; ++c) {
        const int am = B.col_m[c];
        double pmm = 1.0, fact = 1.0;
        for (int i = 0; i < am; ++i) { pmm *= -fact * somx2; fact += 2.0; }
        double sn, cs;
        sincos_cw((double)am * phi, &sn, &cs);
        double p0 = pmm, p1 = ct * (double)(2 * am + 1) * pmm;     // P_am^am, P_{am+1}^am
        int n = am;
        for (int e = B.col_ptr[c]; e < B.col_ptr[c + 1]; ++e) {
            const int want = B.ent_n[e];
            double p_val;
            if (want == am) {
                p_val = p0;
            } else {
                if (n < am + 1) n = am + 1;
                for (; n < want; ++n) {
                    const double pll = (ct * (double)(2 * n + 1) * p1 - (double)(n + am) * p0) / (double)(n + 1 - am);
                    p0 = p1; p1 = pll;
                }
                p_val = p1;
            }
            const double np_ = B.ent_norm[e] * p_val;
            const int q = B.ent_q[e];
            Y[base + (long)q * LC_YS] = np_ * cs;
            if (am) Y[base + (long)(q + 1) * LC_YS] = np_ * sn;
        }
    }
}

__global__ void __launch_bounds__(128)
lc_prep_kernel(LcPrepArgs a, LcImage *img_out, double *Ys, double *Yr) {
    long img = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= a.n_pad) return;
    LcImage rec;
    rec.cx = rec.cy = rec.cz = 1.0;
    for (int i = 0; i < 8; ++i) rec.e[i] = 0;
    const long tile = img / LC_TI, it = img % LC_TI;
    const long ys_base = tile * (long)a.Js_pad * LC_YS + it;
    const long yr_base = tile * (long)a.Jr_pad * LC_YS + it;
    if (it < 4) {   // the pad columns of this tile's rows
        for (int j = 0; j < a.Js_pad; ++j) Ys[ys_base + (long)j * LC_YS + LC_TI] = 0.0;
        for (int j = 0; j < a.Jr_pad; ++j) Yr[yr_base + (long)j * LC_YS + LC_TI] = 0.0;
    }
    rec.rot = cmake(1.0, 0.0);
    rec.rot8 = cmake(1.0, 0.0);
    if (img >= a.n_images) {
        rec.r = 1.0; rec.inv_r = 0.0;           // weight 0: contributes exactly nothing
        rec.a_flat = cmake(0.0, 0.0);
        img_out[img] = rec;
        fill_y(Ys, ys_base, 0, a.Js_pad, 0.0, 1.0, false);
        fill_y(Yr, yr_base, 1, a.Jr_pad, 0.0, 1.0, false);
        return;
    }
    double phi_s = (double)a.R_s_rI[img * 3 + 0], th_s = (double)a.R_s_rI[img * 3 + 1];
    rec.r = (double)a.R_s_rI[img * 3 + 2];
    rec.inv_r = 1.0 / rec.r;
    {
        double sn, cs;
        sincos_cw(*reinterpret_cast<const double *>(a.flags + 2) * rec.r, &sn, &cs);
        rec.rot = cmake(cs, -sn);
        sincos_cw(8.0 * *reinterpret_cast<const double *>(a.flags + 2) * rec.r, &sn, &cs);
        rec.rot8 = cmake(cs, -sn);
    }
    double phi_r = (double)a.R_r_sI[img * 3 + 0], th_r = (double)a.R_r_sI[img * 3 + 1];
    double s, ct_s, ct_r;
    sincos_cw(th_s, &s, &ct_s);
    sincos_cw(th_r, &s, &ct_r);
    fill_y(Ys, ys_base, 0, a.Js_pad, phi_s, ct_s, true);
    fill_y(Yr, yr_base, 1, a.Jr_pad, phi_r, ct_r, true);
    AttenGeom g;
    g.cx = g.cy = g.cz = 1.0;
    for (int i = 0; i < 8; ++i) g.e[i] = 0;
    rec.a_flat = cmake(rec.inv_r, 0.0);
    if (a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES) {
        if (a.mode == DEISM_ATTEN_FUSED_ROOM)
            atten_geom_room(a.A + img * 6, a.room.LL, a.room.xs, a.room.xr, a.angdep, g);
        else
            atten_geom_angles(a.A + img * 6, a.R_sI_r + img * 3, a.angdep, g);
        cplx at = shoebox_atten(a.Z, a.K, g.cx, g.cy, g.cz, g.e);  // Z[:,0]
        if (a.mode == DEISM_ATTEN_FUSED_ROOM) at = round_c64(at);
        rec.a_flat = cmake(at.x * rec.inv_r, at.y * rec.inv_r);
    }
    rec.cx = g.cx; rec.cy = g.cy; rec.cz = g.cz;
    for (int i = 0; i < 8; ++i) rec.e[i] = g.e[i];
    img_out[img] = rec;
}

// C planes: complex64 [K][J] -> [K_pad/TF][q < J_pad][plane 0..1][TF + 2] doubles, the real-basis
// coefficients C'_q (plane = Re | Im, frequency within the tile fastest, zero padded)
__global__ void __launch_bounds__(256)
lc_coef_kernel(const float2 *__restrict__ in, const double *__restrict__ kvec, long K, int J,
               long n_ftiles, int J_pad, int pass, int TF, double *out) {
    const int LC_CS = TF + 2, LC_TF = TF;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = n_ftiles * J_pad * LC_CS;
    if (i >= total) return;
    const LcBasis &B = c_lc_basis[pass];
    int col = (int)(i % LC_CS);
    int q = (int)((i / LC_CS) % J_pad);
    long ft = i / ((long)LC_CS * J_pad);
    long k = ft * LC_TF + col;
    double re = 0.0, im = 0.0;
    if (col < LC_TF && q < B.Jq && k < K) {
        for (int c = B.cptr[q]; c < B.cptr[q + 1]; ++c) {
            float2 v = in[k * J + B.cj[c]];
            double x = (double)v.x, y = (double)v.y;
            switch (B.cp[c] & 3) {                 // times i^cp
            case 0: re += x; im += y; break;
            case 1: re -= y; im += x; break;
            case 2: re -= x; im -= y; break;
            default: re += y; im -= x; break;
            }
        }
        if (pass == 0) {     // the source planes carry the factor's -4pi/k^2 (pb:314-317)
            const double w = -4.0 * 3.14159265358979323846 / (kvec[k] * kvec[k]);
            re *= w; im *= w;
        }
    }
    double *p = out + (ft * J_pad + q) * 2 * LC_CS + col;
    p[0] = re;
    p[LC_CS] = im;
}

struct LcMainArgs {
    long n_images, n_pad, K;
    int Js_pad, Jr_pad;      // harmonic counts padded to the DMMA k extent (multiples of 4)
    int ncs, ncr, JCs, JCr;  // chunks per pass and harmonics per chunk (multiples of 4)
    int nst;                 // ring stages (<= LC_MAXST)
    const double *CTs, *CTr; // [K_pad/TF][J_pad][2][TF+2]
    const double *Ys, *Yr;   // [n_pad/64][J_pad][LC_YS]
    const LcImage *img;      // [n_pad]
    const double *k;
    int mode;                // DEISM_ATTEN_*
    const void *atten;       // modes C64/C128
    const cplx *Z;           // fused modes
    const int *flags;        // flags[0]: impedance is frequency-flat
    long tiles_per_chunk;    // image tiles (of LC_TI) per CTA chunk: uniform chunks ...
    int n_guided;            // ... or, when > 0, that many chunks of DEcreasing size:
    int chunk_lo[LC_MAXCHUNKS + 1];   // image tiles [chunk_lo[c], chunk_lo[c+1]) of chunk c
    cplx *partial;           // [n_chunks, K]
};

// Shared-memory carve-up of lc_main_kernel<WN, CRES> (host and device use the same arithmetic).
struct LcCarve {
    unsigned cres, ring, stage_bytes, facp, rec, Z, kf, dkd, dkd8, bars, total;
};
__host__ __device__ inline LcCarve lc_carve(int WN, bool cres_on, int rows, int nst, int jc_rows) {
    const unsigned TF = lc_tf(WN), CS = lc_cs(WN), threads = 128 * WN;
    LcCarve c;
    unsigned o = 0;
    c.cres = o;  o += cres_on ? rows * 2 * CS * 8 : 0;                      // resident C planes
    c.stage_bytes = jc_rows * (LC_YS + (cres_on ? 0 : 2 * CS)) * 8;          // Y chunk (+ C chunk)
    c.ring = o;  o += nst * c.stage_bytes;
    c.facp = o;  o += 4 * LC_NT * threads * 16;     // thread-private factor, then S' = factor * S
    c.rec = o;   o += 2 * LC_TI * (unsigned)sizeof(LcImage);                // records: this / next tile
    c.Z = o;     o += 6 * TF * 16;
    c.kf = o;    o += TF * 8;
    c.dkd = o;   o += TF * 8;      // ulp-level deviations of 1 and 8 steps from the uniform grid
    c.dkd8 = o;  o += TF * 8;
    c.bars = o;  o += (LC_MAXST + 3) * 8 + (LC_MAXST + 2) * 4;
    c.total = (o + 15) & ~15u;
    return c;
}

// Factors (atten/r) exp(-ikr) of ONE image at the four frequencies a thread owns, for an impedance
// that varies with frequency: the wall exponents are shared by the four, so their attenuations are
// built together (shoebox_atten_n).  Only the general instantiation of the main kernel carries it.
#ifndef LC_REAL_CHAIN
#define LC_REAL_CHAIN 1     // 0: the general (complex-capable) wall-factor chain everywhere (A/B builds)
#endif
__device__ __forceinline__ void lc_general_factors(const cplx *Zs, int TF, const double *kf, const double *dkd,
                                                const double *dkd8, const LcImage *recp, int fqa,
                                                bool round_to_c64, bool uniform_k, bool real_tile,
                                                bool pairs_tile, double2 *fac, int stride) {
    constexpr int NQ = 2 * LC_NT;
    static_assert(LC_NT == 2, "frequency offsets below assume two n-tiles per warp");
    const LcImage &rec = *recp;
    const int fqs[NQ] = {fqa, fqa + 1, fqa + 8, fqa + 9};
    // E = exp(-i k r) / r first: independent of the wall factors, so its sincos and rotation steps
    // fill the latency of the reciprocal and power chains below
    cplx Eq[NQ];
    auto rot_step = [](cplx F, cplx rot, double dr) {
        F = cmul(F, rot);
        return cmake(fma(F.y, dr, F.x), fma(-F.x, dr, F.y));
    };
    if (uniform_k) {
        double sn, cs;
        sincos_cw(kf[fqa] * rec.r, &sn, &cs);
        Eq[0] = cmake(cs * rec.inv_r, -sn * rec.inv_r);
        Eq[1] = rot_step(Eq[0], rec.rot, dkd[fqa] * rec.r);
        Eq[2] = rot_step(Eq[0], rec.rot8, dkd8[fqa] * rec.r);
        Eq[3] = rot_step(Eq[2], rec.rot, dkd[fqa + 8] * rec.r);
    } else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double sn, cs;
            sincos_cw(kf[fqs[q]] * rec.r, &sn, &cs);
            Eq[q] = cmake(cs * rec.inv_r, -sn * rec.inv_r);
        }
    }
    if (real_tile) {
        double att_r[NQ];
        if (round_to_c64) {
            if (pairs_tile) shoebox_atten_real_n<NQ, 1, true>(Zs, TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, att_r);
            else shoebox_atten_real_n<NQ, 1, false>(Zs, TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, att_r);
#pragma unroll
            for (int q = 0; q < NQ; ++q) att_r[q] = (double)__double2float_rn(att_r[q]);
        } else {
            if (pairs_tile) shoebox_atten_real_n<NQ, 2, true>(Zs, TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, att_r);
            else shoebox_atten_real_n<NQ, 2, false>(Zs, TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, att_r);
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) fac[(long)q * stride] = cmake(att_r[q] * Eq[q].x, att_r[q] * Eq[q].y);
        return;
    }
    cplx att[NQ];
    shoebox_atten_n<NQ>(Zs, TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, att);
#pragma unroll
    for (int q = 0; q < NQ; ++q) fac[(long)q * stride] = cmul(round_to_c64 ? round_c64(att[q]) : att[q], Eq[q]);
}

// The main loop has NO CTA-wide barrier and no waiting producer: Y (and, when they stream, C)
// chunks and the per-image records arrive by TMA on "full" mbarriers; every warp counts itself out
// of a buffer when it is done with it and the LAST warp out refills it (issues the next bulk
// copies) on the spot.  The factor of a pair is computed, parked and consumed by the one thread
// that owns the pair in the DMMA accumulator layout.  Warps of a CTA drift by up to nst chunks.
template <int WN, bool CRES, bool FAST>
__global__ void __launch_bounds__(128 * WN, CRES ? 1 : 2)
lc_main_kernel(LcMainArgs a) {
    constexpr int NT = LC_NT;
    constexpr int TF = lc_tf(WN), CS = lc_cs(WN);
    constexpr int THREADS = 128 * WN, NWARPS = 4 * WN;
    extern __shared__ __align__(128) unsigned char lc_smem_raw[];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;          // DMMA fragment coordinates
    const int wm = warp / WN, wn = warp % WN;        // warp tile: images wm*16.., freqs wn*16..
    const long f0 = (long)blockIdx.x * TF;
    const long chunk = blockIdx.y;
    const long n_img_tiles = a.n_pad / LC_TI;
    const long tile_lo = a.n_guided ? (long)a.chunk_lo[chunk] : chunk * a.tiles_per_chunk;
    long tile_hi = a.n_guided ? (long)a.chunk_lo[chunk + 1] : tile_lo + a.tiles_per_chunk;
    if (tile_hi > n_img_tiles) tile_hi = n_img_tiles;
    const int n_tiles = tile_hi > tile_lo ? (int)(tile_hi - tile_lo) : 0;
    const int npt = a.ncs + a.ncr;                         // chunks per tile
    const int total = n_tiles * npt;                       // < 2^31: host caps the chunk sizes
    const int nst = a.nst;

    const int jc_rows = a.JCs > a.JCr ? a.JCs : a.JCr;    // harmonics a ring stage holds
    const LcCarve cv = lc_carve(WN, CRES, a.Js_pad + a.Jr_pad, nst, jc_rows);
    double *Cres = reinterpret_cast<double *>(lc_smem_raw + cv.cres);
    unsigned char *ring = lc_smem_raw + cv.ring;
    double2 (*facp)[THREADS] = reinterpret_cast<double2 (*)[THREADS]>(lc_smem_raw + cv.facp);
    LcImage (*recs)[LC_TI] = reinterpret_cast<LcImage (*)[LC_TI]>(lc_smem_raw + cv.rec);
    double2 (*Zs)[TF] = reinterpret_cast<double2 (*)[TF]>(lc_smem_raw + cv.Z);
    double *kf = reinterpret_cast<double *>(lc_smem_raw + cv.kf);
    double *dkd = reinterpret_cast<double *>(lc_smem_raw + cv.dkd);
    double *dkd8 = reinterpret_cast<double *>(lc_smem_raw + cv.dkd8);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(lc_smem_raw + cv.bars);
    unsigned long long *recfull = full + LC_MAXST;         // [2]
    unsigned long long *cfull = recfull + 2;               // resident C planes landed
    int *done = reinterpret_cast<int *>(cfull + 1);        // warps that have released a stage
    int *recdone = done + LC_MAXST;                        // ... a record buffer

    const bool fused = a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES;
    const bool flat = fused && a.flags[0] != 0;
    const bool uniform_k = a.flags[1] != 0;
    // Two instantiations are launched back to back: FAST carries only the frequency-flat /
    // uniform-grid factor code, the other one only the general code.  Which case applies is decided
    // on the device (flags), so the instantiation it does not belong to returns at once — the hot
    // path keeps its registers to itself and the host never has to read the flags.
    if ((flat && uniform_k) != FAST) return;

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
        for (int s = 0; s < 2; ++s) { mbar_init(&recfull[s], 1); recdone[s] = 0; }
        mbar_init(cfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const double dk = *reinterpret_cast<const double *>(a.flags + 2);
        for (int q = tid; q < TF; q += THREADS) {
            long f = f0 + q;
            double kk = f < a.K ? a.k[f] : 1.0;
            kf[q] = kk;
            dkd[q] = f + 1 < a.K ? (a.k[f + 1] - kk) - dk : 0.0;
            dkd8[q] = f + 8 < a.K ? (a.k[f + 8] - kk) - 8.0 * dk : 0.0;
        }
    }
    // Per-term attenuation with every impedance of the frequency tile real (what the material
    // conversions return): the real chain of common.cuh (shoebox_atten_real_n), decided per CTA.
    // ... and whether the two walls of every axis carry the same impedance throughout the tile (an
    // axis is then one base with the summed exponent).
    bool z_real_here = true, pairs_here = true;
    if (fused && !flat) {
        for (int i = tid; i < 3 * TF; i += THREADS) {
            int w = i / TF, q = i % TF;
            long f = f0 + q;
            const cplx z1 = f < a.K ? a.Z[(long)(2 * w) * a.K + f] : cmake(1.0, 0.0);
            const cplx z2 = f < a.K ? a.Z[(long)(2 * w + 1) * a.K + f] : cmake(1.0, 0.0);
            Zs[2 * w][q] = z1;
            Zs[2 * w + 1][q] = z2;
            z_real_here = z_real_here && z_is_real(z1) && z_is_real(z2);
            pairs_here = pairs_here && z1.x == z2.x;
        }
    }
    const bool real_tile = __syncthreads_and(z_real_here) && fused && !flat && LC_REAL_CHAIN;
    const bool pairs_tile = __syncthreads_and(pairs_here) && real_tile;

    // producer: chunk = JC harmonics of one pass for one image tile -> one bulk copy (Y), plus the
    // matching coefficient rows when they are not resident
    auto issue = [&](int tile_ord, int local, int stage) {
        const long tile = tile_lo + tile_ord;
        const bool src_pass = local < a.ncs;
        const int JCfull = src_pass ? a.JCs : a.JCr;
        const int j0 = (src_pass ? local : local - a.ncs) * JCfull;
        const int Jp = src_pass ? a.Js_pad : a.Jr_pad;
        const int JC = Jp - j0 < JCfull ? Jp - j0 : JCfull;       // the last chunk of a pass may be short
        const double *Yg = (src_pass ? a.Ys : a.Yr) + (tile * Jp + j0) * LC_YS;
        unsigned char *S = ring + (unsigned)stage * cv.stage_bytes;
        const unsigned yb = (unsigned)(JC * LC_YS * 8);
        if (CRES) {
            mbar_expect_tx(&full[stage], yb);
        } else {
            const double *CT = (src_pass ? a.CTs : a.CTr) + ((long)blockIdx.x * Jp + j0) * 2 * CS;
            const unsigned cb = (unsigned)(JC * 2 * CS * 8);
            mbar_expect_tx(&full[stage], cb + yb);
            bulk_g2s(S + jc_rows * LC_YS * 8, CT, cb, &full[stage]);
        }
        bulk_g2s(S, Yg, yb, &full[stage]);
    };
    auto issue_records = [&](int tt) {       // tile ordinal tt within this CTA
        const int rb = tt & 1;
        mbar_expect_tx(&recfull[rb], (unsigned)(LC_TI * sizeof(LcImage)));
        bulk_g2s(&recs[rb][0], a.img + (tile_lo + tt) * LC_TI, (unsigned)(LC_TI * sizeof(LcImage)),
                 &recfull[rb]);
    };

    // accumulators of the two real GEMMs (Re, Im of the pass): [plane][m-tile][n-tile][2]
    double T[2][2][NT][2];
    // P[freq]: this thread's 2*NT frequencies wn*16 + 8*nt + 2*t4 + e, summed over its images
    cplx P[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { P[nt][0] = cmake(0.0, 0.0); P[nt][1] = cmake(0.0, 0.0); }

    // chunk ordinal c = tile * npt + local (source chunks first); (pt, pl) = the chunk nst ahead of
    // c, the one whose operands replace chunk c's when the last warp leaves its stage
    int pt = nst / npt, pl = nst % npt;
    if (tid == 0 && total > 0) {
        if (CRES) {     // the ftile's coefficient planes: two contiguous blocks, loaded once
            const unsigned sb = (unsigned)(a.Js_pad * 2 * CS * 8), rbts = (unsigned)(a.Jr_pad * 2 * CS * 8);
            mbar_expect_tx(cfull, sb + rbts);
            bulk_g2s(Cres, a.CTs + (long)blockIdx.x * a.Js_pad * 2 * CS, sb, cfull);
            bulk_g2s(Cres + a.Js_pad * 2 * CS, a.CTr + (long)blockIdx.x * a.Jr_pad * 2 * CS, rbts, cfull);
        }
        issue_records(0);
        if (n_tiles > 1) issue_records(1);
        int l0 = 0, t0 = 0;
        for (int s = 0; s < nst && s < total; ++s) {
            issue(t0, l0, s);
            if (++l0 == npt) { l0 = 0; ++t0; }
        }
    }
    if (CRES && total > 0) mbar_wait(cfull, 0);

    const int fqa = wn * NT * 8 + 2 * t4;     // this thread's first frequency within the tile
    int c = 0, stage = 0;
    unsigned phase = 0;
    // one operand chunk: wait for its bytes, JC/4 k-steps of DMMA, count the warp out of the stage.
    // crow = first coefficient row of the chunk within the resident planes.
    // `side` is independent scalar work (the factor of the next pairs) placed in the same basic
    // block as the unrolled k-steps, so that it issues in the shadow of the DMMAs' pipe occupancy.
    // `pre` (the previous pass's epilogue, which ends by clearing T) runs after the wait so that it,
    // too, shares a basic block with the first k-step.
    auto run_chunk = [&](const int JC, const int crow, auto &&pre, auto &&side0, auto &&side1) {
        mbar_wait(&full[stage], phase);
        const double *Ystage = reinterpret_cast<const double *>(ring + (unsigned)stage * cv.stage_bytes);
        const double *Yw = Ystage + t4 * LC_YS + wm * 16 + g;
        const double *Cw = (CRES ? Cres + (crow + t4) * 2 * CS
                                 : Ystage + jc_rows * LC_YS + t4 * 2 * CS) + wn * NT * 8 + g;
        // a k-step = 4 harmonics: 2 + 2*NT fragment loads, 4*NT DMMAs
        auto load_frag = [&](int ks, double (&af)[2], double (&bf)[2][NT]) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) af[mt] = Yw[ks * LC_YS + mt * 8];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bf[p][nt] = Cw[(ks * 2 + p) * CS + nt * 8];
        };
        auto mma_frag = [&](const double (&af)[2], const double (&bf)[2][NT]) {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        dmma884(T[p][mt][nt][0], T[p][mt][nt][1], af[mt], bf[p][nt]);
        };
        double af[2], bf[2][NT];
        pre();
        load_frag(0, af, bf);
        mma_frag(af, bf);
        side0();
        if (JC > 4) {
            load_frag(4, af, bf);
            mma_frag(af, bf);
            side1();
            if (JC > 8) {
                // remaining k-steps, two per trip with ping-pong fragment registers: the loads of one
                // step are in flight while the other multiplies
                double a2[2], b2[2][NT];
                load_frag(8, af, bf);
                if (JC == 36) {          // a whole order-5 pass: straight-line code
#pragma unroll
                    for (int kq = 12; kq + 4 < 36; kq += 8) {
                        load_frag(kq, a2, b2);
                        mma_frag(af, bf);
                        load_frag(kq + 4, af, bf);
                        mma_frag(a2, b2);
                    }
                    mma_frag(af, bf);
                } else {
                int ks = 12;
#pragma unroll 1
                for (; ks + 4 < JC; ks += 8) {
                    load_frag(ks, a2, b2);
                    mma_frag(af, bf);
                    load_frag(ks + 4, af, bf);
                    mma_frag(a2, b2);
                }
                if (ks < JC) {
                    load_frag(ks, a2, b2);
                    mma_frag(af, bf);
                    mma_frag(a2, b2);
                } else {
                    mma_frag(af, bf);
                }
                }
            }
        } else {
            side1();
        }
        __syncwarp();
        if (lane == 0 && atomicAdd(&done[stage], 1) == NWARPS - 1) {
            // last warp out of this stage: refill it with chunk c + nst
            done[stage] = 0;
            if (c + nst < total) issue(pt, pl, stage);
        }
        ++c;
        if (++stage == nst) { stage = 0; phase ^= 1u; }
        if (++pl == npt) { pl = 0; ++pt; }
    };
    auto clear_T = [&]() {
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) { T[p][mt][nt][0] = 0.0; T[p][mt][nt][1] = 0.0; }
    };
    // one frequency step of the factor: exp(-i (dk + d) r) = rot * (1 - i d r) to first order in d
    auto rot_step = [](cplx F, cplx rot, double dr) {
        F = cmul(F, rot);
        return cmake(fma(F.y, dr, F.x), fma(-F.x, dr, F.y));
    };

    auto no_side = [] {};
    auto chunk_rows = [](int Jpad, int JC, int l) { return Jpad - l * JC < JC ? Jpad - l * JC : JC; };
    // end of a source pass: S = T0 + i T1;  park S' = fac * S
    auto epi_S = [&] {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = (mt * NT + nt) * 2 + e;
                    facp[j][tid] = cmul(facp[j][tid], cmake(T[0][mt][nt][e], T[1][mt][nt][e]));
                }
        clear_T();
    };
    // end of a receiver pass: P += S' * R  (a no-op before the first tile: T = 0, facp = 0)
    auto epi_P = [&] {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    cfma(P[nt][e], facp[(mt * NT + nt) * 2 + e][tid], cmake(T[0][mt][nt][e], T[1][mt][nt][e]));
        clear_T();
    };
#pragma unroll
    for (int j = 0; j < 4 * NT; ++j) facp[j][tid] = cmake(0.0, 0.0);
    clear_T();
    // fast path: frequency-flat impedance on a uniform wave-number grid.  One sincos per image
    // starts a rotation recurrence over the thread's frequencies; branch-free, so it can ride along
    // with the first source chunks of the tile.
    constexpr bool fast = FAST;

#pragma unroll 1
    for (int tt = 0; tt < n_tiles; ++tt) {
        const int rb = tt & 1;
        mbar_wait(&recfull[rb], (unsigned)((tt >> 1) & 1));
        // ---- factor fac = -atten * 4pi/k * exp(-ikr)/(k r) (pb:314-317) of the pairs this thread
        // owns (images wm*16 + mt*8 + g, frequencies fqa + 8*nt + e).  -4pi/k^2 is folded into the
        // source coefficient planes, so here fac = (atten / r) * exp(-i k r). ----
        auto factor_fast = [&](const int mt) {
            const LcImage &rec = recs[rb][wm * 16 + mt * 8 + g];
            const double r = rec.r;
            double sn, cs;
            sincos_cw(kf[fqa] * r, &sn, &cs);
            cplx F = cmul(rec.a_flat, cmake(cs, -sn));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int fq = fqa + 8 * nt;
                facp[(mt * NT + nt) * 2][tid] = F;
                facp[(mt * NT + nt) * 2 + 1][tid] = rot_step(F, rec.rot, dkd[fq] * r);
                if (nt + 1 < NT) F = rot_step(F, rec.rot8, dkd8[fq] * r);
            }
        };
        if (fast) {
            run_chunk(chunk_rows(a.Js_pad, a.JCs, 0), 0, epi_P, [&] { factor_fast(0); }, [&] { factor_fast(1); });
#pragma unroll 1
            for (int l = 1; l < a.ncs; ++l)
                run_chunk(chunk_rows(a.Js_pad, a.JCs, l), l * a.JCs, no_side, no_side, no_side);
        } else {
            // general path: per-term sincos off the uniform grid, per-term attenuation for a
            // frequency-dependent impedance or a materialised atten[I,K] (rolled: rarely hot)
            epi_P();
            const long i0 = (tile_lo + tt) * LC_TI;
            if (fused && !flat) {
#pragma unroll 1
                for (int mt = 0; mt < 2; ++mt)
                    lc_general_factors(&Zs[0][0], TF, kf, dkd, dkd8, &recs[rb][wm * 16 + mt * 8 + g], fqa,
                                       a.mode == DEISM_ATTEN_FUSED_ROOM, uniform_k, real_tile, pairs_tile,
                                       &facp[mt * 2 * NT][tid], THREADS);
            } else {
            cplx E = cmake(1.0, 0.0), E0 = E;
            static_assert(NT == 1 || NT == 2, "rotation order below assumes one or two n-tiles per warp");
#pragma unroll 1
            for (int j = 0; j < 4 * NT; ++j) {
                const int mt = j / (2 * NT), nt = (j >> 1) % NT, e = j & 1;
                const int il = wm * 16 + mt * 8 + g, fq = fqa + 8 * nt + e;
                const LcImage &rec = recs[rb][il];
                cplx att;
                if (flat) {
                    att = rec.a_flat;
                } else {
                    if (fused) {
                        att = shoebox_atten_term(&Zs[0][fq], TF, rec.cx, rec.cy, rec.cz, rec.e);
                        if (a.mode == DEISM_ATTEN_FUSED_ROOM) att = round_c64(att);
                    } else {
                        const long img = i0 + il, f = f0 + fq;
                        att = cmake(0.0, 0.0);
                        if (img < a.n_images && f < a.K) {
                            if (a.mode == DEISM_ATTEN_C64) {
                                float2 v = ((const float2 *)a.atten)[img * a.K + f];
                                att = cmake((double)v.x, (double)v.y);
                            } else {
                                att = ((const cplx *)a.atten)[img * a.K + f];
                            }
                        }
                    }
                    att = cmake(att.x * rec.inv_r, att.y * rec.inv_r);
                }
                // exp(-i k r): per-term sincos, or on a uniform grid one sincos per image and the
                // rotation recurrence (base, +1, +8, +9 steps for nt = 0,0,1,1 / e = 0,1,0,1)
                const int jj = j % (2 * NT);
                if (!uniform_k || jj == 0) {
                    double sn, cs;
                    sincos_cw(kf[fq] * rec.r, &sn, &cs);
                    E = cmake(cs, -sn);
                    E0 = E;
                } else if (NT == 2 && jj == 2) {
                    E = rot_step(E0, rec.rot8, dkd8[fq - 8] * rec.r);
                } else {
                    E = rot_step(E, rec.rot, dkd[fq - 1] * rec.r);
                }
                facp[j][tid] = cmul(att, E);
            }
            }
#pragma unroll 1
            for (int l = 0; l < a.ncs; ++l)
                run_chunk(chunk_rows(a.Js_pad, a.JCs, l), l * a.JCs, no_side, no_side, no_side);
        }
        __syncwarp();
        if (lane == 0 && atomicAdd(&recdone[rb], 1) == NWARPS - 1) {
            // last warp out of this record buffer: refill it with the records of tile tt+2
            recdone[rb] = 0;
            if (tt + 2 < n_tiles) issue_records(tt + 2);
        }
        // ---- receiver pass; its first chunk carries the source pass's epilogue ----
        run_chunk(chunk_rows(a.Jr_pad, a.JCr, 0), a.Js_pad, epi_S, no_side, no_side);
#pragma unroll 1
        for (int l = 1; l < a.ncr; ++l)
            run_chunk(chunk_rows(a.Jr_pad, a.JCr, l), a.Js_pad + l * a.JCr, no_side, no_side, no_side);
    }
    epi_P();

    // ---- reduce over the 8 row groups x 4 image warps that share a frequency (aliases facp) ----
    __syncthreads();
    double2 (*red)[TF] = reinterpret_cast<double2 (*)[TF]>(lc_smem_raw + cv.facp);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) red[wm * 8 + g][(wn * NT + nt) * 8 + 2 * t4 + e] = P[nt][e];
    __syncthreads();
    if (tid < TF) {
        cplx s = cmake(0.0, 0.0);
#pragma unroll 8
        for (int y = 0; y < 32; ++y) s = cadd(s, red[y][tid]);
        long f = f0 + tid;
        if (f < a.K) a.partial[chunk * a.K + f] = s;
    }
}

// ---------------------------------------------------------------------------------------
// Monopole source and receiver (J = 1 each; BASELINE C1 and the monopole run of C2): S and R do
// not depend on the image, so
//     P[k] = -C_s[k] C_r[k] Y_00^2 4pi/k^2 * sum_img (atten/r) exp(-i k r)
// and the contraction disappears.  Thread = one image x 16 consecutive frequencies: one sincos,
// then 15 steps of the rotation recurrence; 64 images x 64 frequencies per CTA pass; partial sums
// over the CTA's images stay in registers and are reduced once at the end.
// ---------------------------------------------------------------------------------------
constexpr int LM_THREADS = 256;          // 64 image lanes x 4 frequency groups
constexpr int LM_FPT_FAST = 16;          // frequencies per thread: rotation recurrence (flat impedance, uniform grid)
constexpr int LM_FPT_GEN = 4;            //                         per-term attenuation / arbitrary grid

struct LcMonoArgs {
    long n_images, n_pad, K;
    const LcImage *img;
    const double *k;
    const float2 *Cs, *Cr;   // [K][1]
    int mode;
    const void *atten;
    const cplx *Z;
    const int *flags;
    long tiles_per_chunk;
    long partial_rows;       // rows of partial[] the reduction will read (>= gridDim.y)
    cplx *partial;
};

// Like the main kernel, two instantiations go out back to back and the one the device-side flags
// do not select returns at once.  FAST: 16 frequencies per thread, one sincos + 15 rotation steps.
// General: 4 frequencies per thread; the four share the image's wall exponents, so their
// attenuations are built together (shoebox_atten_n), and the grid may be non-uniform.
template <int FPT, bool FAST>
__global__ void __launch_bounds__(LM_THREADS, 2)
lc_mono_kernel(LcMonoArgs a) {
    constexpr int TF = 4 * FPT;
    __shared__ double kf[TF], dkd[TF];
    __shared__ double2 Zs[6][TF];
    __shared__ double2 red[LM_THREADS / 32][FPT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int il = tid & (LC_TI - 1), fg = tid / LC_TI;          // image lane, frequency group
    const long f0 = (long)blockIdx.x * TF;
    const long n_img_tiles = a.n_pad / LC_TI;
    long tile_lo = (long)blockIdx.y * a.tiles_per_chunk;
    long tile_hi = tile_lo + a.tiles_per_chunk;
    if (tile_hi > n_img_tiles) tile_hi = n_img_tiles;
    const bool fused = a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES;
    const bool flat = fused && a.flags[0] != 0;
    const bool uniform_k = a.flags[1] != 0;
    if ((flat && uniform_k) != FAST) return;
    {
        const double dk = *reinterpret_cast<const double *>(a.flags + 2);
        for (int q = tid; q < TF; q += LM_THREADS) {
            const long f = f0 + q;
            const double kk = f < a.K ? a.k[f] : 1.0;
            kf[q] = kk;
            dkd[q] = f + 1 < a.K ? (a.k[f + 1] - kk) - dk : 0.0;
        }
    }
    // Real impedances, equal on the two walls of every axis, throughout the frequency tile: the
    // three-base simultaneous power of common.cuh.  (With six bases it does not pay here — a warp
    // holds 32 different images, so every bit of every exponent slot is taken by some lane: measured
    // 0.61 ms against 0.55 ms for the per-wall loops at C2 size.)
    bool pairs_here = true;
    if (fused && !flat)
        for (int i = tid; i < 3 * TF; i += LM_THREADS) {
            const int w = i / TF, q = i % TF;
            const long f = f0 + q;
            const cplx z1 = f < a.K ? a.Z[(long)(2 * w) * a.K + f] : cmake(1.0, 0.0);
            const cplx z2 = f < a.K ? a.Z[(long)(2 * w + 1) * a.K + f] : cmake(1.0, 0.0);
            Zs[2 * w][q] = z1;
            Zs[2 * w + 1][q] = z2;
            pairs_here = pairs_here && z_is_real(z1) && z_is_real(z2) && z1.x == z2.x;
        }
    const bool pairs_tile = __syncthreads_and(pairs_here) && fused && !flat && LC_REAL_CHAIN;
    const int q0 = fg * FPT;
    cplx P[FPT];
#pragma unroll
    for (int q = 0; q < FPT; ++q) P[q] = cmake(0.0, 0.0);

#pragma unroll 1
    for (long tile = tile_lo; tile < tile_hi; ++tile) {
        const long img = tile * LC_TI + il;
        const LcImage rec = a.img[img];
        const double r = rec.r;
        if (FAST) {
            double sn, cs;
            sincos_cw(kf[q0] * r, &sn, &cs);
            cplx F = cmul(rec.a_flat, cmake(cs, -sn));
#pragma unroll
            for (int q = 0; q < FPT; ++q) {
                P[q] = cadd(P[q], F);
                if (q + 1 < FPT) {
                    F = cmul(F, rec.rot);                      // exp(-i (dk + d) r) ~ rot (1 - i d r)
                    const double dr = dkd[q0 + q] * r;
                    F = cmake(fma(F.y, dr, F.x), fma(-F.x, dr, F.y));
                }
            }
        } else {
            cplx att[FPT], E[FPT];
            if (flat) {
#pragma unroll
                for (int q = 0; q < FPT; ++q) att[q] = rec.a_flat;
            } else {
                if (fused) {
                    int fqs[FPT];
#pragma unroll
                    for (int q = 0; q < FPT; ++q) fqs[q] = q0 + q;
                    if (pairs_tile) {
                        double ar[FPT];
                        if (a.mode == DEISM_ATTEN_FUSED_ROOM) {
                            shoebox_atten_real_n<FPT, 1, true>(&Zs[0][0], TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, ar);
#pragma unroll
                            for (int q = 0; q < FPT; ++q) att[q] = cmake((double)__double2float_rn(ar[q]), 0.0);
                        } else {
                            shoebox_atten_real_n<FPT, 2, true>(&Zs[0][0], TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, ar);
#pragma unroll
                            for (int q = 0; q < FPT; ++q) att[q] = cmake(ar[q], 0.0);
                        }
                    } else {
                        shoebox_atten_n<FPT>(&Zs[0][0], TF, fqs, rec.cx, rec.cy, rec.cz, rec.e, att);
                        if (a.mode == DEISM_ATTEN_FUSED_ROOM) {
#pragma unroll
                            for (int q = 0; q < FPT; ++q) att[q] = round_c64(att[q]);
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < FPT; ++q) {
                        const long f = f0 + q0 + q;
                        att[q] = cmake(0.0, 0.0);
                        if (img < a.n_images && f < a.K) {
                            if (a.mode == DEISM_ATTEN_C64) {
                                float2 v = ((const float2 *)a.atten)[img * a.K + f];
                                att[q] = cmake((double)v.x, (double)v.y);
                            } else {
                                att[q] = ((const cplx *)a.atten)[img * a.K + f];
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < FPT; ++q) att[q] = cmake(att[q].x * rec.inv_r, att[q].y * rec.inv_r);
            }
            if (uniform_k) {
                double sn, cs;
                sincos_cw(kf[q0] * r, &sn, &cs);
                E[0] = cmake(cs, -sn);
#pragma unroll
                for (int q = 1; q < FPT; ++q) {
                    const cplx F = cmul(E[q - 1], rec.rot);
                    const double dr = dkd[q0 + q - 1] * r;
                    E[q] = cmake(fma(F.y, dr, F.x), fma(-F.x, dr, F.y));
                }
            } else {
#pragma unroll
                for (int q = 0; q < FPT; ++q) {
                    double sn, cs;
                    sincos_cw(kf[q0 + q] * r, &sn, &cs);
                    E[q] = cmake(cs, -sn);
                }
            }
#pragma unroll
            for (int q = 0; q < FPT; ++q) cfma(P[q], att[q], E[q]);
        }
    }
    // ---- sum over the 64 image lanes of each frequency group: shuffles, then two warps per group ----
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        double x = P[q].x, y = P[q].y;
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_down_sync(0xffffffffu, x, o);
            y += __shfl_down_sync(0xffffffffu, y, o);
        }
        if (lane == 0) red[warp][q] = cmake(x, y);
    }
    __syncthreads();
    if (tid < TF) {
        const int g = tid / FPT, q = tid % FPT;
        const cplx s = cadd(red[2 * g][q], red[2 * g + 1][q]);
        const long f = f0 + tid;
        if (f < a.K) {
            // S R (-4pi/k^2):  C_s Y_00 * C_r Y_00, Y_00 = 1/(2 sqrt(pi))
            const float2 cs = a.Cs[f], cr = a.Cr[f];
            const double y00 = sph_harm_dev(0, 0, 0.0, 1.0).x;
            const cplx SR = cmul(cmake((double)cs.x * y00, (double)cs.y * y00),
                                 cmake((double)cr.x * y00, (double)cr.y * y00));
            const double w = -4.0 * 3.14159265358979323846 / (kf[tid] * kf[tid]);
            a.partial[(long)blockIdx.y * a.K + f] = cmul(cmake(SR.x * w, SR.y * w), s);
            // this instantiation may run with fewer image chunks than partial[] has rows
            for (long row = blockIdx.y + gridDim.y; row < a.partial_rows; row += gridDim.y)
                a.partial[row * a.K + f] = cmake(0.0, 0.0);
        }
    }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const cplx *__restrict__ partial, long n_chunks, long K, cplx *out,
                       int accumulate) {
    long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= K) return;
    cplx s = accumulate ? out[f] : cmake(0.0, 0.0);
    for (long c = 0; c < n_chunks; ++c) s = cadd(s, partial[c * K + f]);
    out[f] = s;
}

int launch_reduce_partials(const cplx *partial, long n_chunks, long K, double *d_out, int accumulate,
                           cudaStream_t st) {
    reduce_partials_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(partial, n_chunks, K,
                                                                       (cplx *)d_out, accumulate);
    DEISM_LAUNCH_CHECK("reduce_partials_kernel");
    return DEISM_OK;
}

int validate_atten(const deism_shoebox_atten *at) {
    if (!at) return set_error(DEISM_ERR_INVALID, "null attenuation spec");
    switch (at->mode) {
    case DEISM_ATTEN_C64:
    case DEISM_ATTEN_C128:
        if (!at->d_atten) return set_error(DEISM_ERR_INVALID, "materialised mode needs d_atten");
        break;
    case DEISM_ATTEN_FUSED_ANGLES:
        if (!at->d_R_sI_r) return set_error(DEISM_ERR_INVALID, "FUSED_ANGLES needs d_R_sI_r");
        /* fallthrough */
    case DEISM_ATTEN_FUSED_ROOM:
        if (!at->d_A || !at->d_Z_S) return set_error(DEISM_ERR_INVALID, "fused mode needs d_A, d_Z_S");
        break;
    default:
        return set_error(DEISM_ERR_INVALID, "unknown attenuation mode %d", at->mode);
    }
    return DEISM_OK;
}

// J harmonics (padded to Jpad, a multiple of 4) -> nc chunks of JC rows (multiple of 4, <= jcmax);
// the last chunk holds what is left.  Fewest chunks, near-equal sizes.
static void pick_chunks(int J, int jcmax, int &nc, int &JC, int &Jpad) {
    Jpad = (J + 3) / 4 * 4;
    nc = (Jpad + jcmax - 1) / jcmax;
    JC = ((Jpad / 4 + nc - 1) / nc) * 4;
}

// Real-harmonic basis of the caller's (n_j, m_j) list (see LcBasis): distinct (n,|m|) in order of
// first appearance, function a then (|m| > 0) b; repeated (n,m) entries simply add up.
static void build_real_basis(int J, const int64_t *n, const int64_t *m, LcBasis &B) {
    int first[LC_JMAX], cnt[LC_QMAX];
    B.Jq = 0;
    for (int j = 0; j < J; ++j) {
        const int nn = (int)n[j], am = (int)llabs(m[j]);
        int q = -1;
        for (int t = 0; t < B.Jq; ++t)
            if (!B.qim[t] && B.qn[t] == nn && B.qm[t] == am) { q = t; break; }
        if (q < 0) {
            q = B.Jq;
            for (int part = 0; part < (am ? 2 : 1); ++part, ++B.Jq) {
                B.qn[B.Jq] = (unsigned char)nn; B.qm[B.Jq] = (unsigned char)am;
                B.qim[B.Jq] = (unsigned char)part; cnt[B.Jq] = 0;
            }
        }
        first[j] = q;
        cnt[q]++;
        if (am) cnt[q + 1]++;
    }
    B.cptr[0] = 0;
    for (int q = 0; q < B.Jq; ++q) { B.cptr[q + 1] = (short)(B.cptr[q] + cnt[q]); cnt[q] = 0; }
    for (int j = 0; j < J; ++j) {
        const int nn = (int)n[j], mm = (int)m[j], am = abs(mm), q = first[j];
        const int pa = mm >= 0 ? nn : nn + 2 * am;          // weight of a: i^pa
        int c = B.cptr[q] + cnt[q]++;
        B.cj[c] = (short)j; B.cp[c] = (unsigned char)(pa & 3);
        if (am) {
            c = B.cptr[q + 1] + cnt[q + 1]++;
            B.cj[c] = (short)j; B.cp[c] = (unsigned char)((pa + (mm > 0 ? 1 : 3)) & 3);
        }
    }
    // evaluation program: columns of equal |m|, ascending n
    B.ncol = 0;
    int ne = 0;
    for (int am = 0; am <= 64; ++am) {
        bool any = false;
        for (int n = am; n <= 64; ++n)
            for (int q = 0; q < B.Jq; ++q) {
                if (B.qim[q] || B.qm[q] != am || B.qn[q] != n) continue;
                if (!any) { B.col_m[B.ncol] = (unsigned char)am; B.col_ptr[B.ncol] = (short)ne; any = true; }
                double ratio = 1.0;
                for (int i = n - am + 1; i <= n + am; ++i) ratio *= (double)i;
                B.ent_n[ne] = (unsigned char)n;
                B.ent_q[ne] = (short)q;
                B.ent_norm[ne] = sqrt((double)(2 * n + 1) / (4.0 * 3.14159265358979323846) / ratio);
                ++ne;
            }
        if (any) ++B.ncol;
    }
    B.col_ptr[B.ncol] = (short)ne;
}

int launch_k_uniform(const double *d_k, long K, int *flags, cudaStream_t st) {
    k_uniform_kernel<<<1, 256, 0, st>>>(d_k, K, flags);
    DEISM_LAUNCH_CHECK("k_uniform_kernel");
    return DEISM_OK;
}

int launch_z_flat(const deism_shoebox_atten *at, long K, int *flags, cudaStream_t st) {
    DEISM_CUDA_TRY(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
    if (at->mode == DEISM_ATTEN_FUSED_ROOM || at->mode == DEISM_ATTEN_FUSED_ANGLES) {
        // flags[0] = 1 (little endian) without a host source: the call stays CUDA-graph capturable
        DEISM_CUDA_TRY(cudaMemsetAsync(flags, 1, 1, st));
        z_flat_kernel<<<32, 256, 0, st>>>((const cplx *)at->d_Z_S, K, flags);
        DEISM_LAUNCH_CHECK("z_flat_kernel");
    }
    return DEISM_OK;
}

}  // namespace deism

using namespace deism;

extern "C" int deism_lc_shoebox(int64_t n_images, int64_t K, int Js, int Jr, const int64_t *h_n_all,
                                const int64_t *h_m_all, const int64_t *h_v_all, const int64_t *h_u_all,
                                const float *d_C_s_vec, const float *d_C_r_vec, const float *d_R_s_rI,
                                const float *d_R_r_sI, const deism_shoebox_atten *atten,
                                const double *d_k, double *d_out, int accumulate, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 0) return DEISM_OK;
    if (!d_out || !d_k) return set_error(DEISM_ERR_INVALID, "null d_out/d_k");
    if (n_images <= 0) {   // pb:607-608: zero images -> zeros
        if (!accumulate) DEISM_CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)K * 16, st));
        return DEISM_OK;
    }
    if (Js <= 0 || Jr <= 0 || Js > LC_JMAX || Jr > LC_JMAX)
        return set_error(DEISM_ERR_INVALID, "Js/Jr out of range (1..%d)", LC_JMAX);
    if (!h_n_all || !h_m_all || !h_v_all || !h_u_all || !d_C_s_vec || !d_C_r_vec || !d_R_s_rI ||
        !d_R_r_sI)
        return set_error(DEISM_ERR_INVALID, "null input pointer");
    rc = validate_atten(atten);
    if (rc) return rc;

    // The basis tables live in constant memory.  They depend only on the caller's (n,m) lists, so
    // the upload (and the host synchronisation that protects the staging copy) is skipped when the
    // device already holds the same tables: back-to-back calls stay fully asynchronous.
    static std::mutex basis_mu;
    static LcBasis uploaded[2];
    static int uploaded_dev = -1;
    LcBasis hb[2];
    memset(hb, 0, sizeof(hb));
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t *hn = pass ? h_v_all : h_n_all, *hm = pass ? h_u_all : h_m_all;
        const int J = pass ? Jr : Js;
        for (int j = 0; j < J; ++j)
            if (hn[j] < 0 || hn[j] > 64 || llabs(hm[j]) > hn[j])
                return set_error(DEISM_ERR_INVALID, "bad (%s) = (%ld,%ld) at j=%d", pass ? "v,u" : "n,m",
                                 (long)hn[j], (long)hm[j], j);
        build_real_basis(J, hn, hm, hb[pass]);
    }
    {
        int dev = -1;
        DEISM_CUDA_TRY(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(basis_mu);
        if (dev != uploaded_dev || memcmp(hb, uploaded, sizeof(hb)) != 0) {
            uploaded_dev = -1;
            bump_generation();
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_lc_basis, hb, sizeof(hb), 0, cudaMemcpyHostToDevice, st));
            DEISM_CUDA_TRY(cudaStreamSynchronize(st));   // hb is on the stack
            memcpy(uploaded, hb, sizeof(hb));
            uploaded_dev = dev;
        }
    }
    const int Jqs = hb[0].Jq, Jqr = hb[1].Jq;
    // monopole source and receiver: dedicated kernel (no contraction)
    const bool monopole = Js == 1 && Jr == 1 && h_n_all[0] == 0 && h_m_all[0] == 0 && h_v_all[0] == 0 &&
                          h_u_all[0] == 0 && !getenv("DEISM_LC_NO_MONO");

    const long n_img_tiles = (n_images + LC_TI - 1) / LC_TI;
    const long n_pad = n_img_tiles * LC_TI;
    // kernel shape: resident coefficient planes (512 threads, 64 frequencies, 1 CTA/SM, Y chunks of
    // up to 36 harmonics) when they leave room for a ring of >= 2 such chunks, else the streaming
    // shape (256 threads, 32 frequencies, 2 CTAs/SM, 4-stage ring of C+Y chunks of <= 12 harmonics).
    // Harmonics go in ncs (ncr) chunks of JCs (JCr) each (multiples of 4, the DMMA k extent), sized
    // to waste as little zero padding as possible.
    const int smem_max = smem_optin_bytes();
    static const char *force = getenv("DEISM_LC_SHAPE");     // "stream": A/B runs
    bool resident = !(force && !strcmp(force, "stream"));
    int nst = 4, ncs, ncr, JCs, JCr, Js_pad, Jr_pad;
    if (resident) {
        pick_chunks(Jqs, lc_jcmax(true), ncs, JCs, Js_pad);
        pick_chunks(Jqr, lc_jcmax(true), ncr, JCr, Jr_pad);
        const LcCarve c0 = lc_carve(LC_WN_RES, true, Js_pad + Jr_pad, 0, JCs > JCr ? JCs : JCr);
        const int room = smem_max > (int)c0.total ? (smem_max - (int)c0.total) / (int)c0.stage_bytes : 0;
        if (room >= 2) nst = room < LC_MAXST ? room : LC_MAXST;
        else resident = false;
    }
    if (!resident) {
        pick_chunks(Jqs, lc_jcmax(false), ncs, JCs, Js_pad);
        pick_chunks(Jqr, lc_jcmax(false), ncr, JCr, Jr_pad);
        nst = LC_STREAM_NST;
    }
    const int WN = resident ? LC_WN_RES : LC_WN_STR, TF = lc_tf(WN), CS = lc_cs(WN);
    const int ctas_per_sm = resident ? 1 : 2;
    const LcCarve carve = lc_carve(WN, resident, Js_pad + Jr_pad, nst, JCs > JCr ? JCs : JCr);
    const long n_ftiles = (K + TF - 1) / TF;

    // image tiles are cut into n_chunks CTAs per ftile: ~`waves` waves of CTAs over the SMs, picked
    // so that the last wave is as full as possible
    // (a CTA's prologue — the resident coefficient planes, the first operand chunks — is ~4 us: a CTA
    // should live for >= ~48 tiles, so a frequency slab of a multi-GPU run takes fewer, longer waves)
    static const int waves_env = getenv("DEISM_LC_WAVES") ? atoi(getenv("DEISM_LC_WAVES")) : 0;
    const long slots = (long)sm_count() * ctas_per_sm;
    int waves = 16;
    if (waves_env > 0) {
        waves = waves_env;
    } else {
        const long per_slot = n_ftiles * n_img_tiles / slots;
        waves = (int)(per_slot / 48);
        if (waves < 4) waves = 4;
        if (waves > 16) waves = 16;
    }
    long n_chunks = 1, tiles_per_chunk = n_img_tiles;
    {
        const long min_chunks = n_img_tiles * (ncs + ncr) / (1L << 30) + 1;  // chunk counter is an int
        long lo = slots * waves / 2 / n_ftiles, hi = slots * waves * 3 / 2 / n_ftiles;
        if (lo < min_chunks) lo = min_chunks;
        if (hi < lo) hi = lo;
        if (hi > 65535) hi = 65535;
        if (lo > hi) lo = hi;
        double best = -1.0;
        for (long n = lo; n <= hi; ++n) {
            if (n > n_img_tiles) break;
            const long tpc = (n_img_tiles + n - 1) / n, nn = (n_img_tiles + tpc - 1) / tpc;
            const long ctas = nn * n_ftiles, rounds = (ctas + slots - 1) / slots;
            // time ~ rounds * (tpc + fixed per-CTA cost of ~1/2 tile); useful work = n_img_tiles * n_ftiles
            const double eff = (double)n_img_tiles * n_ftiles / ((double)rounds * slots * (tpc + 0.5));
            if (eff > best) { best = eff; n_chunks = nn; tiles_per_chunk = tpc; }
        }
        if (best < 0.0) {
            n_chunks = n_img_tiles < lo ? n_img_tiles : lo;
            if (n_chunks < 1) n_chunks = 1;
            tiles_per_chunk = (n_img_tiles + n_chunks - 1) / n_chunks;
            n_chunks = (n_img_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
        }
    }

    // Chunks of DEcreasing size ("guided" schedule: the block scheduler hands out CTAs in launch
    // order, so the last ones to start should be the shortest) when a simulation of that greedy
    // hand-out says they finish earlier than the uniform cut above; cached per problem shape.
    GuidedPlan gplan;
    // (the plan arrives with the second call of a shape; the partial sums are sized for it from the
    // first call on, so that call — possibly a graph capture — allocates nothing)
    const bool may_guide = !monopole && guided_chunks_possible(n_img_tiles, n_ftiles, n_chunks);
    const bool guided = may_guide && guided_chunks(n_img_tiles, n_ftiles, slots, n_chunks, tiles_per_chunk,
                                                   ((1L << 30) - 1) / (ncs + ncr), gplan);
    if (guided) n_chunks = gplan.n;
    const long partial_rows = may_guide && n_chunks < DEISM_MAXCHUNKS ? DEISM_MAXCHUNKS : n_chunks;

    LcImage *img = (LcImage *)workspace(WS_LC_IMG, (size_t)n_pad * sizeof(LcImage));
    double *Ys = (double *)workspace(WS_LC_YS, (size_t)n_img_tiles * Js_pad * LC_YS * sizeof(double));
    double *Yr = (double *)workspace(WS_LC_YR, (size_t)n_img_tiles * Jr_pad * LC_YS * sizeof(double));
    double *CT = (double *)workspace(WS_SMALL,
                                     (size_t)n_ftiles * (Js_pad + Jr_pad) * 2 * CS * sizeof(double));
    cplx *partial = (cplx *)workspace(WS_PARTIAL, (size_t)partial_rows * K * sizeof(cplx));
    int *flags = (int *)workspace(WS_FLAGS, 64);
    if (!img || !Ys || !Yr || !CT || !partial || !flags) return DEISM_ERR_NOMEM;
    double *CTs = CT, *CTr = CT + (size_t)n_ftiles * Js_pad * 2 * CS;

    rc = launch_z_flat(atten, K, flags, st);
    if (rc) return rc;
    k_uniform_kernel<<<1, 256, 0, st>>>(d_k, K, flags);
    DEISM_LAUNCH_CHECK("k_uniform_kernel");

    if (monopole) {
        LcPrepArgs mp;
        mp.flags = flags;
        mp.n_images = n_images; mp.n_pad = n_pad; mp.K = K; mp.Js = 1; mp.Jr = 1;
        mp.Js_pad = 0; mp.Jr_pad = 0;                       // records only, no harmonic rows
        mp.R_s_rI = d_R_s_rI; mp.R_r_sI = d_R_r_sI;
        mp.mode = atten->mode; mp.angdep = atten->ang_dep_flag;
        mp.A = atten->d_A; mp.R_sI_r = atten->d_R_sI_r; mp.Z = (const cplx *)atten->d_Z_S;
        for (int i = 0; i < 3; ++i) {
            mp.room.LL[i] = atten->room_size[i]; mp.room.xs[i] = atten->pos_source[i];
            mp.room.xr[i] = atten->pos_receiver[i];
        }
        lc_prep_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, st>>>(mp, img, Ys, Yr);
        DEISM_LAUNCH_CHECK("lc_prep_kernel");
        const long m_ftiles = (K + 4 * LM_FPT_FAST - 1) / (4 * LM_FPT_FAST);
        long m_chunks = (long)sm_count() * 8 * 4 / m_ftiles;          // ~4 waves of 8 CTAs/SM
        if (m_chunks < 1) m_chunks = 1;
        if (m_chunks > n_img_tiles) m_chunks = n_img_tiles;
        if (m_chunks > n_chunks) m_chunks = n_chunks;                 // partial[] was sized for n_chunks
        const long m_tpc = (n_img_tiles + m_chunks - 1) / m_chunks;
        m_chunks = (n_img_tiles + m_tpc - 1) / m_tpc;
        LcMonoArgs mo;
        mo.n_images = n_images; mo.n_pad = n_pad; mo.K = K; mo.img = img; mo.k = d_k;
        mo.Cs = (const float2 *)d_C_s_vec; mo.Cr = (const float2 *)d_C_r_vec;
        mo.mode = atten->mode; mo.atten = atten->d_atten; mo.Z = (const cplx *)atten->d_Z_S;
        mo.flags = flags; mo.tiles_per_chunk = m_tpc; mo.partial = partial; mo.partial_rows = m_chunks;
        // the general instantiation has 4x the frequency tiles: give it 1/4 of the image chunks so
        // that, when it does not apply, its empty grid costs no more than the fast one's
        const long g_ftiles = (K + 4 * LM_FPT_GEN - 1) / (4 * LM_FPT_GEN);
        long g_chunks = m_chunks / 4 > 0 ? m_chunks / 4 : 1;
        const long g_tpc = (n_img_tiles + g_chunks - 1) / g_chunks;
        g_chunks = (n_img_tiles + g_tpc - 1) / g_tpc;
        if (g_ftiles > 2147483647L) return set_error(DEISM_ERR_INVALID, "too many frequencies");
        prof_begin("lc_main", st);
        lc_mono_kernel<LM_FPT_FAST, true><<<dim3((unsigned)m_ftiles, (unsigned)m_chunks), LM_THREADS, 0, st>>>(mo);
        LcMonoArgs mg = mo;
        mg.tiles_per_chunk = g_tpc;
        lc_mono_kernel<LM_FPT_GEN, false><<<dim3((unsigned)g_ftiles, (unsigned)g_chunks), LM_THREADS, 0, st>>>(mg);
        prof_end("lc_main", st);
        count_launch();
        DEISM_LAUNCH_CHECK("lc_mono_kernel");
        return launch_reduce_partials(partial, m_chunks, K, d_out, accumulate, st);
    }

    LcPrepArgs pa;
    pa.flags = flags;
    pa.n_images = n_images; pa.n_pad = n_pad; pa.K = K; pa.Js = Js; pa.Jr = Jr;
    pa.Js_pad = Js_pad; pa.Jr_pad = Jr_pad;
    pa.R_s_rI = d_R_s_rI; pa.R_r_sI = d_R_r_sI;
    pa.mode = atten->mode; pa.angdep = atten->ang_dep_flag;
    pa.A = atten->d_A; pa.R_sI_r = atten->d_R_sI_r; pa.Z = (const cplx *)atten->d_Z_S;
    for (int i = 0; i < 3; ++i) {
        pa.room.LL[i] = atten->room_size[i]; pa.room.xs[i] = atten->pos_source[i];
        pa.room.xr[i] = atten->pos_receiver[i];
    }
    lc_prep_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, st>>>(pa, img, Ys, Yr);
    DEISM_LAUNCH_CHECK("lc_prep_kernel");
    lc_coef_kernel<<<(unsigned)((n_ftiles * Js_pad * CS + 255) / 256), 256, 0, st>>>(
        (const float2 *)d_C_s_vec, d_k, K, Js, n_ftiles, Js_pad, 0, TF, CTs);
    DEISM_LAUNCH_CHECK("lc_coef_kernel");
    lc_coef_kernel<<<(unsigned)((n_ftiles * Jr_pad * CS + 255) / 256), 256, 0, st>>>(
        (const float2 *)d_C_r_vec, d_k, K, Jr, n_ftiles, Jr_pad, 1, TF, CTr);
    DEISM_LAUNCH_CHECK("lc_coef_kernel");

    LcMainArgs ma;
    ma.n_images = n_images; ma.n_pad = n_pad; ma.K = K;
    ma.Js_pad = Js_pad; ma.Jr_pad = Jr_pad; ma.ncs = ncs; ma.ncr = ncr; ma.JCs = JCs; ma.JCr = JCr;
    ma.CTs = CTs; ma.CTr = CTr; ma.Ys = Ys; ma.Yr = Yr; ma.img = img; ma.k = d_k;
    ma.mode = atten->mode; ma.atten = atten->d_atten; ma.Z = (const cplx *)atten->d_Z_S;
    ma.flags = flags; ma.tiles_per_chunk = tiles_per_chunk; ma.partial = partial;
    ma.nst = nst;
    ma.n_guided = 0;
    memset(ma.chunk_lo, 0, sizeof(ma.chunk_lo));
    if (guided) {
        ma.n_guided = gplan.n;
        memcpy(ma.chunk_lo, gplan.lo, sizeof(int) * (gplan.n + 1));
    }
    dim3 grid((unsigned)n_ftiles, (unsigned)n_chunks);
    if ((int)carve.total > smem_max)
        return set_error(DEISM_ERR_INVALID, "LC kernel needs %u B of shared memory, device has %d",
                         carve.total, smem_max);
    prof_begin("lc_main", st);
    if (resident) {
        DEISM_CUDA_TRY(cudaFuncSetAttribute(lc_main_kernel<LC_WN_RES, true, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)carve.total));
        DEISM_CUDA_TRY(cudaFuncSetAttribute(lc_main_kernel<LC_WN_RES, true, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)carve.total));
        lc_main_kernel<LC_WN_RES, true, true><<<grid, 128 * LC_WN_RES, carve.total, st>>>(ma);
        lc_main_kernel<LC_WN_RES, true, false><<<grid, 128 * LC_WN_RES, carve.total, st>>>(ma);
    } else {
        DEISM_CUDA_TRY(cudaFuncSetAttribute(lc_main_kernel<LC_WN_STR, false, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)carve.total));
        DEISM_CUDA_TRY(cudaFuncSetAttribute(lc_main_kernel<LC_WN_STR, false, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)carve.total));
        lc_main_kernel<LC_WN_STR, false, true><<<grid, 128 * LC_WN_STR, carve.total, st>>>(ma);
        lc_main_kernel<LC_WN_STR, false, false><<<grid, 128 * LC_WN_STR, carve.total, st>>>(ma);
    }
    prof_end("lc_main", st);
    count_launch();                          // two instantiations went out
    DEISM_LAUNCH_CHECK("lc_main_kernel");

    return launch_reduce_partials(partial, n_chunks, K, d_out, accumulate, st);
}
