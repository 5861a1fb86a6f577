// lc.cu — shoebox DEISM-LC accumulation with the wall reflection factors fused in.
//
// Replaces _numba_LC_matrix_batch (deism/parallel_backends.py:264-325), its batching driver
// (pb:594-654) and, in FUSED modes, the attenuation producers cd:2900-2928 / pb:140-186.
//
//   P[k] = sum_img  -atten[img,k] * 4pi/k * exp(-i k r_s)/(k r_s) * S[img,k] * R[img,k]
//   S[img,k] = sum_j i^{n_j} C_s[k,j] Y_{n_j}^{m_j}(theta_s,phi_s)       (source->receiver image)
//   R[img,k] = sum_j i^{v_j} C_r[k,j] Y_{v_j}^{u_j}(theta_r,phi_r)       (receiver->source image)
//
// Structure:
//   lc_prep_kernel   one thread per image: phase-folded harmonics Ys'[img,j], Yr'[img,j],
//                    r, 1/r, direction cosines + wall exponents, flat-impedance attenuation.
//   lc_main_kernel   CTA tile = 32 frequencies x 64 images, 128 threads, 4x4 (freq x image)
//                    register tile per thread: S and R are [I x J] . [J x K] complex
//                    contractions streamed through shared memory in chunks of JC harmonics;
//                    the epilogue applies exp(-ikr), 1/(k^2 r) and atten and accumulates
//                    P[k] in registers over the CTA's image chunk.
//   lc_reduce_kernel deterministic sum of the per-chunk partials into out[K].
#include "common.cuh"

namespace deism {

constexpr int LC_TF = 32;       // frequencies per CTA tile
constexpr int LC_TI = 64;       // images per inner tile
constexpr int LC_JC = 16;       // harmonics per shared-memory chunk
constexpr int LC_THREADS = 128; // 8 (freq) x 16 (image) threads, 4x4 pairs each

struct LcImage {        // 64 bytes per image
    double r, inv_r;
    double cx, cy, cz;
    unsigned char e[8];
    double2 atten_flat;  // c64-rounded (mode ROOM) attenuation for frequency-flat impedance
};

struct LcPrepArgs {
    long n_images, n_pad, K;
    int Js, Jr, Js_pad, Jr_pad;
    const float *R_s_rI, *R_r_sI;
    int mode, angdep;
    const int32_t *A;
    const float *R_sI_r;
    const cplx *Z;
    RoomGeom room;
};

// n_all/m_all/v_all/u_all live in constant memory (J <= 256 each: order <= 15)
constexpr int LC_JMAX = 256;
__constant__ int c_lc_n[LC_JMAX], c_lc_m[LC_JMAX], c_lc_v[LC_JMAX], c_lc_u[LC_JMAX];

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

// Per-image tables, padded to whole tiles: rows >= n_images and columns >= J are zero so that
// the main kernel's bulk copies never need a bounds check.
__global__ void __launch_bounds__(128)
lc_prep_kernel(LcPrepArgs a, LcImage *img_out, cplx *Ys, cplx *Yr) {
    long img = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= a.n_pad) return;
    LcImage rec;
    rec.cx = rec.cy = rec.cz = 1.0;
    for (int i = 0; i < 8; ++i) rec.e[i] = 0;
    // Y tables are tile-major [tile][j][image in tile]: a chunk of harmonics of one image tile is
    // one contiguous block (a single bulk copy) and lands conflict-free in shared memory
    const long ys_base = (img / LC_TI) * (long)a.Js_pad * LC_TI + (img % LC_TI);
    const long yr_base = (img / LC_TI) * (long)a.Jr_pad * LC_TI + (img % LC_TI);
    if (img >= a.n_images) {
        rec.r = 1.0; rec.inv_r = 0.0;           // weight 0: contributes exactly nothing
        rec.atten_flat = cmake(0.0, 0.0);
        img_out[img] = rec;
        for (int j = 0; j < a.Js_pad; ++j) Ys[ys_base + (long)j * LC_TI] = cmake(0.0, 0.0);
        for (int j = 0; j < a.Jr_pad; ++j) Yr[yr_base + (long)j * LC_TI] = cmake(0.0, 0.0);
        return;
    }
    double phi_s = (double)a.R_s_rI[img * 3 + 0], th_s = (double)a.R_s_rI[img * 3 + 1];
    rec.r = (double)a.R_s_rI[img * 3 + 2];
    rec.inv_r = 1.0 / rec.r;
    double phi_r = (double)a.R_r_sI[img * 3 + 0], th_r = (double)a.R_r_sI[img * 3 + 1];
    double s, ct_s, ct_r;
    sincos_cw(th_s, &s, &ct_s);
    sincos_cw(th_r, &s, &ct_r);
    for (int j = 0; j < a.Js_pad; ++j)
        Ys[ys_base + (long)j * LC_TI] = j < a.Js
            ? cmul_ipow(sph_harm_dev(c_lc_m[j], c_lc_n[j], phi_s, ct_s), c_lc_n[j]) : cmake(0.0, 0.0);
    for (int j = 0; j < a.Jr_pad; ++j)
        Yr[yr_base + (long)j * LC_TI] = j < a.Jr
            ? cmul_ipow(sph_harm_dev(c_lc_u[j], c_lc_v[j], phi_r, ct_r), c_lc_v[j]) : cmake(0.0, 0.0);
    AttenGeom g;
    g.cx = g.cy = g.cz = 1.0;
    for (int i = 0; i < 8; ++i) g.e[i] = 0;
    rec.atten_flat = cmake(1.0, 0.0);
    if (a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES) {
        if (a.mode == DEISM_ATTEN_FUSED_ROOM)
            atten_geom_room(a.A + img * 6, a.room.LL, a.room.xs, a.room.xr, a.angdep, g);
        else
            atten_geom_angles(a.A + img * 6, a.R_sI_r + img * 3, a.angdep, g);
        cplx at = shoebox_atten(a.Z, a.K, g.cx, g.cy, g.cz, g.e);  // Z[:,0]
        rec.atten_flat = a.mode == DEISM_ATTEN_FUSED_ROOM ? round_c64(at) : at;
    }
    rec.cx = g.cx; rec.cy = g.cy; rec.cz = g.cz;
    for (int i = 0; i < 8; ++i) rec.e[i] = g.e[i];
    img_out[img] = rec;
}

// complex64 [K][J] -> complex128 [K_pad/32][J_pad][32], zero padded: the harmonics j0..j0+JC of
// one frequency tile are one contiguous block (a single bulk copy)
__global__ void __launch_bounds__(256)
lc_coef_kernel(const float2 *__restrict__ in, long K, int J, long K_pad, int J_pad, cplx *out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K_pad * J_pad) return;
    long ft = i / ((long)J_pad * LC_TF);
    int j = (int)((i / LC_TF) % J_pad);
    long k = ft * LC_TF + (i % LC_TF);
    cplx v = cmake(0.0, 0.0);
    if (j < J && k < K) {
        float2 c = in[k * J + j];
        v = cmake((double)c.x, (double)c.y);
    }
    out[i] = v;
}

// ---- mbarrier / bulk-copy (TMA) primitives ----------------------------------------------
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

struct LcMainArgs {
    long n_images, n_pad, K, K_pad;
    int Js_pad, Jr_pad;      // padded harmonic counts = ncs*JCs, ncr*JCr
    int ncs, ncr, JCs, JCr;  // chunks per pass and harmonics per chunk
    const cplx *CTs, *CTr;   // [K_pad/32][J_pad][32]
    const cplx *Ys, *Yr;     // [n_pad/64][J_pad][64]
    const LcImage *img;      // [n_pad]
    const double *k;
    int mode;                // DEISM_ATTEN_*
    const void *atten;       // modes C64/C128
    const cplx *Z;           // fused modes
    const int *flags;        // flags[0]: impedance is frequency-flat
    long tiles_per_chunk;    // image tiles (of LC_TI) per CTA chunk
    cplx *partial;           // [n_chunks, K]
};

constexpr int LC_STAGES = 2;
struct LcStage {
    double2 C[LC_JC][LC_TF];
    double2 Y[LC_JC][LC_TI];
};
struct LcSmem {
    LcStage st[LC_STAGES];
    double2 fac[LC_TI][LC_TF + 1];   // factor[img][freq] of the current tile
    double2 red[16][LC_TF];
    double2 Z[6][LC_TF];
    double2 aflat[LC_TI];
    double r[LC_TI], invr[LC_TI];
    double cosd[3][LC_TI];
    unsigned char e[LC_TI][8];
    unsigned long long full[LC_STAGES];
};

__global__ void __launch_bounds__(LC_THREADS, 2)
lc_main_kernel(LcMainArgs a) {
    extern __shared__ __align__(128) unsigned char lc_smem_raw[];
    LcSmem &sm = *reinterpret_cast<LcSmem *>(lc_smem_raw);

    const int tid = threadIdx.x;
    const int tx = tid & 7;        // frequency lane: freqs f0 + tx + 8*q
    const int ty = tid >> 3;       // image lane:     images i0 + ty + 16*b
    const long f0 = (long)blockIdx.x * LC_TF;
    const long chunk = blockIdx.y;
    const long n_img_tiles = a.n_pad / LC_TI;
    long tile_lo = chunk * a.tiles_per_chunk;
    long tile_hi = tile_lo + a.tiles_per_chunk;
    if (tile_hi > n_img_tiles) tile_hi = n_img_tiles;
    const int npt = a.ncs + a.ncr;                         // chunks per tile
    const long total = tile_hi > tile_lo ? (tile_hi - tile_lo) * npt : 0;

    const bool fused = a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES;
    const bool flat = fused && a.flags[0] != 0;

    if (tid == 0) {
        for (int s = 0; s < LC_STAGES; ++s) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer: warp 0 stages chunk c (harmonics of one pass for one image tile) with bulk copies
    auto issue = [&](long c, int stage) {
        const long tile = tile_lo + c / npt;
        const int local = (int)(c % npt);
        const bool src_pass = local < a.ncs;
        const int JC = src_pass ? a.JCs : a.JCr;
        const int j0 = (src_pass ? local : local - a.ncs) * JC;
        const int Jp = src_pass ? a.Js_pad : a.Jr_pad;
        const cplx *CT = (src_pass ? a.CTs : a.CTr) + ((long)blockIdx.x * Jp + j0) * LC_TF;
        const cplx *Yg = (src_pass ? a.Ys : a.Yr) + (tile * Jp + j0) * LC_TI;
        LcStage &S = sm.st[stage];
        if ((tid & 31) == 0) {
            mbar_expect_tx(&sm.full[stage], (unsigned)(JC * (LC_TF + LC_TI) * 16));
            bulk_g2s(&S.C[0][0], CT, (unsigned)(JC * LC_TF * 16), &sm.full[stage]);
            bulk_g2s(&S.Y[0][0], Yg, (unsigned)(JC * LC_TI * 16), &sm.full[stage]);
        }
    };

    // per-thread frequency data
    double kf[4], wk[4];
    bool fvalid[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        long f = f0 + tx + 8 * q;
        fvalid[q] = f < a.K;
        kf[q] = fvalid[q] ? a.k[f] : 1.0;
        wk[q] = -4.0 * 3.14159265358979323846 / (kf[q] * kf[q]);   // -4pi/k^2
    }
    if (fused && !flat) {
        for (int i = tid; i < 6 * LC_TF; i += LC_THREADS) {
            int w = i / LC_TF, q = i % LC_TF;
            long f = f0 + q;
            sm.Z[w][q] = f < a.K ? a.Z[(long)w * a.K + f] : cmake(1.0, 0.0);
        }
    }

    cplx P[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) P[q] = cmake(0.0, 0.0);
    cplx S[4][4], R[4][4];   // [freq q][image b]

    if (tid < 32 && total > 0) issue(0, 0);

#pragma unroll 1
    for (long c = 0; c < total; ++c) {
        const int stage = (int)(c & 1);
        const unsigned parity = (unsigned)((c >> 1) & 1);
        const int local = (int)(c % npt);
        const long i0 = (tile_lo + c / npt) * LC_TI;

        if (local == 0) {
            // ---- new image tile: per-image records, then the factor
            //      fac[img][freq] = -atten * 4pi/k * exp(-ikr)/k/r   (pb:314-317),
            // computed before the contraction so that its registers are dead while the 4x4
            // S and R accumulators are live; each thread produces exactly the 16 pairs it
            // consumes in the epilogue.
            for (int i = tid; i < LC_TI; i += LC_THREADS) {
                LcImage rec = a.img[i0 + i];
                sm.r[i] = rec.r; sm.invr[i] = rec.inv_r;
                sm.cosd[0][i] = rec.cx; sm.cosd[1][i] = rec.cy; sm.cosd[2][i] = rec.cz;
#pragma unroll
                for (int b = 0; b < 8; ++b) sm.e[i][b] = rec.e[b];
                sm.aflat[i] = rec.atten_flat;
            }
            __syncthreads();
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int il = ty + 16 * b;
                const long img = i0 + il;
                const double r = sm.r[il], inv_r = sm.invr[il];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double sn, cs;
                    sincos_cw(kf[q] * r, &sn, &cs);
                    cplx att;
                    if (flat) {
                        att = sm.aflat[il];
                    } else if (fused) {
                        att = shoebox_atten(&sm.Z[0][tx + 8 * q], LC_TF, sm.cosd[0][il], sm.cosd[1][il],
                                            sm.cosd[2][il], sm.e[il]);
                        if (a.mode == DEISM_ATTEN_FUSED_ROOM) att = round_c64(att);
                    } else {
                        long f = f0 + tx + 8 * q;
                        att = cmake(0.0, 0.0);
                        if (img < a.n_images && fvalid[q]) {
                            if (a.mode == DEISM_ATTEN_C64) {
                                float2 v = ((const float2 *)a.atten)[img * a.K + f];
                                att = cmake((double)v.x, (double)v.y);
                            } else {
                                att = ((const cplx *)a.atten)[img * a.K + f];
                            }
                        }
                    }
                    double w = wk[q] * inv_r;
                    sm.fac[il][tx + 8 * q] = cmul(att, cmake(cs * w, -sn * w));
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int b = 0; b < 4; ++b) { S[q][b] = cmake(0.0, 0.0); R[q][b] = cmake(0.0, 0.0); }
        }

        // stage (c+1)&1 was released by the barrier that closed iteration c-1
        if (tid < 32 && c + 1 < total) issue(c + 1, stage ^ 1);
        mbar_wait(&sm.full[stage], parity);

        const LcStage &T = sm.st[stage];
        if (local < a.ncs) {
            const int jn = a.JCs;
#pragma unroll 4
            for (int j = 0; j < jn; ++j) {
                cplx cq[4], yb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) cq[q] = T.C[j][tx + 8 * q];
#pragma unroll
                for (int b = 0; b < 4; ++b) yb[b] = T.Y[j][ty + 16 * b];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < 4; ++b) cfma(S[q][b], cq[q], yb[b]);
            }
        } else {
            const int jn = a.JCr;
#pragma unroll 4
            for (int j = 0; j < jn; ++j) {
                cplx cq[4], yb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) cq[q] = T.C[j][tx + 8 * q];
#pragma unroll
                for (int b = 0; b < 4; ++b) yb[b] = T.Y[j][ty + 16 * b];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < 4; ++b) cfma(R[q][b], cq[q], yb[b]);
            }
        }
        __syncthreads();   // every warp is done with this stage (and, on a tile's last chunk,
                           // with the per-image arrays)

        if (local == npt - 1) {
            // ---- epilogue: P += factor * S * R ----
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int il = ty + 16 * b;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    cfma(P[q], cmul(sm.fac[il][tx + 8 * q], S[q][b]), R[q][b]);
            }
        }
    }

    // ---- reduce the 16 image lanes that share a frequency, write the chunk partial ----
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) sm.red[ty][tx + 8 * q] = P[q];
    __syncthreads();
    if (tid < LC_TF) {
        cplx s = cmake(0.0, 0.0);
#pragma unroll
        for (int y = 0; y < 16; ++y) s = cadd(s, sm.red[y][tid]);
        long f = f0 + tid;
        if (f < a.K) a.partial[chunk * a.K + f] = s;
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

int launch_z_flat(const deism_shoebox_atten *at, long K, int *flags, cudaStream_t st) {
    DEISM_CUDA_TRY(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
    if (at->mode == DEISM_ATTEN_FUSED_ROOM || at->mode == DEISM_ATTEN_FUSED_ANGLES) {
        int one = 1;
        DEISM_CUDA_TRY(cudaMemcpyAsync(flags, &one, sizeof(int), cudaMemcpyHostToDevice, st));
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

    int hn[LC_JMAX], hm[LC_JMAX], hv[LC_JMAX], hu[LC_JMAX];
    for (int j = 0; j < Js; ++j) {
        hn[j] = (int)h_n_all[j]; hm[j] = (int)h_m_all[j];
        if (hn[j] < 0 || hn[j] > 64 || abs(hm[j]) > hn[j])
            return set_error(DEISM_ERR_INVALID, "bad (n,m) = (%d,%d) at j=%d", hn[j], hm[j], j);
    }
    for (int j = 0; j < Jr; ++j) {
        hv[j] = (int)h_v_all[j]; hu[j] = (int)h_u_all[j];
        if (hv[j] < 0 || hv[j] > 64 || abs(hu[j]) > hv[j])
            return set_error(DEISM_ERR_INVALID, "bad (v,u) = (%d,%d) at j=%d", hv[j], hu[j], j);
    }
    DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_lc_n, hn, Js * sizeof(int), 0, cudaMemcpyHostToDevice, st));
    DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_lc_m, hm, Js * sizeof(int), 0, cudaMemcpyHostToDevice, st));
    DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_lc_v, hv, Jr * sizeof(int), 0, cudaMemcpyHostToDevice, st));
    DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_lc_u, hu, Jr * sizeof(int), 0, cudaMemcpyHostToDevice, st));
    // the host arrays above are on the stack: the copies must have been consumed before return
    DEISM_CUDA_TRY(cudaStreamSynchronize(st));

    const long n_img_tiles = (n_images + LC_TI - 1) / LC_TI;
    const long n_pad = n_img_tiles * LC_TI;
    const long n_ftiles = (K + LC_TF - 1) / LC_TF;
    const long K_pad = n_ftiles * LC_TF;
    // harmonics are streamed in ncs (ncr) chunks of JCs (JCr) <= LC_JC each, sized to waste
    // as little padding as possible (J = 36 -> 3 x 12)
    const int ncs = (Js + LC_JC - 1) / LC_JC, ncr = (Jr + LC_JC - 1) / LC_JC;
    const int JCs = (Js + ncs - 1) / ncs, JCr = (Jr + ncr - 1) / ncr;
    const int Js_pad = ncs * JCs, Jr_pad = ncr * JCr;
    long n_chunks = (long)sm_count() * 2 * 8 / n_ftiles;   // ~8 waves of 2 CTAs/SM
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > n_img_tiles) n_chunks = n_img_tiles;
    if (n_chunks > 65535) n_chunks = 65535;
    long tiles_per_chunk = (n_img_tiles + n_chunks - 1) / n_chunks;
    n_chunks = (n_img_tiles + tiles_per_chunk - 1) / tiles_per_chunk;

    LcImage *img = (LcImage *)workspace(WS_LC_IMG, (size_t)n_pad * sizeof(LcImage));
    cplx *Ys = (cplx *)workspace(WS_LC_YS, (size_t)n_pad * Js_pad * sizeof(cplx));
    cplx *Yr = (cplx *)workspace(WS_LC_YR, (size_t)n_pad * Jr_pad * sizeof(cplx));
    cplx *CT = (cplx *)workspace(WS_SMALL, (size_t)(Js_pad + Jr_pad) * K_pad * sizeof(cplx));
    cplx *partial = (cplx *)workspace(WS_PARTIAL, (size_t)n_chunks * K * sizeof(cplx));
    int *flags = (int *)workspace(WS_FLAGS, 64);
    if (!img || !Ys || !Yr || !CT || !partial || !flags) return DEISM_ERR_NOMEM;
    cplx *CTs = CT, *CTr = CT + (size_t)Js_pad * K_pad;

    rc = launch_z_flat(atten, K, flags, st);
    if (rc) return rc;

    LcPrepArgs pa;
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
    lc_coef_kernel<<<(unsigned)((K_pad * Js_pad + 255) / 256), 256, 0, st>>>(
        (const float2 *)d_C_s_vec, K, Js, K_pad, Js_pad, CTs);
    DEISM_LAUNCH_CHECK("lc_coef_kernel");
    lc_coef_kernel<<<(unsigned)((K_pad * Jr_pad + 255) / 256), 256, 0, st>>>(
        (const float2 *)d_C_r_vec, K, Jr, K_pad, Jr_pad, CTr);
    DEISM_LAUNCH_CHECK("lc_coef_kernel");

    LcMainArgs ma;
    ma.n_images = n_images; ma.n_pad = n_pad; ma.K = K; ma.K_pad = K_pad;
    ma.Js_pad = Js_pad; ma.Jr_pad = Jr_pad; ma.ncs = ncs; ma.ncr = ncr; ma.JCs = JCs; ma.JCr = JCr;
    ma.CTs = CTs; ma.CTr = CTr; ma.Ys = Ys; ma.Yr = Yr; ma.img = img; ma.k = d_k;
    ma.mode = atten->mode; ma.atten = atten->d_atten; ma.Z = (const cplx *)atten->d_Z_S;
    ma.flags = flags; ma.tiles_per_chunk = tiles_per_chunk; ma.partial = partial;
    dim3 grid((unsigned)n_ftiles, (unsigned)n_chunks);
    DEISM_CUDA_TRY(cudaFuncSetAttribute(lc_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(LcSmem)));
    prof_begin("lc_main", st);
    lc_main_kernel<<<grid, LC_THREADS, sizeof(LcSmem), st>>>(ma);
    prof_end("lc_main", st);
    DEISM_LAUNCH_CHECK("lc_main_kernel");

    return launch_reduce_partials(partial, n_chunks, K, d_out, accumulate, st);
}
