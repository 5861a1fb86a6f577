// org.cu — DEISM-ORG accumulation: the Wigner-3j coupled bilinear sum over source and receiver
// spherical-harmonic directivity coefficients.
//
// Replaces _numba_ORG_batch (deism/parallel_backends.py:189-261), _numba_ARG_ORG_batch
// (pb:388-460) and the batching driver pb:543-591.
//
// Reference form, one image:
//   P[k] = sum_{n,m,v,u} mirror * atten[k] * C_s[k,n,m] * S * C_r[k,v,-u] * (i/k) * (-1)^u
//   S    = 4pi i^{v-n} (-1)^{m'} sum_l i^l h_l(kr) Y_l^{m'-u}(theta,phi) W1[n,v,l] W2[n,v,l,m',u] Xi
//   mirror = (-1)^{(p_y+p_z) m + p_z n},  m' = (-1)^{p_x+p_y} m,  Xi = sqrt((2n+1)(2v+1)(2l+1)/4pi)
//
// W1 = (n v l; 0 0 0) vanishes unless n+v+l is even, so i^{v-n+l} = +-1 and every quintuple
// carries a REAL coefficient  G = 4pi (-1)^{(v-n+l)/2} (-1)^{m'+u} W1 W2 Xi.  Re-ordering the
// sum by (l, mu = m'-u):
//
//   P[img,k] = atten[img,k] * sum_l h_l(k r_img) sum_mu Y_l^mu(img) * B[par(img)][l,mu][k]
//   B[par][l,mu][k] = (i/k) sum_{(n,m,v,u): m'-u = mu} mirror*G * C_s[k,n,m] * C_r[k,v,-u]
//
// Kernels:
//   org_btable2_kernel  B for the 8 parity classes (image independent): CTA = a tile of frequencies
//                       x all (l, |mu|) jobs, coefficient tile in shared memory      [algo 1]
//                       (org_btable_kernel: the earlier form, kept behind ORG_BTABLE_V1)
//   org_sort/prep       images grouped by parity class, Y_l^mu(img) rows (real harmonics, signed for
//                       the Clenshaw sum), attenuation geometry
//   org_consume_kernel  64 image x 32 freq CTA tile on the FP64 tensor cores; per l a (2l+1)-row
//                       real x complex contraction from TMA-staged shared memory; the sum over l by
//                       Clenshaw's recurrence with the tensor core adding e_{l+2}
//   org_small_kernel    the same factorised sum in plain FMAs for a few images (MIX early images)
//   org_direct_kernel   the quintuple sum done per (image, frequency) pair with per-image
//                       source coefficients: the ARG form (pb:388-460) and the shoebox
//                       cross-check [algo 2]
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace deism {

int launch_reduce_partials(const cplx *partial, long n_chunks, long K, double *d_out, int accumulate,
                           cudaStream_t st);
int validate_atten(const deism_shoebox_atten *at);
int launch_z_flat(const deism_shoebox_atten *at, long K, int *flags, cudaStream_t st);
int launch_k_uniform(const double *d_k, long K, int *flags, cudaStream_t st);   // flags[1], dk at flags+2

// One non-zero quintuple (n,m,v,u,l) of the coupling, stored in the segment of its (l,mu).
struct OrgEntry {
    double G;        // real coefficient without the mirror sign
    int16_t idx_s;   // n*(2N+1) + wrap(m)      -> C_s[k, n, m]
    int16_t idx_r;   // v*(2V+1) + wrap(-u)     -> C_r[k, v, -u]
    int8_t m, n;     // for the mirror sign (-1)^{b m + c n}
    int16_t pad;
};
static_assert(sizeof(OrgEntry) == 16, "OrgEntry must be 16 bytes");

// one 128-bit load + register decode (a struct copy with int8 fields goes through local memory)
struct OrgEntryReg { double G; int idx_s, idx_r, m, n; };
__device__ __forceinline__ OrgEntryReg load_entry(const OrgEntry *p) {
    int4 raw = __ldg(reinterpret_cast<const int4 *>(p));
    OrgEntryReg e;
    e.G = __hiloint2double(raw.y, raw.x);
    e.idx_s = (int)(short)(raw.z & 0xffff);
    e.idx_r = (int)(short)((raw.z >> 16) & 0xffff);
    e.m = (int)(signed char)(raw.w & 0xff);
    e.n = (int)(signed char)((raw.w >> 8) & 0xff);
    return e;
}

// B-table form of the same list.  The class-1 list (m' = -m) is the class-0 list with the source
// coefficient of (n, -m) in place of (n, m): same (v, u), same G, same segment.  One entry therefore
// feeds both classes from ONE receiver-coefficient load.  Inside a segment the entries are grouped
// by the parities (m & 1, n & 1) that decide the mirror sign (-1)^{b m + c n}, then ordered by (n, m)
// so that consecutive entries reuse the source coefficients.
struct BtEntry {
    double G;
    int16_t idx_s;    // n*(2N+1) + wrap(m)    -> C_s[k, n, m]     (class 0)
    int16_t idx_sn;   // n*(2N+1) + wrap(-m)   -> C_s[k, n, -m]    (class 1)
    int16_t idx_r;    // v*(2V+1) + wrap(-u)   -> C_r[k, v, -u]
    int16_t pad;
};
static_assert(sizeof(BtEntry) == 16, "BtEntry must be 16 bytes");

struct OrgLists {
    // class a = (p_x+p_y)&1 selects m' = m (a=0) or -m (a=1)
    std::vector<OrgEntry> entries[2];
    std::vector<int> seg[2];   // Lmax^2 + 1 offsets, segment index = l*l + l + mu
    std::vector<BtEntry> bt;   // class-0 entries, by (segment, parity group, n, m, v)
    std::vector<int> bt_seg;   // 4 * Lmax^2 + 1 offsets, index = 4 * segment + 2 * (m & 1) + (n & 1)
    std::vector<int2> jobs;    // (l, |mu|) pairs, most entries first
    long nnz = 0;
};

static void build_lists(int N, int V, const float *W1, const float *W2, OrgLists &L) {
    const int Lmax = N + V + 1, M2 = 2 * N + 1, U2 = 2 * V + 1, nseg = Lmax * Lmax;
    for (int a = 0; a < 2; ++a) {
        std::vector<std::vector<OrgEntry>> buckets(nseg);
        for (int n = 0; n <= N; ++n)
        for (int m = -n; m <= n; ++m) {
            int mp = a ? -m : m;
            for (int v = 0; v <= V; ++v)
            for (int u = -v; u <= v; ++u)
            for (int l = abs(n - v); l <= n + v; ++l) {
                if (abs(u - mp) > l) continue;
                long i1 = ((long)n * (V + 1) + v) * Lmax + l;
                long i2 = (i1 * M2 + (mp < 0 ? mp + M2 : mp)) * U2 + (u < 0 ? u + U2 : u);
                // complex64 tables: a value counts as zero only if both parts are (pb:229)
                double w1 = (double)W1[2 * i1], w1i = (double)W1[2 * i1 + 1];
                double w2 = (double)W2[2 * i2], w2i = (double)W2[2 * i2 + 1];
                if ((w1 == 0.0 && w1i == 0.0) || (w2 == 0.0 && w2i == 0.0)) continue;
                double Xi = sqrt((double)((2 * n + 1) * (2 * v + 1) * (2 * l + 1)) /
                                 (4.0 * 3.14159265358979323846));
                // i^{v-n+l}: real (+-1) whenever W1 != 0; guard the impossible odd case
                int e = v - n + l;
                if (e & 1) continue;
                double ph = ((e / 2) & 1) ? -1.0 : 1.0;
                double sg = ((mp + u) & 1) ? -1.0 : 1.0;
                OrgEntry en;
                en.G = 4.0 * 3.14159265358979323846 * ph * sg * w1 * w2 * Xi;
                en.idx_s = (int16_t)(n * M2 + (m < 0 ? m + M2 : m));
                en.idx_r = (int16_t)(v * U2 + ((-u) < 0 ? -u + U2 : -u));
                en.m = (int8_t)m; en.n = (int8_t)n; en.pad = 0;
                int mu = mp - u;
                buckets[l * l + l + mu].push_back(en);
            }
        }
        L.seg[a].assign(nseg + 1, 0);
        L.entries[a].clear();
        for (int s = 0; s < nseg; ++s) {
            L.seg[a][s] = (int)L.entries[a].size();
            L.entries[a].insert(L.entries[a].end(), buckets[s].begin(), buckets[s].end());
        }
        L.seg[a][nseg] = (int)L.entries[a].size();
    }
    L.nnz = (long)L.entries[0].size();
    // ---- B-table lists from the class-0 buckets (already in (n, m, v) order inside a segment) ----
    L.bt.clear();
    L.bt_seg.assign(4 * nseg + 1, 0);
    for (int sgm = 0; sgm < nseg; ++sgm) {
        for (int pg = 0; pg < 4; ++pg) {
            L.bt_seg[4 * sgm + pg] = (int)L.bt.size();
            for (int j = L.seg[0][sgm]; j < L.seg[0][sgm + 1]; ++j) {
                const OrgEntry &e = L.entries[0][j];
                if ((((int)e.m & 1) << 1 | ((int)e.n & 1)) != pg) continue;
                BtEntry b;
                b.G = e.G; b.idx_s = e.idx_s; b.idx_r = e.idx_r; b.pad = 0;
                const int m = e.m, n = e.n;
                b.idx_sn = (int16_t)(n * M2 + ((-m) < 0 ? -m + M2 : -m));
                L.bt.push_back(b);
            }
        }
    }
    L.bt_seg[4 * nseg] = (int)L.bt.size();
    L.jobs.clear();
    for (int l = 0; l < Lmax; ++l)
        for (int am = 0; am <= l; ++am) L.jobs.push_back(make_int2(l, am));
    auto job_size = [&](const int2 &j) {
        const int s1 = j.x * j.x + j.x + j.y, s2 = j.x * j.x + j.x - j.y;
        return (L.bt_seg[4 * s1 + 4] - L.bt_seg[4 * s1]) + (j.y ? L.bt_seg[4 * s2 + 4] - L.bt_seg[4 * s2] : 0);
    };
    std::stable_sort(L.jobs.begin(), L.jobs.end(),
                     [&](const int2 &a, const int2 &b) { return job_size(a) > job_size(b); });
}

// ---------------------------------------------------------------------------------------
// coefficient transposes: complex64 [K][J] -> complex64 [J][K] (the B-table build streams them
// once per coupling entry: keeping them at 8 bytes halves its L2 traffic)
// ---------------------------------------------------------------------------------------
// The transposed tables are widened to complex128 once: the B-table loop then spends no FP64-pipe
// slots on F32->F64 conversions (4 per coupling entry; measured 0.577 -> 0.522 ms at C3 size, the
// tables stay L2-resident at 16 bytes per element: 15 MB at C3 size).
typedef double2 bt_elem;
__device__ __forceinline__ bt_elem bt_make(float2 v) { return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ bt_elem bt_make(double2 v) { return v; }
// IN = float2 (the reference's complex64 storage, cd:1184) or double2 (a caller's own complex128
// coefficients, which the reference's ORG path would use at full precision, pb:556-557)
template <typename IN>
__global__ void transpose_coef_kernel(const IN *__restrict__ in, long K, int J, bt_elem *out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * J) return;
    long k = i / J;
    int j = (int)(i % J);
    out[(long)j * K + k] = bt_make(in[i]);
}

// ---------------------------------------------------------------------------------------
// factorised path: operand-plane geometry.  Segment (l, mu) lives at row rowoff(l) + mu + l of a
// plane set; the block of an odd l has one zero pad row (2l + 2 rows = whole DMMA k-steps), the block of an
// even l none (see org_consume_kernel).
// ---------------------------------------------------------------------------------------
#ifndef ORG_TI_
#define ORG_TI_ 64
#endif
constexpr int ORG_TI = ORG_TI_;   // images per tile
constexpr int ORG_TF = 32;        // frequencies per tile
constexpr int ORG_WM = ORG_TI / 16;            // warps along images (16 images each)
constexpr int ORG_WN = 4;                      // warps along frequency (8 frequencies each)
constexpr int ORG_THREADS = 32 * ORG_WM * ORG_WN;
constexpr int ORG_CTAS_PER_SM = ORG_THREADS <= 256 ? 2 : 1;
#ifndef ORG_RC_
#define ORG_RC_ 96                // 2 stages of 104 KB: the bulk-copy engine takes one stage at a time and
#endif                            // >= ~550 clk per stage however small (scripts/probes/tma_probe.cu)
constexpr int ORG_RC = ORG_RC_;   // (l,mu) rows per staged chunk (whole degrees; grows to the widest degree)
constexpr int ORG_OS = ORG_TI + 4;// Y row stride in doubles (tile + 4 pad: 4 mod 16 -> conflict-free fragments)
constexpr int ORG_BS = 34;        // B plane stride: a row of 2 planes is 68 = 4 mod 16 doubles
constexpr int ORG_LMAX = 48;
constexpr int ORG_MAXCHUNKS = 256;
#ifndef ORG_MAXST
#define ORG_MAXST 4               // most ring stages
#endif
constexpr int SORT_THREADS = 1024;
__constant__ int c_org_rowoff[ORG_LMAX + 1];
// normalisation sqrt((2l+1)/(4 pi) (l-|mu|)!/(l+|mu|)!) of Y_l^mu, index l (l + 1) / 2 + |mu|: computed on
// the host with the operations of _sph_harm_numba (pb:70-76: the factorial ratio as a running
// product, then one division and one square root — IEEE on both sides, so the table holds the values
// the per-element evaluation produced)
__device__ double g_org_norm[ORG_LMAX * (ORG_LMAX + 1) / 2];
__constant__ int4 c_org_chunks[ORG_MAXCHUNKS];   // row0, nrows, first degree, one past the last degree

// Real-harmonic rows.  Y_l^{-mu} = (-1)^mu conj(Y_l^mu), so with Y_l^mu = a + i b (mu > 0)
//     sum_mu Y_l^mu B[l,mu]  =  a_0 B[l,0] + sum_{mu>0} a_mu (B[l,mu] + (-1)^mu B[l,-mu])
//                                          + b_mu i (B[l,mu] - (-1)^mu B[l,-mu])
// is a REAL-by-complex contraction: two real GEMMs instead of the three of the 3-multiplication
// complex form.  Row order inside the block of degree l: a_0, a_1, b_1, a_2, b_2, ...
//
// B planes: [8 parity][K_pad/32][R_tot][Re|Im][ORG_BS] doubles.
// grid (ceil(K/128), Lmax(Lmax+1)/2 pairs (l, mu >= 0), 2 m'-classes); each thread evaluates the
// coupling D = C_s*C_r of an entry ONCE and feeds the four mirror-sign variants (b, c) of its class.
struct BtableArgs {
    long K, n_ftiles;
    int nseg, R_tot;
    const OrgEntry *ent[2];
    const int *seg[2];
    const bt_elem *CsT, *CrT; // [nm][K] complex128 (widened copies of the caller's complex64)
    const double *k;
    double *Bp;
};

__global__ void __launch_bounds__(128) org_btable_kernel(BtableArgs a) {
    long k = (long)blockIdx.x * 128 + threadIdx.x;
    const int cls = blockIdx.z;
    if (k >= a.K) return;
    int l = 0;
    while ((l + 1) * (l + 2) / 2 <= (int)blockIdx.y) ++l;
    const int am = (int)blockIdx.y - l * (l + 1) / 2;        // |mu|
    const OrgEntry *ent = cls ? a.ent[1] : a.ent[0];
    const int *segp = cls ? a.seg[1] : a.seg[0];
    const bt_elem *cs_base = a.CsT + k, *cr_base = a.CrT + k;
    const unsigned K32 = (unsigned)a.K;
    cplx acc[2][4];     // [mu sign][mirror variant]
#pragma unroll
    for (int sg = 0; sg < 2; ++sg)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[sg][v] = cmake(0.0, 0.0);
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
        if (sg == 1 && am == 0) break;
        const int s = l * l + l + (sg ? -am : am);
        const int lo = segp[s], hi = segp[s + 1];
#pragma unroll 4
        for (int j = lo; j < hi; ++j) {
            OrgEntryReg e = load_entry(ent + j);
            // 32-bit row offsets (the launcher checks rows * K < 2^31)
            const bt_elem cs = __ldg(cs_base + (unsigned)e.idx_s * K32), cr = __ldg(cr_base + (unsigned)e.idx_r * K32);
            cplx d = cmul(cmake((double)cs.x, (double)cs.y), cmake((double)cr.x, (double)cr.y));
            // mirror variants v = 2*b + c carry the sign (-1)^{b m + c n}: flip G's sign bit
            const int ghi = __double2hiint(e.G), glo = __double2loint(e.G);
            const int sn = (e.n & 1) << 31, sm_ = (e.m & 1) << 31;
            const double Gv[4] = {e.G, __hiloint2double(ghi ^ sn, glo), __hiloint2double(ghi ^ sm_, glo),
                                  __hiloint2double(ghi ^ sn ^ sm_, glo)};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                acc[sg][v].x = fma(Gv[v], d.x, acc[sg][v].x);
                acc[sg][v].y = fma(Gv[v], d.y, acc[sg][v].y);
            }
        }
    }
    const long row_a = c_org_rowoff[l] + (am ? 2 * am - 1 : 0);
    const long ft = k / ORG_TF;
    const int q = (int)(k % ORG_TF);
    const double ik = 1.0 / a.k[k];
    const double sm = (am & 1) ? -1.0 : 1.0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int b = v >> 1, c = v & 1;
        const int pz = c, py = b ^ c, px = cls ^ py;
        const int par = px * 4 + py * 2 + pz;
        // B = (i/k) acc;  a-row: B+ + (-1)^mu B-;  b-row: i (B+ - (-1)^mu B-)
        const double sx = acc[0][v].x + sm * acc[1][v].x, sy = acc[0][v].y + sm * acc[1][v].y;
        const double dx = acc[0][v].x - sm * acc[1][v].x, dy = acc[0][v].y - sm * acc[1][v].y;
        double *p = a.Bp + (((long)par * a.n_ftiles + ft) * a.R_tot + row_a) * 2 * ORG_BS + q;
        p[0] = -sy * ik;            // (i/k)(sx + i sy) = (-sy + i sx)/k
        p[ORG_BS] = sx * ik;
        if (am) {
            p += 2 * ORG_BS;        // i (i/k)(dx + i dy) = (-dx - i dy)/k
            p[0] = -dx * ik;
            p[ORG_BS] = -dy * ik;
        }
    }
}

// ---------------------------------------------------------------------------------------
// B table, second form (round 2): CTA = FT frequencies x every (l, |mu|) job.
//
// The first form streams the transposed coefficient tables from L2 once per coupling entry and is
// bound by that stream (64 B/clk/SM: 66 clk per warp-entry against 26 clk of FP64 work).  Here the
// coefficient tile of the CTA's FT frequencies — every (n, m) row of C_s and C_r, widened to
// complex128, 118 KB at N = V = 10 and FT = 16 — is staged in shared memory once, straight from
// the caller's [K][(n,m)] tables (no transposed copies), and every job reads it from there.
//   * lane = (frequency kk, slice h): the 32 / FT slices of a warp split a sub-segment's entries
//     and are summed by shuffles at the end of the job;
//   * one entry feeds both m'-classes from one receiver-coefficient load (BtEntry);
//   * entries of a sub-segment share their mirror parities: the four mirror-sign variants are
//     formed from the sub-segment's two sums instead of four FMAs per entry (6 FP64 instructions
//     per entry and class instead of 12);
//   * warps pull jobs (largest first) from a counter in shared memory.
// ---------------------------------------------------------------------------------------
struct Btable2Args {
    long K, n_ftiles;
    int R_tot, NMs, NMr, n_jobs, n_split;
    const BtEntry *ent;
    const int *seg;           // 4 * nseg + 1
    const int2 *jobs;
    const void *Cs, *Cr;      // the caller's [K][NMs] / [K][NMr] tables (complex64 or complex128)
    const double *k;
    double *Bp;
};

constexpr int BT2_THREADS = 512;

template <int FT, typename IN>
__global__ void __launch_bounds__(BT2_THREADS, 1) org_btable2_kernel(Btable2Args a) {
    constexpr int RS = FT + 1;        // row stride in double2 (odd: staging stores spread over the banks)
    constexpr int NH = 32 / FT;       // entry slices per warp
    extern __shared__ __align__(16) unsigned char bt_smem_raw[];
    double2 *CsS = reinterpret_cast<double2 *>(bt_smem_raw);
    double2 *CrS = CsS + (size_t)a.NMs * RS;
    int *counter = reinterpret_cast<int *>(CrS + (size_t)a.NMr * RS);
    const int tid = threadIdx.x, lane = tid & 31;
    int4 *ebuf = reinterpret_cast<int4 *>(counter + 4) + (tid >> 5) * 32;     // this warp's entry batch
    const long k0 = (long)blockIdx.x * FT;
    {
        const IN *ins = reinterpret_cast<const IN *>(a.Cs), *inr = reinterpret_cast<const IN *>(a.Cr);
        for (int i = tid; i < a.NMs * FT; i += BT2_THREADS) {
            const int kk = i / a.NMs, j = i - kk * a.NMs;
            CsS[j * RS + kk] = k0 + kk < a.K ? bt_make(ins[(k0 + kk) * a.NMs + j]) : make_double2(0.0, 0.0);
        }
        for (int i = tid; i < a.NMr * FT; i += BT2_THREADS) {
            const int kk = i / a.NMr, j = i - kk * a.NMr;
            CrS[j * RS + kk] = k0 + kk < a.K ? bt_make(inr[(k0 + kk) * a.NMr + j]) : make_double2(0.0, 0.0);
        }
        if (tid == 0) *counter = 0;
    }
    __syncthreads();
    const int kk = lane % FT, h = lane / FT;
    const long k = k0 + kk;
    const double ik = k < a.K ? 1.0 / a.k[k] : 0.0;
    const double2 *csl = CsS + kk, *crl = CrS + kk;
    const long ft = k / ORG_TF;
    const int q = (int)(k % ORG_TF);

    for (;;) {
        int jn = 0;
        if (lane == 0) jn = atomicAdd(counter, 1);
        jn = __shfl_sync(0xffffffffu, jn, 0);
        const int job = (int)blockIdx.y + a.n_split * jn;      // this CTA's share: every n_split-th job
        if (job >= a.n_jobs) break;
        const int2 jb = a.jobs[job];
        const int l = jb.x, am = jb.y;
        cplx V[2][2][4];      // [class][mu sign][mirror variant 2b + c]
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int sg = 0; sg < 2; ++sg)
#pragma unroll
                for (int v = 0; v < 4; ++v) V[c][sg][v] = cmake(0.0, 0.0);
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            if (sg == 1 && am == 0) break;
            const int sgm = l * l + l + (sg ? -am : am);
#pragma unroll
            for (int pg = 0; pg < 4; ++pg) {
                const int lo = a.seg[4 * sgm + pg], hi = a.seg[4 * sgm + pg + 1];
                if (hi == lo) continue;
                const int per = (hi - lo + NH - 1) / NH;
                const int my_lo = lo + h * per;
                const int my_hi = my_lo + per < hi ? my_lo + per : hi;
                cplx a0 = cmake(0.0, 0.0), a1 = cmake(0.0, 0.0);
                int prev = -1;
                double2 cs0 = make_double2(0.0, 0.0), cs1 = cs0;
                // Entries are fetched FT per slice at a time, one per lane (a coalesced 16 x FT byte
                // read per slice), parked in the warp's shared-memory buffer and consumed from there:
                // a 16-byte entry read straight from global memory by every lane is a 32-byte-sector
                // L2 round trip every other entry (measured: 440 clk per entry).  The next batch is in
                // flight while the current one is consumed.
                auto fetch = [&](int b) {
                    const int j = my_lo + b + kk;
                    int4 raw = __ldg(reinterpret_cast<const int4 *>(a.ent + (j < my_hi ? j : hi - 1)));
                    if (j >= my_hi) { raw.x = 0; raw.y = 0; }      // G = 0: a padding entry adds nothing
                    return raw;
                };
                int4 nxt = fetch(0);
#pragma unroll 1
                for (int b = 0; b < per; b += FT) {
                    __syncwarp();
                    ebuf[lane] = nxt;
                    __syncwarp();
                    if (b + FT < per) nxt = fetch(b + FT);
                    const int nt = per - b < FT ? per - b : FT;
#pragma unroll 2
                    for (int t = 0; t < nt; ++t) {
                        const int4 raw = ebuf[h * FT + t];
                        const double G = __hiloint2double(raw.y, raw.x);
                        const int is = (int)(short)(raw.z & 0xffff), isn = (int)(short)((raw.z >> 16) & 0xffff);
                        const int ir = (int)(short)(raw.w & 0xffff);
                        if (is != prev) { cs0 = csl[is * RS]; cs1 = csl[isn * RS]; prev = is; }
                        const double2 cr = crl[ir * RS];
                        const cplx d0 = cmul(cs0, cr), d1 = cmul(cs1, cr);
                        a0.x = fma(G, d0.x, a0.x); a0.y = fma(G, d0.y, a0.y);
                        a1.x = fma(G, d1.x, a1.x); a1.y = fma(G, d1.y, a1.y);
                    }
                }
                // mirror variants v = 2b + c carry (-1)^{b m + c n}; pg = 2 (m & 1) + (n & 1)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const bool neg = ((((v >> 1) & (pg >> 1)) ^ ((v & 1) & (pg & 1))) & 1) != 0;
                    if (neg) {
                        V[0][sg][v].x -= a0.x; V[0][sg][v].y -= a0.y;
                        V[1][sg][v].x -= a1.x; V[1][sg][v].y -= a1.y;
                    } else {
                        V[0][sg][v].x += a0.x; V[0][sg][v].y += a0.y;
                        V[1][sg][v].x += a1.x; V[1][sg][v].y += a1.y;
                    }
                }
            }
        }
        // sum the entry slices of the warp
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int sg = 0; sg < 2; ++sg)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
#pragma unroll
                    for (int o = FT; o < 32; o <<= 1) {
                        V[c][sg][v].x += __shfl_xor_sync(0xffffffffu, V[c][sg][v].x, o);
                        V[c][sg][v].y += __shfl_xor_sync(0xffffffffu, V[c][sg][v].y, o);
                    }
                }
        if (h == 0 && k < a.K) {
            const long row_a = c_org_rowoff[l] + (am ? 2 * am - 1 : 0);
            const double sm = (am & 1) ? -1.0 : 1.0;
#pragma unroll
            for (int cls = 0; cls < 2; ++cls)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int b = v >> 1, c = v & 1;
                    const int pz = c, py = b ^ c, px = cls ^ py;
                    const int par = px * 4 + py * 2 + pz;
                    // B = (i/k) acc;  a-row: B+ + (-1)^mu B-;  b-row: i (B+ - (-1)^mu B-)
                    const cplx p_ = V[cls][0][v], m_ = V[cls][1][v];
                    const double sx = p_.x + sm * m_.x, sy = p_.y + sm * m_.y;
                    const double dx = p_.x - sm * m_.x, dy = p_.y - sm * m_.y;
                    double *p = a.Bp + (((long)par * a.n_ftiles + ft) * a.R_tot + row_a) * 2 * ORG_BS + q;
                    p[0] = -sy * ik;            // (i/k)(sx + i sy) = (-sy + i sx)/k
                    p[ORG_BS] = sx * ik;
                    if (am) {
                        p += 2 * ORG_BS;        // i (i/k)(dx + i dy) = (-dx - i dy)/k
                        p[0] = -dx * ik;
                        p[ORG_BS] = -dy * ik;
                    }
                }
        }
    }
}

// ---------------------------------------------------------------------------------------
// image grouping by parity class (stable counting sort, one CTA) and per-image tables
// ---------------------------------------------------------------------------------------
struct OrgImage {   // 96 bytes
    double r, inv_r;
    double cx, cy, cz;
    unsigned char e[8];
    double2 atten_flat;
    double2 rot;    // exp(-i dk r): one frequency step on a uniform wave-number grid
    long orig;      // index in the caller's image list (-1 = padding)
    int par, pad;
};

// meta[0..8]   : first padded position of class c (multiple of ORG_TI); meta[8] = total padded
// meta[9..16]  : number of real images of class c (they sit at the start of the class's positions)
// perm[pos]    : original image index or -1
__global__ void __launch_bounds__(SORT_THREADS)
org_sort_kernel(const int32_t *__restrict__ A, long n_images, long *perm, long *meta, long perm_cap,
                int par_smem) {
    // Stable counting sort by parity class in one CTA: thread t owns the contiguous slice
    // [t * per, (t + 1) * per) of the image list, so "stable" = (thread order, then order inside the
    // slice).  Per class: counts per thread -> exclusive scan over the threads (warp shuffles + one
    // pass over the warp totals) -> every thread places its own images in order.
    __shared__ long base[9];
    __shared__ int wsum[8][SORT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long per = (n_images + SORT_THREADS - 1) / SORT_THREADS;
    const long lo = (long)tid * per, hi = lo + per < n_images ? lo + per : n_images;
    // parity classes of all images, read once with coalesced loads into shared memory (when they
    // fit: par_smem != 0); the slices then walk shared memory instead of strided global memory
    extern __shared__ unsigned char sort_par[];
    auto parity = [&](long i) {
        const int32_t *a = A + i * 6;
        return ((a[3] & 1) << 2) | ((a[4] & 1) << 1) | (a[5] & 1);
    };
    if (par_smem) {
        for (long i = tid; i < n_images; i += SORT_THREADS) sort_par[i] = (unsigned char)parity(i);
        __syncthreads();
    }
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long i = lo; i < hi; ++i) {
        const int par = par_smem ? (int)sort_par[i] : parity(i);
#pragma unroll
        for (int c = 0; c < 8; ++c) cnt[c] += (par == c);
    }
    int off[8];        // images of class c in lower threads of this warp
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int v = cnt[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        off[c] = v - cnt[c];
        if (lane == 31) wsum[c][warp] = v;
    }
    __syncthreads();
    if (tid < 8) {     // exclusive scan of the warp totals of class tid
        int run = 0;
        for (int w = 0; w < SORT_THREADS / 32; ++w) { const int t = wsum[tid][w]; wsum[tid][w] = run; run += t; }
        base[tid] = run;      // class total, for now
    }
    __syncthreads();
    if (tid == 0) {
        long pos = 0;
        for (int c = 0; c < 8; ++c) {
            const long n = base[c];
            meta[c] = pos;
            meta[9 + c] = n;               // real images of class c: positions meta[c] .. meta[c] + n - 1
            base[c] = pos;
            pos += (n + ORG_TI - 1) / ORG_TI * ORG_TI;
        }
        meta[8] = pos;
        base[8] = pos;
    }
    __syncthreads();
    long p[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = base[c] + wsum[c][warp] + off[c];
    for (long i = lo; i < hi; ++i) {
        const int par = par_smem ? (int)sort_par[i] : parity(i);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (par == c) perm[p[c]++] = i;
    }
    // padding positions: behind the real images of every class, and behind the last class
    for (int c = 0; c < 8; ++c) {
        const long first = meta[c] + meta[9 + c], last = c < 7 ? meta[c + 1] : base[8];
        for (long q = first + tid; q < last; q += SORT_THREADS) perm[q] = -1;
    }
    for (long q = base[8] + tid; q < perm_cap; q += SORT_THREADS) perm[q] = -1;
}

struct OrgPrepArgs {
    long n_images, n_pad_cap;   // Y row stride = n_pad_cap
    int Lmax;
    const long *perm, *meta;    // nullptr perm = identity (direct kernel)
    const float *R_sI_r;        // [I,3] (shoebox) or [3,I] (ARG, component_major = 1)
    int component_major;
    int mode, angdep;
    const int32_t *A;
    const cplx *Z;
    long K;
    RoomGeom room;
    double *Yp;                 // factorised path: real rows [tile][R_tot][ORG_OS] (pre-zeroed)
    int R_tot;
    const int *flags;           // flags[1]: uniform k grid, *(double *)(flags + 2) = dk (or nullptr)
};

constexpr int ORG_PREP_Y = 8;
// Y_l^{|mu|} for all l >= |mu| in one upward pass of the Legendre recurrence per |mu| (the same
// recurrence, normalisation and operation order as sph_harm_dev, pb:40-83, which restarts it for
// every (l, mu)); columns |mu| = am0, am0 + ORG_PREP_Y, ... of image `pos`.
__device__ __forceinline__ void org_harmonic_rows(const OrgPrepArgs &a, long pos, double phi, double ct, int am0) {
    const long tile = pos / ORG_TI;
    const int it = (int)(pos % ORG_TI);
    const double somx2 = sqrt(fma(-ct, ct, 1.0));
    for (int am = am0; am < a.Lmax; am += ORG_PREP_Y) {      // rows a_0, a_1, b_1, a_2, b_2, ...
        double pmm = 1.0, fact = 1.0;
        for (int i = 0; i < am; ++i) { pmm *= -fact * somx2; fact += 2.0; }
        double sn, cs;
        sincos_cw((double)am * phi, &sn, &cs);
        double p0 = pmm, p1 = ct * (double)(2 * am + 1) * pmm;     // P_am^am, P_{am+1}^am
        for (int l = am; l < a.Lmax; ++l) {
            double p_val;
            if (l == am) p_val = p0;
            else if (l == am + 1) p_val = p1;
            else {
                p_val = (ct * (double)(2 * l - 1) * p1 - (double)(l + am - 1) * p0) / (double)(l - am);
                p0 = p1; p1 = p_val;
            }
            const double np_ = g_org_norm[l * (l + 1) / 2 + am] * p_val;
            double *p = a.Yp + (tile * a.R_tot + c_org_rowoff[l] + (am ? 2 * am - 1 : 0)) * ORG_OS + it;
            p[0] = np_ * cs;
            if (am) p[ORG_OS] = np_ * sn;
        }
    }
}

// Block = 32 images (x) by ORG_PREP_Y values of |mu| (y): the per-image record is written by the y = 0
// thread, the real-harmonic rows of one |mu| column by thread (image, |mu| mod ORG_PREP_Y) — a
// one-thread-per-image version left 4 warps per SM walking 231 dependent values each (0.29 ms at C3
// size).  Every value is produced by the same operations in the same order as before.
__global__ void __launch_bounds__(32 * ORG_PREP_Y)
org_prep_kernel(OrgPrepArgs a, OrgImage *rec_out, cplx *Y) {
    long pos = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = a.perm ? a.meta[8] : a.n_images;
    if (pos >= total) return;
    long img = a.perm ? a.perm[pos] : pos;
    if (threadIdx.y != 0) {
        // harmonic rows only (factorised path); padding rows stay zero
        if (!a.Yp || img < 0) return;
        double phi, theta;
        if (a.component_major) { phi = (double)a.R_sI_r[img]; theta = (double)a.R_sI_r[a.n_images + img]; }
        else { phi = (double)a.R_sI_r[img * 3 + 0]; theta = (double)a.R_sI_r[img * 3 + 1]; }
        double st, ct;
        sincos_cw(theta, &st, &ct);
        org_harmonic_rows(a, pos, phi, ct, (int)threadIdx.y);
        return;
    }
    OrgImage rec;
    rec.orig = img; rec.pad = 0; rec.par = 0;
    rec.cx = rec.cy = rec.cz = 1.0;
    for (int i = 0; i < 8; ++i) rec.e[i] = 0;
    rec.atten_flat = cmake(1.0, 0.0);
    rec.rot = cmake(1.0, 0.0);
    if (img < 0) {   // padding: contributes exactly zero
        rec.r = 1.0; rec.inv_r = 1.0; rec.atten_flat = cmake(0.0, 0.0);
        rec_out[pos] = rec;
        if (Y)
            for (int s = 0; s < a.Lmax * a.Lmax; ++s) Y[(long)s * a.n_pad_cap + pos] = cmake(0.0, 0.0);
        return;
    }
    double phi, theta;
    if (a.component_major) {
        phi = (double)a.R_sI_r[img]; theta = (double)a.R_sI_r[a.n_images + img];
        rec.r = (double)a.R_sI_r[2 * a.n_images + img];
    } else {
        phi = (double)a.R_sI_r[img * 3 + 0]; theta = (double)a.R_sI_r[img * 3 + 1];
        rec.r = (double)a.R_sI_r[img * 3 + 2];
    }
    rec.inv_r = 1.0 / rec.r;
    if (a.flags) {
        double sn, cs;
        sincos_cw(*reinterpret_cast<const double *>(a.flags + 2) * rec.r, &sn, &cs);
        rec.rot = cmake(cs, -sn);
    }
    if (a.A) {
        const int32_t *A6 = a.A + img * 6;
        rec.par = ((A6[3] & 1) << 2) | ((A6[4] & 1) << 1) | (A6[5] & 1);
    }
    if (a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES) {
        AttenGeom g;
        if (a.mode == DEISM_ATTEN_FUSED_ROOM)
            atten_geom_room(a.A + img * 6, a.room.LL, a.room.xs, a.room.xr, a.angdep, g);
        else
            atten_geom_angles(a.A + img * 6, a.R_sI_r + img * 3, a.angdep, g);
        cplx at = shoebox_atten(a.Z, a.K, g.cx, g.cy, g.cz, g.e);
        rec.atten_flat = a.mode == DEISM_ATTEN_FUSED_ROOM ? round_c64(at) : at;
        rec.cx = g.cx; rec.cy = g.cy; rec.cz = g.cz;
        for (int i = 0; i < 8; ++i) rec.e[i] = g.e[i];
    }
    rec_out[pos] = rec;
    double st, ct;
    sincos_cw(theta, &st, &ct);
    if (a.Yp) {
        org_harmonic_rows(a, pos, phi, ct, 0);
        return;
    }
    for (int l = 0; l < a.Lmax; ++l)
        for (int mu = -l; mu <= l; ++mu)
            Y[(long)(l * l + l + mu) * a.n_pad_cap + pos] = sph_harm_dev(mu, l, phi, ct);
}

// ---------------------------------------------------------------------------------------
// consume kernel (factorised form) on the FP64 tensor cores.
//
// Per degree l: T_l[img,k] = sum_mu Y_l^mu(img) * B[par][l,mu][k], in the real-harmonic row basis a
// real [64 img x rows(l)] by complex [rows(l) x 32 freq] product = two real DMMA m8n8k4 GEMMs;
// then P[k] += (atten h_l)(k r_img) * T_l[img,k] with the h_l recurrence of every pair kept in
// registers (pb:86-118, j and y carried as one complex number) and the attenuation of the pair
// folded into the recurrence's two start values (it is linear).  (l, mu) rows are streamed in
// chunks of whole degrees through a ring of TMA bulk copies.
//
// CTA = 64 images x 32 frequencies, 4 x 4 warps of 16 x 8 (a thread owns four (image, frequency)
// pairs: its recurrence state is 20 doubles), one CTA per SM.  Against the 32 x 32 tile of round 1
// every B row is shared by four image groups and every Y row by four frequency groups: 35 % less
// L2 -> SM operand traffic per DMMA and half as many stage hand-overs.
// ---------------------------------------------------------------------------------------
struct OrgConsumeArgs {
    long K, n_ftiles;
    int Lmax, R_tot, ncpt;    // chunks per tile
    int rc;                   // rows a stage holds
    int nst;                  // ring stages
    const double *Bp;         // [8][n_ftiles][R_tot][2][ORG_BS]
    const double *Yp;         // [tiles][R_tot][ORG_OS]
    const OrgImage *rec;      // [n_pad]
    const long *meta;         // class starts + total
    const double *k;
    int mode;
    const void *atten;
    const cplx *Z;
    const int *flags;
    long tiles_per_chunk;     // uniform image-tile chunks ...
    cplx *partial;            // [n_chunks][K]
    int n_guided;             // ... or, when > 0, that many chunks of decreasing size (common.cuh)
    int chunk_lo[DEISM_MAXCHUNKS + 1];
};

// fixed part of the shared memory; the operand stages (rc rows of B then rc rows of Y each) follow
// it.  The final reduction over the row groups that share a frequency aliases the stages.
struct OrgSmem {
    OrgImage rec[2][ORG_TI];           // per-image records of the current / next tile (TMA)
    double2 Z[6][ORG_TF];
    double kf[ORG_TF], invk[ORG_TF], dkd[ORG_TF];      // dkd: deviation of one step from the uniform grid
    unsigned long long full[ORG_MAXST], recfull[2];
    int done[ORG_MAXST], recdone[2];
};
static_assert(sizeof(OrgSmem) % 16 == 0, "stages must stay 16-byte aligned");
static_assert(sizeof(OrgImage) % 16 == 0, "records are moved by bulk copies");

// D = A*B + C with C and D in DIFFERENT registers: the first k-step of a degree reads the finished
// sum two degrees up as its accumulator input and leaves it intact for the scalar step that still
// needs it (no copies, no cleared registers).
__device__ __forceinline__ void dmma884_c(double &d0, double &d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// The sum over degrees  sum_l h_l(kr) T_l  (h_l the spherical Hankel function, T_l = Y_l . B_l the
// degree's contraction) is taken by CLENSHAW's backward recurrence instead of by running h_l upwards
// and accumulating h_l T_l:  h_{l+1} = a_l h_l - h_{l-1}, a_l = (2l+1)/(kr), gives
//     b_l = T_l + a_l b_{l+1} - b_{l+2}  (l = L-1 .. 0),   sum = h_0 b_0 - h_{-1} b_1 = (e^{-ikr}/kr)(i b_0 - b_1).
// With signs s_l = + + - - + + ... folded into the harmonic rows (e_l = s_l b_l, the prep kernel's
// normalisation table carries s_l) the recurrence reads  e_l = s_l T_l + (+-a_l) e_{l+1} + e_{l+2}:
// e_{l+2} IS the DMMA accumulator input of degree l, so the tensor core adds it for free, and what
// is left per (image, frequency) pair and degree is one multiply (a_l) and two FMAs (Re, Im) — three
// FP64 instructions where the upward recurrence + complex accumulate took seven, and 16 registers
// less.  As accurate as the upward form over kr = 0.04 .. 2000 (checked against 60-digit sums).
// Degrees are whole inside a chunk and run DOWNWARDS; the two register sets ping-pong with the
// parity of l, so the scalar step that finishes e_{l+1} issues in the shadow of degree l's DMMAs.
// Operand chunks and the per-image records arrive by TMA; the last warp out of a buffer refills it:
// no CTA-wide barrier in the tile loop.
__global__ void __launch_bounds__(ORG_THREADS, ORG_CTAS_PER_SM)
org_consume_kernel(OrgConsumeArgs a) {
    constexpr int NWARPS = ORG_THREADS / 32;
    extern __shared__ __align__(128) unsigned char org_smem_raw[];
    OrgSmem &sm = *reinterpret_cast<OrgSmem *>(org_smem_raw);
    const unsigned stage_bytes = (unsigned)a.rc * (2 * ORG_BS + ORG_OS) * 8;
    unsigned char *stages = org_smem_raw + sizeof(OrgSmem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm = warp / ORG_WN, wn = warp % ORG_WN;   // warp tile: 16 images x 8 frequencies
    const long f0 = (long)blockIdx.x * ORG_TF;
    const long n_tiles_all = a.meta[8] / ORG_TI;
    long tile_lo = a.n_guided ? (long)a.chunk_lo[blockIdx.y] : (long)blockIdx.y * a.tiles_per_chunk;
    long tile_hi = a.n_guided ? (long)a.chunk_lo[blockIdx.y + 1] : tile_lo + a.tiles_per_chunk;
    if (tile_hi > n_tiles_all) tile_hi = n_tiles_all;
    const int n_tiles = tile_hi > tile_lo ? (int)(tile_hi - tile_lo) : 0;
    const int ncpt = a.ncpt, nst = a.nst;
    const int total = n_tiles * ncpt;

    const bool fused = a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES;
    const bool flat = fused && a.flags[0] != 0;

    if (tid == 0) {
#ifdef ORG_LDGSTS
        for (int b = 0; b < nst; ++b) { mbar_init(&sm.full[b], 32); sm.done[b] = 0; }
#else
        for (int b = 0; b < nst; ++b) { mbar_init(&sm.full[b], 1); sm.done[b] = 0; }
#endif
        for (int b = 0; b < 2; ++b) { mbar_init(&sm.recfull[b], 1); sm.recdone[b] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool uniform_k = a.flags[1] != 0;
    {
        const double dk = *reinterpret_cast<const double *>(a.flags + 2);
        for (int q = tid; q < ORG_TF; q += ORG_THREADS) {
            long f = f0 + q;
            double kk = f < a.K ? a.k[f] : 1.0;
            sm.kf[q] = kk;
            sm.invk[q] = 1.0 / kk;
            sm.dkd[q] = f + 1 < a.K ? (a.k[f + 1] - kk) - dk : 0.0;
        }
    }
    if (fused && !flat) {
        for (int i = tid; i < 6 * ORG_TF; i += ORG_THREADS) {
            int w = i / ORG_TF, q = i % ORG_TF;
            long f = f0 + q;
            sm.Z[w][q] = f < a.K ? a.Z[(long)w * a.K + f] : cmake(1.0, 0.0);
        }
    }
    __syncthreads();

    auto issue = [&](int tile_ord, int local, int stage) {
        const long tile = tile_lo + tile_ord;
        const int4 ch = c_org_chunks[local];
        const long p0 = tile * ORG_TI;
        int par = 0;
#pragma unroll
        for (int q = 1; q < 8; ++q) par += (p0 >= a.meta[q]);
        const double *Bg = a.Bp + (((long)par * a.n_ftiles + blockIdx.x) * a.R_tot + ch.x) * 2 * ORG_BS;
        const double *Yg = a.Yp + (tile * a.R_tot + ch.x) * ORG_OS;
        const unsigned bb = (unsigned)(ch.y * 2 * ORG_BS * 8), yb = (unsigned)(ch.y * ORG_OS * 8);
        unsigned char *S = stages + (unsigned)stage * stage_bytes;
        mbar_expect_tx(&sm.full[stage], bb + yb);
#ifndef ORG_TMA_SPLIT
        bulk_g2s(S, Bg, bb, &sm.full[stage]);
        bulk_g2s(S + (unsigned)a.rc * 2 * ORG_BS * 8, Yg, yb, &sm.full[stage]);
#else
        // several copies of whole rows in flight per stage instead of two large ones
        for (int r0 = 0; r0 < ch.y; r0 += ORG_TMA_SPLIT) {
            const int nr = ch.y - r0 < ORG_TMA_SPLIT ? ch.y - r0 : ORG_TMA_SPLIT;
            bulk_g2s(S + (unsigned)r0 * 2 * ORG_BS * 8, Bg + (long)r0 * 2 * ORG_BS, (unsigned)(nr * 2 * ORG_BS * 8),
                     &sm.full[stage]);
            bulk_g2s(S + (unsigned)a.rc * 2 * ORG_BS * 8 + (unsigned)r0 * ORG_OS * 8, Yg + (long)r0 * ORG_OS,
                     (unsigned)(nr * ORG_OS * 8), &sm.full[stage]);
        }
#endif
    };
#ifdef ORG_LDGSTS
    // The same refill done by a whole warp with 16-byte cp.async (LDGSTS): the bulk-copy engine takes
    // one stage at a time (>= ~550 clk per stage however small, scripts/probes/tma_probe.cu), plain
    // asynchronous copies spread over the L2 slices and pipeline freely.
    auto issue_warp = [&](int tile_ord, int local, int stage) {
        const long tile = tile_ord + tile_lo;
        const int4 ch = c_org_chunks[local];
        const long p0 = tile * ORG_TI;
        int par = 0;
#pragma unroll
        for (int q = 1; q < 8; ++q) par += (p0 >= a.meta[q]);
        const unsigned char *Bg = reinterpret_cast<const unsigned char *>(
            a.Bp + (((long)par * a.n_ftiles + blockIdx.x) * a.R_tot + ch.x) * 2 * ORG_BS);
        const unsigned char *Yg = reinterpret_cast<const unsigned char *>(a.Yp + (tile * a.R_tot + ch.x) * ORG_OS);
        const unsigned bb = (unsigned)(ch.y * 2 * ORG_BS * 8), yb = (unsigned)(ch.y * ORG_OS * 8);
        unsigned char *S = stages + (unsigned)stage * stage_bytes;
        unsigned char *SY = S + (unsigned)a.rc * 2 * ORG_BS * 8;
        for (unsigned off = lane * 16; off < bb; off += 512) cp_async16(S + off, Bg + off);
        for (unsigned off = lane * 16; off < yb; off += 512) cp_async16(SY + off, Yg + off);
        cp_async_mbar_arrive(&sm.full[stage]);
    };
#endif
    auto issue_records = [&](int tile_ord) {
        const int rb = tile_ord & 1;
        mbar_expect_tx(&sm.recfull[rb], (unsigned)(ORG_TI * sizeof(OrgImage)));
        bulk_g2s(&sm.rec[rb][0], a.rec + (tile_lo + tile_ord) * ORG_TI, (unsigned)(ORG_TI * sizeof(OrgImage)),
                 &sm.recfull[rb]);
    };
    if (tid == 0 && total > 0) {
        issue_records(0);
        if (n_tiles > 1) issue_records(1);
#ifndef ORG_LDGSTS
        int l0 = 0, t0 = 0;
        for (int s = 0; s < nst && s < total; ++s) {
            issue(t0, l0, s);
            if (++l0 == ncpt) { l0 = 0; ++t0; }
        }
#endif
    }
#ifdef ORG_LDGSTS
    if (warp == 0 && total > 0) {
        int l0 = 0, t0 = 0;
        for (int s = 0; s < nst && s < total; ++s) {
            issue_warp(t0, l0, s);
            if (++l0 == ncpt) { l0 = 0; ++t0; }
        }
    }
#endif
    // (pt, pl): the chunk nst ahead of the current one
    int pt = nst / ncpt, pl = nst % ncpt;
    int c = 0, stage = 0;
    unsigned phase = 0;

    // E[parity of l][Re|Im][m-tile][e]: e_l once finished, before that degree l's accumulator
    double E[2][2][2][2];
    cplx A[2][2];               // atten * exp(-i k r) / (k r) of the pair
    double xinv[2][2];          // 1 / (k r) of the pair
    cplx P[2] = {cmake(0.0, 0.0), cmake(0.0, 0.0)};      // [e], summed over this thread's images

    // one degree l of parity PAR = l & 1 (compile-time roles of the two register sets): on entry
    // E[PAR] = e_{l+2} (finished) and E[PAR^1] = degree l+1's accumulator, its DMMAs possibly still
    // in flight.  Row block of degree l: a_0, a_1, b_1, ..., a_l, b_l (2l + 1 real rows).  For ODD l
    // that is 3 mod 4: one zero pad row completes the last DMMA k-step.  For EVEN l the 2l rows
    // behind a_0 are whole k-steps, and the a_0 row is a rank-one update done in plain FMAs — 8 DFMA
    // instead of a mostly empty k-step of 4 DMMAs.
    auto degree = [&](auto par_tag, const int l, const double *Bs, const double *Ys, const int row0) {
        constexpr int PAR = decltype(par_tag)::value;
        const int row = c_org_rowoff[l] - row0 + (PAR == 0 ? 1 : 0);      // first row of the k-steps
        const int ksn = PAR == 0 ? (l >> 1) : ((l + 1) >> 1);              // k-steps of degree l
        const double *Yw = Ys + (row + t4) * ORG_OS + wm * 16 + g;
        const double *Bw = Bs + (row + t4) * 2 * ORG_BS + wn * 8 + g;
        auto load_frag = [&](int ks, double (&af)[2], double (&bf)[2]) {
            af[0] = Yw[ks * 4 * ORG_OS];
            af[1] = Yw[ks * 4 * ORG_OS + 8];
            bf[0] = Bw[ks * 8 * ORG_BS];
            bf[1] = Bw[ks * 8 * ORG_BS + ORG_BS];
        };
        double acc[2][2][2];   // [Re|Im][m-tile][e]
        auto mma_frag = [&](const double (&af)[2], const double (&bf)[2]) {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma884(acc[p][mt][0], acc[p][mt][1], af[mt], bf[p]);
        };
        double af[2], bf[2];
        int ks = 0;            // k-steps issued so far
        if (PAR == 0) {
            // a_0 row: this thread's two images (rows g, g + 8 of the D fragment) and two frequencies
            const double *y0 = Ys + (row - 1) * ORG_OS + wm * 16 + g;
            const double *b0 = Bs + (row - 1) * 2 * ORG_BS + wn * 8 + 2 * t4;
            const double ya[2] = {y0[0], y0[8]};
            const double2 bre = *reinterpret_cast<const double2 *>(b0);
            const double2 bim = *reinterpret_cast<const double2 *>(b0 + ORG_BS);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                acc[0][mt][0] = fma(ya[mt], bre.x, E[0][0][mt][0]); acc[0][mt][1] = fma(ya[mt], bre.y, E[0][0][mt][1]);
                acc[1][mt][0] = fma(ya[mt], bim.x, E[0][1][mt][0]); acc[1][mt][1] = fma(ya[mt], bim.y, E[0][1][mt][1]);
            }
            if (ksn > 0) {
                load_frag(0, af, bf);
                mma_frag(af, bf);
                ks = 1;
            }
        } else {
            load_frag(0, af, bf);
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    dmma884_c(acc[p][mt][0], acc[p][mt][1], af[mt], bf[p], E[1][p][mt][0], E[1][p][mt][1]);
            ks = 1;
        }
        // finish degree l+1 in the shadow of the k-step above: e_{l+1} += (+-a_{l+1}) e_{l+2}, the sign
        // + for even l+1 (s_{l+1} s_{l+2})
        {
            const double cl = PAR == 1 ? (double)(2 * l + 3) : -(double)(2 * l + 3);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double cf = cl * xinv[mt][e];
                    E[PAR ^ 1][0][mt][e] = fma(cf, E[PAR][0][mt][e], E[PAR ^ 1][0][mt][e]);
                    E[PAR ^ 1][1][mt][e] = fma(cf, E[PAR][1][mt][e], E[PAR ^ 1][1][mt][e]);
                }
        }
        // remaining k-steps, two per trip with ping-pong fragment registers: the loads of one step
        // are in flight while the other multiplies
        if (ks < ksn) {
            double a2[2], b2[2];
            load_frag(ks, af, bf);
            ++ks;
#pragma unroll 1
            for (; ks + 1 < ksn; ks += 2) {
                load_frag(ks, a2, b2);
                mma_frag(af, bf);
                load_frag(ks + 1, af, bf);
                mma_frag(a2, b2);
            }
            if (ks < ksn) {
                load_frag(ks, a2, b2);
                mma_frag(af, bf);
                mma_frag(a2, b2);
            } else {
                mma_frag(af, bf);
            }
        }
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { E[PAR][p][mt][0] = acc[p][mt][0]; E[PAR][p][mt][1] = acc[p][mt][1]; }
    };
    using Even = std::integral_constant<int, 0>;
    using Odd = std::integral_constant<int, 1>;

#pragma unroll 1
    for (int tt = 0; tt < n_tiles; ++tt) {
        const int rb = tt & 1;
        mbar_wait(&sm.recfull[rb], (unsigned)((tt >> 1) & 1));
        // A = atten * h_{-1}(kr) = atten * exp(-i k r) / (k r) of this thread's DMMA-mapped pairs
        // (pb:86-118, h = j - i y; h_0 = i h_{-1}), and the empty sums e_L = e_{L+1} = 0
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const OrgImage &rec = sm.rec[rb][wm * 16 + mt * 8 + g];
            const double r = rec.r, ir = rec.inv_r;
            double cs0 = 1.0, sn0 = 0.0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int fq = wn * 8 + 2 * t4 + e;
                cplx att;
                if (rec.orig < 0) {
                    att = cmake(0.0, 0.0);
                } else if (flat) {
                    att = rec.atten_flat;
                } else if (fused) {
                    att = shoebox_atten_term(&sm.Z[0][fq], ORG_TF, rec.cx, rec.cy, rec.cz, rec.e);
                    if (a.mode == DEISM_ATTEN_FUSED_ROOM) att = round_c64(att);
                } else {
                    const long f = f0 + fq;
                    att = cmake(0.0, 0.0);
                    if (f < a.K) {
                        if (a.mode == DEISM_ATTEN_C64) {
                            float2 v = ((const float2 *)a.atten)[rec.orig * a.K + f];
                            att = cmake((double)v.x, (double)v.y);
                        } else {
                            att = ((const cplx *)a.atten)[rec.orig * a.K + f];
                        }
                    }
                }
                // exp(-i k r) = (cs, -sn): on a uniform grid the odd frequency of the thread's pair is
                // one rotation by exp(-i dk r) (and the first-order term of the grid's own rounding)
                // away from the even one (Q8, as in lc.cu)
                double sn, cs;
                if (e == 0 || !uniform_k) {
                    sincos_cw(sm.kf[fq] * r, &sn, &cs);
                } else {
                    const double dr = sm.dkd[fq - 1] * r;
                    const cplx F = cmul(cmake(cs0, -sn0), rec.rot);
                    cs = fma(F.y, dr, F.x);
                    sn = -fma(-F.x, dr, F.y);
                }
                cs0 = cs; sn0 = sn;
                const double i1 = sm.invk[fq] * ir;
                xinv[mt][e] = i1;
                A[mt][e] = cmul(att, cmake(cs * i1, -sn * i1));
#pragma unroll
                for (int par = 0; par < 2; ++par) { E[par][0][mt][e] = 0.0; E[par][1][mt][e] = 0.0; }
            }
        }
        __syncwarp();
        if (lane == 0 && atomicAdd(&sm.recdone[rb], 1) == NWARPS - 1) {
            // last warp out of this record buffer: refill it with the records of tile tt+2
            sm.recdone[rb] = 0;
            if (tt + 2 < n_tiles) issue_records(tt + 2);
        }
#pragma unroll 1
        for (int local = 0; local < ncpt; ++local) {
            const int4 ch = c_org_chunks[local];          // chunks and degrees run downwards
            mbar_wait(&sm.full[stage], phase);
            const double *Bs = reinterpret_cast<const double *>(stages + (unsigned)stage * stage_bytes);
            const double *Ys = Bs + a.rc * 2 * ORG_BS;
            int l = ch.w - 1;
            if (!(l & 1)) { degree(Even{}, l, Bs, Ys, ch.x); --l; }
#pragma unroll 1
            for (; l - 1 >= ch.z; l -= 2) {
                degree(Odd{}, l, Bs, Ys, ch.x);
                degree(Even{}, l - 1, Bs, Ys, ch.x);
            }
            if (l >= ch.z) degree(Odd{}, l, Bs, Ys, ch.x);
            __syncwarp();
#ifdef ORG_LDGSTS
            {
                int last = 0;
                if (lane == 0) {
                    last = atomicAdd(&sm.done[stage], 1) == NWARPS - 1;
                    if (last) sm.done[stage] = 0;
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last && c + nst < total) issue_warp(pt, pl, stage);
            }
#else
            if (lane == 0 && atomicAdd(&sm.done[stage], 1) == NWARPS - 1) {
                // last warp out of this stage: refill it with the chunk nst ahead
                sm.done[stage] = 0;
                if (c + nst < total) issue(pt, pl, stage);
            }
#endif
            ++c;
            if (++stage == nst) { stage = 0; phase ^= 1u; }
            if (++pl == ncpt) { pl = 0; ++pt; }
        }
        // degree 0 was the last one issued: E[1] = e_1 (finished in its shadow), E[0] its accumulator.
        // e_0 = acc_0 + a_0 e_1;  P += A (i e_0 - e_1)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double e0r = fma(xinv[mt][e], E[1][0][mt][e], E[0][0][mt][e]);
                const double e0i = fma(xinv[mt][e], E[1][1][mt][e], E[0][1][mt][e]);
                cfma(P[e], A[mt][e], cmake(-e0i - E[1][0][mt][e], e0r - E[1][1][mt][e]));
            }
    }

    // ---- reduce over the 8 row groups x ORG_WM image warps that share a frequency (aliases the stages) ----
    __syncthreads();
    double2 (*red)[ORG_TF] = reinterpret_cast<double2 (*)[ORG_TF]>(stages);
#pragma unroll
    for (int e = 0; e < 2; ++e) red[wm * 8 + g][wn * 8 + 2 * t4 + e] = P[e];
    __syncthreads();
    if (tid < ORG_TF) {
        cplx s = cmake(0.0, 0.0);
#pragma unroll 8
        for (int y = 0; y < 8 * ORG_WM; ++y) s = cadd(s, red[y][tid]);
        long f = f0 + tid;
        if (f < a.K) a.partial[(long)blockIdx.y * a.K + f] = s;
    }
}

// ---------------------------------------------------------------------------------------
// The same factorised sum for a FEW images (the early images of DEISM-MIX: 25 at mixEarlyOrder 2):
// grouped by parity class and padded to tiles of 64 they would be eight tiles of mostly padding
// (0.25 ms at C5 size for 25 images).  Here a thread owns one frequency and walks the images of its
// lane: T_l = sum_mu Y_l^mu(img) B[par][l,mu][k] in plain FMAs over the 2l+1 real rows (no DMMA
// k-step padding either), same recurrence, same operand tables.  Warp = 32 frequencies of one
// image: the Y value is a broadcast, the B row a coalesced 256-byte read.
// ---------------------------------------------------------------------------------------
constexpr int ORG_SMALL_LANES = 4;        // images in flight per CTA (one warp each)
constexpr int ORG_SMALL_MAX_IMAGES = 512;
struct OrgSmallArgs {
    long K, n_ftiles, n_images;
    int Lmax, R_tot;
    const double *Bp, *Yp;
    const OrgImage *rec;
    const long *meta;
    const double *k;
    int mode;
    const void *atten;
    const cplx *Z;
    const int *flags;
    cplx *partial;
};

__global__ void __launch_bounds__(32 * ORG_SMALL_LANES) org_small_kernel(OrgSmallArgs a) {
    __shared__ double2 red[ORG_SMALL_LANES][32];
    const int q = threadIdx.x & 31, j = threadIdx.x >> 5;
    const long ft = blockIdx.x, f = ft * ORG_TF + q;
    const bool live_f = f < a.K;
    const double kk = live_f ? a.k[f] : 1.0, ik = 1.0 / kk;
    const bool fused = a.mode == DEISM_ATTEN_FUSED_ROOM || a.mode == DEISM_ATTEN_FUSED_ANGLES;
    const bool flat = fused && a.flags[0] != 0;
    cplx P = cmake(0.0, 0.0);
    // image i of the class-grouped order: class c holds meta[9 + c] real images from position meta[c]
    for (long i = (long)blockIdx.y * ORG_SMALL_LANES + j; i < a.n_images; i += (long)gridDim.y * ORG_SMALL_LANES) {
        int par = 0;
        long before = 0;
        for (int c = 0; c < 7; ++c) {
            const long nc = a.meta[9 + c];
            if (i >= before + nc) { before += nc; par = c + 1; }
            else break;
        }
        const long pos = a.meta[par] + (i - before);
        const OrgImage rec = a.rec[pos];
        cplx att;
        if (flat) {
            att = rec.atten_flat;
        } else if (fused) {
            att = shoebox_atten(a.Z + (live_f ? f : 0), a.K, rec.cx, rec.cy, rec.cz, rec.e);
            if (a.mode == DEISM_ATTEN_FUSED_ROOM) att = round_c64(att);
        } else if (!live_f) {
            att = cmake(0.0, 0.0);
        } else if (a.mode == DEISM_ATTEN_C64) {
            const float2 v = ((const float2 *)a.atten)[rec.orig * a.K + f];
            att = cmake((double)v.x, (double)v.y);
        } else {
            att = ((const cplx *)a.atten)[rec.orig * a.K + f];
        }
        double sn, cs;
        sincos_cw(kk * rec.r, &sn, &cs);
        const double i1 = ik * rec.inv_r;
        const double ci = cs * i1, si = sn * i1;
        cplx hc = cmul(att, cmake(ci, -si));                                   // atten * h_{-1}
        cplx hm = cmul(att, cmake(fma(-ci, i1, -si), fma(si, i1, -ci)));       // atten * h_{-2}
        const double *Yrow = a.Yp + (pos / ORG_TI) * (long)a.R_tot * ORG_OS + pos % ORG_TI;
        const double *Brow = a.Bp + (((long)par * a.n_ftiles + ft) * a.R_tot) * 2 * ORG_BS + q;
        for (int l = 0; l < a.Lmax; ++l) {
            const double cf = (double)(2 * l - 1) * i1;
            const cplx h = cmake(fma(cf, hc.x, -hm.x), fma(cf, hc.y, -hm.y));
            hm = hc; hc = h;
            const int r0 = c_org_rowoff[l];
            double tx = 0.0, ty = 0.0, ux = 0.0, uy = 0.0;       // two partial sums: twice the loads in flight
            int r = r0;
#pragma unroll 4
            for (; r + 1 < r0 + 2 * l + 1; r += 2) {
                const double y0 = __ldg(Yrow + (long)r * ORG_OS), y1 = __ldg(Yrow + (long)(r + 1) * ORG_OS);
                tx = fma(y0, __ldg(Brow + (long)r * 2 * ORG_BS), tx);
                ty = fma(y0, __ldg(Brow + (long)r * 2 * ORG_BS + ORG_BS), ty);
                ux = fma(y1, __ldg(Brow + (long)(r + 1) * 2 * ORG_BS), ux);
                uy = fma(y1, __ldg(Brow + (long)(r + 1) * 2 * ORG_BS + ORG_BS), uy);
            }
            {   // 2l+1 is odd: the last row
                const double y0 = __ldg(Yrow + (long)r * ORG_OS);
                tx = fma(y0, __ldg(Brow + (long)r * 2 * ORG_BS), tx);
                ty = fma(y0, __ldg(Brow + (long)r * 2 * ORG_BS + ORG_BS), ty);
            }
            const double sg = (l & 2) ? -1.0 : 1.0;      // the rows carry the Clenshaw sign s_l
            cfma(P, h, cmake(sg * (tx + ux), sg * (ty + uy)));
        }
    }
    red[j][q] = P;
    __syncthreads();
    if (j == 0 && live_f) {
        cplx s_ = red[0][q];
#pragma unroll
        for (int t = 1; t < ORG_SMALL_LANES; ++t) s_ = cadd(s_, red[t][q]);
        a.partial[(long)blockIdx.y * a.K + f] = s_;
    }
}

// ---------------------------------------------------------------------------------------
// direct kernel: thread = image, blockIdx.y = frequency.  Source coefficients are addressed
// as Cs[k*sk + nm*snm + img*si] (complex64): shoebox [K][NM] (si = 0) or ARG [K][NM][I].
// ---------------------------------------------------------------------------------------
struct OrgDirectArgs {
    long n_sel, n_images, K;
    int Lmax, nseg, NMr;
    const OrgEntry *ent[2];
    const int *seg[2];
    const float2 *Cs;
    long sk, snm, si;
    const float2 *Cr;          // [K][NMr]
    const cplx *Y;             // [nseg][n_images]
    const OrgImage *rec;       // [n_images]
    const long *sel;           // [n_sel] or nullptr
    const double *k;
    int mode;                  // DEISM_ATTEN_* (shoebox) or -1: ARG atten complex64 [K][I]
    const void *atten;
    const cplx *Z;
    int shoebox;
    cplx *partial;             // [gridDim.x][K]
};

__global__ void __launch_bounds__(128) org_direct_kernel(OrgDirectArgs a) {
    __shared__ double2 red[128];
    const long s = (long)blockIdx.x * 128 + threadIdx.x;
    // frequencies beyond the grid's y extent (65535) are taken in further passes
    for (long kidx = blockIdx.y; kidx < a.K; kidx += gridDim.y) {
    cplx P = cmake(0.0, 0.0);
    if (s < a.n_sel) {
        const long img = a.sel ? a.sel[s] : s;
        const OrgImage rec = a.rec[img];
        const double kk = a.k[kidx];
        int cls = 0, b = 0, c = 0;
        if (a.shoebox) {
            int px = rec.par >> 2, py = (rec.par >> 1) & 1, pz = rec.par & 1;
            cls = (px + py) & 1; b = (py + pz) & 1; c = pz;
        }
        const OrgEntry *ent = cls ? a.ent[1] : a.ent[0];
        const int *seg = cls ? a.seg[1] : a.seg[0];
        const float2 *Cs = a.Cs + kidx * a.sk + img * a.si;
        const float2 *Cr = a.Cr + kidx * a.NMr;
        double sn, cs;
        sincos_cw(kk * rec.r, &sn, &cs);
        const double ix = 1.0 / (kk * rec.r);
        cplx hm = cmake(sn * ix, cs * ix);
        cplx hc = cmake(fma(sn * ix, ix, -cs * ix), fma(cs * ix, ix, sn * ix));
        cplx acc = cmake(0.0, 0.0);
        for (int l = 0; l < a.Lmax; ++l) {
            cplx h;
            if (l == 0) h = hm;
            else if (l == 1) h = hc;
            else {
                double cf = (double)(2 * l - 1) * ix;
                h = cmake(fma(cf, hc.x, -hm.x), fma(cf, hc.y, -hm.y));
                hm = hc; hc = h;
            }
            for (int mu = -l; mu <= l; ++mu) {
                int sg = l * l + l + mu;
                int lo = seg[sg], hi = seg[sg + 1];
                if (lo == hi) continue;
                cplx bsum = cmake(0.0, 0.0);
                for (int j = lo; j < hi; ++j) {
                    OrgEntryReg e = load_entry(ent + j);
                    double g = ((b * e.m + c * e.n) & 1) ? -e.G : e.G;
                    float2 c1 = Cs[(long)e.idx_s * a.snm], c2 = Cr[e.idx_r];
                    cplx d = cmul(cmake((double)c1.x, (double)c1.y), cmake((double)c2.x, (double)c2.y));
                    bsum.x = fma(g, d.x, bsum.x);
                    bsum.y = fma(g, d.y, bsum.y);
                }
                cfma(acc, cmul(h, a.Y[(long)sg * a.n_images + img]), bsum);
            }
        }
        // * (i/k) * atten
        const double ik = 1.0 / kk;
        cplx v = cmake(-acc.y * ik, acc.x * ik);
        cplx att;
        if (a.mode == -1) {
            float2 t = ((const float2 *)a.atten)[kidx * a.n_images + img];
            att = cmake((double)t.x, (double)t.y);
        } else if (a.mode == DEISM_ATTEN_C64) {
            float2 t = ((const float2 *)a.atten)[img * a.K + kidx];
            att = cmake((double)t.x, (double)t.y);
        } else if (a.mode == DEISM_ATTEN_C128) {
            att = ((const cplx *)a.atten)[img * a.K + kidx];
        } else {
            att = shoebox_atten(a.Z + kidx, a.K, rec.cx, rec.cy, rec.cz, rec.e);
            if (a.mode == DEISM_ATTEN_FUSED_ROOM) att = round_c64(att);
        }
        P = cmul(att, v);
    }
    red[threadIdx.x] = P;
    __syncthreads();
    for (int st = 64; st > 0; st >>= 1) {
        if (threadIdx.x < st) red[threadIdx.x] = cadd(red[threadIdx.x], red[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partial[(long)blockIdx.x * a.K + kidx] = red[0];
    __syncthreads();
    }
}

}  // namespace deism

// opaque to C callers (include/deism_b200.h)
struct deism_org_plan {
    int N, V, nseg, device;
    void *d_block;
    const deism::OrgEntry *d_ent[2];
    const int *d_seg[2];
    const deism::BtEntry *d_bt;     // B-table form (class-0 entries by segment and parity group)
    const int *d_bt_seg;
    const int2 *d_jobs;
    int n_jobs;
    long nnz;
};

namespace deism {

// A second stream per device for the image-independent half of deism_org_shoebox (coefficient
// transposes + B table), which runs next to the image-dependent half (class sort, records, harmonic
// rows) and joins it before the consume kernel.  DEISM_ORG_NO_FORK=1 keeps everything on one stream.
struct OrgSide {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
// Joins the side stream back into the caller's stream on EVERY way out of the function once the fork
// has happened (also the error returns): whatever the caller's stream does next — including the
// caller's allocator recycling the input buffers — is ordered after the side stream's reads.
struct OrgJoin {
    OrgSide *side = nullptr;
    cudaStream_t st = nullptr;
    bool recorded = false, joined = false;
    void record() {
        if (side && !recorded) { cudaEventRecord(side->join, side->s); recorded = true; }
    }
    void join() {
        if (side && !joined) { record(); cudaStreamWaitEvent(st, side->join, 0); joined = true; }
    }
    ~OrgJoin() { join(); }
};
static OrgSide *org_side() {
    static OrgSide tab[64];
    static std::mutex mu;
    static const bool off = [] { const char *e = getenv("DEISM_ORG_NO_FORK"); return e && e[0] == '1'; }();
    if (off) return nullptr;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    OrgSide &o = tab[dev];
    if (!o.s) {
        if (cudaStreamCreateWithFlags(&o.s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&o.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&o.join, cudaEventDisableTiming) != cudaSuccess) {
            o.s = nullptr;
            return nullptr;
        }
    }
    return &o;
}

static int plan_build(int N, int V, const float *h_W1, const float *h_W2, deism_org_plan *P) {
    const int Lmax = N + V + 1, nseg = Lmax * Lmax;
    OrgLists L;
    build_lists(N, V, h_W1, h_W2, L);
    auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t e0_b = up16(L.entries[0].size() * sizeof(OrgEntry)), e1_b = up16(L.entries[1].size() * sizeof(OrgEntry));
    const size_t bt_b = up16(L.bt.size() * sizeof(BtEntry));
    const size_t seg_b = up16((size_t)(nseg + 1) * sizeof(int)), bts_b = up16(L.bt_seg.size() * sizeof(int));
    const size_t job_b = up16(L.jobs.size() * sizeof(int2));
    unsigned char *base = nullptr;
    if (cudaMalloc(&base, e0_b + e1_b + bt_b + 2 * seg_b + bts_b + job_b + 64) != cudaSuccess) {
        cudaGetLastError();
        return set_error(DEISM_ERR_NOMEM, "cudaMalloc of the coupling plan failed");
    }
    unsigned char *cur = base;
    auto put = [&](const void *src, size_t bytes, size_t padded) -> void * {
        void *dst = cur;
        cur += padded;
        if (bytes && cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return dst;
    };
    void *e0 = put(L.entries[0].data(), L.entries[0].size() * sizeof(OrgEntry), e0_b);
    void *e1 = put(L.entries[1].data(), L.entries[1].size() * sizeof(OrgEntry), e1_b);
    void *bt = put(L.bt.data(), L.bt.size() * sizeof(BtEntry), bt_b);
    void *s0 = put(L.seg[0].data(), (nseg + 1) * sizeof(int), seg_b);
    void *s1 = put(L.seg[1].data(), (nseg + 1) * sizeof(int), seg_b);
    void *bs = put(L.bt_seg.data(), L.bt_seg.size() * sizeof(int), bts_b);
    void *jb = put(L.jobs.data(), L.jobs.size() * sizeof(int2), job_b);
    if (!e0 || !e1 || !bt || !s0 || !s1 || !bs || !jb) {
        cudaError_t e = cudaGetLastError();
        cudaFree(base);
        return set_error(DEISM_ERR_CUDA, "coupling plan upload failed: %s", cudaGetErrorString(e));
    }
    P->N = N; P->V = V; P->nseg = nseg; P->d_block = base; P->nnz = L.nnz;
    P->d_ent[0] = (const OrgEntry *)e0; P->d_ent[1] = (const OrgEntry *)e1;
    P->d_seg[0] = (const int *)s0; P->d_seg[1] = (const int *)s1;
    P->d_bt = (const BtEntry *)bt; P->d_bt_seg = (const int *)bs; P->d_jobs = (const int2 *)jb;
    P->n_jobs = (int)L.jobs.size();
    cudaGetDevice(&P->device);
    return DEISM_OK;
}

// plan supplied by the caller, or a temporary one built from the tables (freed by the guard)
struct PlanGuard {
    deism_org_plan tmp = {};
    const deism_org_plan *use = nullptr;
    ~PlanGuard() { if (tmp.d_block) cudaFree(tmp.d_block); }   // cudaFree waits for the kernels
    int get(const deism_org_plan *plan, int N, int V, const float *h_W1, const float *h_W2) {
        if (plan) {
            if (plan->N != N || plan->V != V)
                return set_error(DEISM_ERR_INVALID, "coupling plan is for orders (%d,%d), call has (%d,%d)",
                                 plan->N, plan->V, N, V);
            use = plan;
            return DEISM_OK;
        }
        if (!h_W1 || !h_W2) return set_error(DEISM_ERR_INVALID, "need a plan or the Wigner tables");
        int rc = plan_build(N, V, h_W1, h_W2, &tmp);
        if (rc) return rc;
        use = &tmp;
        return DEISM_OK;
    }
};

static void fill_room(RoomGeom &room, const deism_shoebox_atten *at) {
    for (int i = 0; i < 3; ++i) {
        room.LL[i] = at->room_size[i]; room.xs[i] = at->pos_source[i]; room.xr[i] = at->pos_receiver[i];
    }
}

}  // namespace deism

using namespace deism;

extern "C" int deism_org_plan_create(int N, int V, const float *h_W1, const float *h_W2,
                                     deism_org_plan **plan) {
    int rc = check_device();
    if (rc) return rc;
    if (!plan || !h_W1 || !h_W2) return set_error(DEISM_ERR_INVALID, "null pointer");
    if (N < 0 || V < 0 || N > 20 || V > 20) return set_error(DEISM_ERR_INVALID, "orders must be 0..20");
    deism_org_plan *P = new deism_org_plan();
    rc = plan_build(N, V, h_W1, h_W2, P);
    if (rc) { delete P; return rc; }
    *plan = P;
    return DEISM_OK;
}

extern "C" int deism_org_plan_destroy(deism_org_plan *plan) {
    if (!plan) return DEISM_OK;
    if (plan->d_block) cudaFree(plan->d_block);
    delete plan;
    return DEISM_OK;
}

static int org_shoebox_impl(int64_t n_images, int64_t K, int N, int V, const void *d_C_nm_s,
                            const void *d_C_vu_r, int coef_c128, const float *h_W1, const float *h_W2,
                            const int32_t *d_A, const float *d_R_sI_r,
                            const deism_shoebox_atten *atten, const double *d_k, double *d_out,
                            int accumulate, int algo, const deism_org_plan *plan, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    CallGuard call_guard(stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 0) return DEISM_OK;
    if (!d_out || !d_k) return set_error(DEISM_ERR_INVALID, "null d_out/d_k");
    if (n_images <= 0) {   // pb:546-547
        if (!accumulate) DEISM_CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)K * 16, st));
        return DEISM_OK;
    }
    if (N < 0 || V < 0 || N > 20 || V > 20) return set_error(DEISM_ERR_INVALID, "orders must be 0..20");
    if (!d_C_nm_s || !d_C_vu_r || !d_A || !d_R_sI_r)
        return set_error(DEISM_ERR_INVALID, "null input pointer");
    if (algo < 0 || algo > 2) return set_error(DEISM_ERR_INVALID, "algo must be 0, 1 or 2");
    rc = validate_atten(atten);
    if (rc) return rc;

    const int Lmax = N + V + 1, nseg = Lmax * Lmax;
    const int NMs = (N + 1) * (2 * N + 1), NMr = (V + 1) * (2 * V + 1);
    PlanGuard guard;
    rc = guard.get(plan, N, V, h_W1, h_W2);
    if (rc) return rc;
    const OrgEntry *d_ent[2] = {guard.use->d_ent[0], guard.use->d_ent[1]};
    const int *d_seg[2] = {guard.use->d_seg[0], guard.use->d_seg[1]};

    int *flags = (int *)workspace(WS_ORG_FLAGS, 64);
    if (!flags) return DEISM_ERR_NOMEM;
    rc = launch_z_flat(atten, K, flags, st);
    if (rc) return rc;
    rc = launch_k_uniform(d_k, K, flags, st);
    if (rc) return rc;

    if (algo == 2 && coef_c128)
        return set_error(DEISM_ERR_INVALID, "the direct ORG sum (algo 2) takes complex64 coefficients");
    if (algo == 2) {
        // ---- direct quintuple sum per (image, frequency) ----
        OrgImage *rec = (OrgImage *)workspace(WS_ORG_IMG, (size_t)n_images * sizeof(OrgImage));
        cplx *Y = (cplx *)workspace(WS_ORG_Y, (size_t)nseg * n_images * sizeof(cplx));
        long nblk = (n_images + 127) / 128;
        cplx *partial = (cplx *)workspace(WS_ORG_PARTIAL, (size_t)nblk * K * sizeof(cplx));
        if (!rec || !Y || !partial) return DEISM_ERR_NOMEM;
        OrgPrepArgs pa;
        pa.n_images = n_images; pa.n_pad_cap = n_images; pa.Lmax = Lmax; pa.perm = nullptr;
        pa.meta = nullptr; pa.R_sI_r = d_R_sI_r; pa.component_major = 0; pa.mode = atten->mode;
        pa.angdep = atten->ang_dep_flag; pa.A = d_A; pa.Z = (const cplx *)atten->d_Z_S; pa.K = K;
        pa.Yp = nullptr; pa.R_tot = 0; pa.flags = nullptr;
        fill_room(pa.room, atten);
        org_prep_kernel<<<(unsigned)((n_images + 127) / 128), 128, 0, st>>>(pa, rec, Y);
        DEISM_LAUNCH_CHECK("org_prep_kernel");
        OrgDirectArgs da;
        da.n_sel = n_images; da.n_images = n_images; da.K = K; da.Lmax = Lmax; da.nseg = nseg;
        da.NMr = NMr;
        for (int i = 0; i < 2; ++i) { da.ent[i] = d_ent[i]; da.seg[i] = d_seg[i]; }
        da.Cs = (const float2 *)d_C_nm_s; da.sk = NMs; da.snm = 1; da.si = 0;
        da.Cr = (const float2 *)d_C_vu_r; da.Y = Y; da.rec = rec; da.sel = nullptr; da.k = d_k;
        da.mode = atten->mode; da.atten = atten->d_atten; da.Z = (const cplx *)atten->d_Z_S;
        da.shoebox = 1; da.partial = partial;
        dim3 grid((unsigned)nblk, (unsigned)(K < 65535 ? K : 65535));
        prof_begin("org_direct", st);
    org_direct_kernel<<<grid, 128, 0, st>>>(da);
    prof_end("org_direct", st);
        DEISM_LAUNCH_CHECK("org_direct_kernel");
        return launch_reduce_partials(partial, nblk, K, d_out, accumulate, st);
    }

    // ---- factorised path ----
    if (Lmax > ORG_LMAX) return set_error(DEISM_ERR_INVALID, "N+V+1 > %d", ORG_LMAX);
    // plane geometry: rows(l) = 2l+1 (even l) or 2l+2 (odd l); chunks hold whole degrees, <= rc rows each
    int rowoff[ORG_LMAX + 1];
    int4 chunks[ORG_MAXCHUNKS];
    int ncpt = 0, R_tot = 0, rc_rows = ORG_RC;
    for (int l = 0; l < Lmax; ++l) {
        rowoff[l] = R_tot;
        // 2l + 1 real rows; odd degrees carry one zero pad row (whole DMMA k-steps), even degrees
        // none (their a_0 row is a rank-one update in the consume kernel)
        const int rows = (l & 1) ? 2 * l + 2 : 2 * l + 1;
        if (rows > rc_rows) rc_rows = rows;
        R_tot += rows;
    }
    rowoff[Lmax] = R_tot;
    // chunks from the top degree downwards (the consume kernel's Clenshaw sum runs that way)
    for (int l1 = Lmax; l1 > 0;) {
        if (ncpt >= ORG_MAXCHUNKS) return set_error(DEISM_ERR_INVALID, "too many (l,mu) chunks");
        int l0 = l1 - 1;
        while (l0 > 0 && rowoff[l1] - rowoff[l0 - 1] <= rc_rows) --l0;
        chunks[ncpt++] = make_int4(rowoff[l0], rowoff[l1] - rowoff[l0], l0, l1);
        l1 = l0;
    }
    {   // the tables depend on Lmax only: upload (and synchronise on the stack copies) once per change
        static std::mutex geom_mu;
        static int uploaded_dev = -1, uploaded_lmax = -1;
        int dev = -1;
        DEISM_CUDA_TRY(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(geom_mu);
        if (dev != uploaded_dev || Lmax != uploaded_lmax) {
            uploaded_dev = -1;
            bump_generation();
            static double norm[ORG_LMAX * (ORG_LMAX + 1) / 2];      // guarded by geom_mu; consumed before the sync below returns
            for (int l = 0; l < Lmax; ++l)
                for (int am = 0; am <= l; ++am) {
                    double ratio = 1.0;
                    for (int i = l - am + 1; i <= l + am; ++i) ratio *= (double)i;
                    // signed with s_l = + + - - ...: the harmonic rows feed the Clenshaw recurrence
                    // of org_consume_kernel (org_small_kernel takes the sign back out)
                    norm[l * (l + 1) / 2 + am] = ((l & 2) ? -1.0 : 1.0) *
                                                 sqrt((double)(2 * l + 1) / (4.0 * 3.14159265358979323846) / ratio);
                }
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(g_org_norm, norm, (size_t)Lmax * (Lmax + 1) / 2 * sizeof(double), 0,
                                                   cudaMemcpyHostToDevice, st));
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_org_rowoff, rowoff, (Lmax + 1) * sizeof(int), 0,
                                                   cudaMemcpyHostToDevice, st));
            DEISM_CUDA_TRY(cudaMemcpyToSymbolAsync(c_org_chunks, chunks, ncpt * sizeof(int4), 0,
                                                   cudaMemcpyHostToDevice, st));
            DEISM_CUDA_TRY(cudaStreamSynchronize(st));   // stack tables consumed
            uploaded_dev = dev; uploaded_lmax = Lmax;
        }
    }

    const long n_ftiles = (K + ORG_TF - 1) / ORG_TF;
    const size_t b_rows = (size_t)R_tot * 2 * ORG_BS, y_rows = (size_t)R_tot * ORG_OS;
    double *Bp = (double *)workspace(WS_ORG_B, (size_t)8 * n_ftiles * b_rows * sizeof(double));
    if (!Bp) return DEISM_ERR_NOMEM;
    // B-table form: frequency tile width FT (16, 8 or 4: its coefficient rows must fit in shared
    // memory) and number of job shares per tile.  CTA time ~ FT / shares, CTAs = tiles x shares, one
    // CTA per SM: take the pair that needs the fewest CTA-times (few frequencies -> narrow tiles and
    // more shares, so that a frequency slab of a multi-GPU run still fills the SMs); a share keeps
    // >= 4 jobs per warp so that the dynamic job queue can balance them.
    int bt_ft = 0, bt_split = 1;
    size_t bt_smem = 0;
    {
        const long sms = sm_count();
        const int n_jobs_ = guard.use->n_jobs;
        const int max_split = n_jobs_ / (4 * (BT2_THREADS / 32)) > 1 ? n_jobs_ / (4 * (BT2_THREADS / 32)) : 1;
        double best = 1e300;
        for (int ft = 16; ft >= 4; ft >>= 1) {
            const size_t need = (size_t)(NMs + NMr) * (ft + 1) * sizeof(double2) + 16 + (size_t)BT2_THREADS * sizeof(int4);
            if ((long)need > (long)smem_optin_bytes()) continue;
            const long nx = (K + ft - 1) / ft;
            for (int sp = 1; sp <= max_split && sp <= 12; ++sp) {
                // a CTA's fixed time (launch, staging, the queue's tail) in units of "all jobs for one
                // frequency": ~4.5 us against ~3.3e-4 us per coupling entry
                const double fixed = 13600.0 / (double)(guard.use->nnz > 0 ? guard.use->nnz : 1);
                // narrow tiles run the same entry with fewer frequencies per lane group: measured
                // 1.8x the time per (entry, frequency) at FT = 4, ~1.3x at FT = 8
                const double slow = ft >= 16 ? 1.0 : ft == 8 ? 1.3 : 1.9;
                const double cost = ((double)ft * slow / sp + fixed) * (double)((nx * sp + sms - 1) / sms);
                if (cost < best * (1.0 - 1e-9)) { best = cost; bt_ft = ft; bt_split = sp; bt_smem = need; }
            }
        }
        if (const char *e = getenv("DEISM_BT_SPLIT")) { const int v = atoi(e); if (v > 0) bt_split = v; }
        if (getenv("DEISM_DEBUG"))
            fprintf(stderr, "[deism] btable: K=%ld nnz=%ld jobs=%d -> FT=%d shares=%d smem=%zu\n", (long)K,
                    guard.use->nnz, n_jobs_, bt_ft, bt_split, bt_smem);
    }
#ifdef ORG_BTABLE_V1
    bt_ft = 0;
#endif
    bt_elem *CsT = nullptr, *CrT = nullptr;
    if (!bt_ft) {
        if ((long)(NMs > NMr ? NMs : NMr) * K >= 2147483647L)
            return set_error(DEISM_ERR_INVALID, "too many frequencies for the B-table's 32-bit row offsets");
        CsT = (bt_elem *)workspace(WS_ORG_SMALL, (size_t)(NMs + NMr) * K * sizeof(bt_elem));
        if (!CsT) return DEISM_ERR_NOMEM;
        CrT = CsT + (size_t)NMs * K;
    }
    OrgSide *side = org_side();
    cudaStream_t bs = st;                 // stream of the image-independent half
    OrgJoin joiner;
    if (side) {
        DEISM_CUDA_TRY(cudaEventRecord(side->fork, st));
        DEISM_CUDA_TRY(cudaStreamWaitEvent(side->s, side->fork, 0));
        bs = side->s;
        joiner.side = side; joiner.st = st;
    }
    DEISM_CUDA_TRY(cudaMemsetAsync(Bp, 0, (size_t)8 * n_ftiles * b_rows * sizeof(double), bs));
    if (bt_ft) {
        Btable2Args b2;
        b2.K = K; b2.n_ftiles = n_ftiles; b2.R_tot = R_tot; b2.NMs = NMs; b2.NMr = NMr;
        b2.n_jobs = guard.use->n_jobs; b2.ent = guard.use->d_bt; b2.seg = guard.use->d_bt_seg;
        b2.jobs = guard.use->d_jobs; b2.Cs = d_C_nm_s; b2.Cr = d_C_vu_r; b2.k = d_k; b2.Bp = Bp;
        const long nx = (K + bt_ft - 1) / bt_ft;
        const int split = bt_split;
        b2.n_split = split;
        dim3 grid((unsigned)nx, (unsigned)split);
        prof_begin("org_btable", bs);
#define DEISM_BT2_LAUNCH(FT, IN)                                                                       \
        do {                                                                                           \
            DEISM_CUDA_TRY(cudaFuncSetAttribute(org_btable2_kernel<FT, IN>,                             \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_smem)); \
            org_btable2_kernel<FT, IN><<<grid, BT2_THREADS, bt_smem, bs>>>(b2);                         \
        } while (0)
        if (coef_c128) {
            if (bt_ft == 16) DEISM_BT2_LAUNCH(16, double2);
            else if (bt_ft == 8) DEISM_BT2_LAUNCH(8, double2);
            else DEISM_BT2_LAUNCH(4, double2);
        } else {
            if (bt_ft == 16) DEISM_BT2_LAUNCH(16, float2);
            else if (bt_ft == 8) DEISM_BT2_LAUNCH(8, float2);
            else DEISM_BT2_LAUNCH(4, float2);
        }
#undef DEISM_BT2_LAUNCH
        prof_end("org_btable", bs);
        DEISM_LAUNCH_CHECK("org_btable2_kernel");
    } else {
    if (coef_c128) {
        transpose_coef_kernel<double2><<<(unsigned)(((long)K * NMs + 255) / 256), 256, 0, bs>>>(
            (const double2 *)d_C_nm_s, K, NMs, CsT);
        DEISM_LAUNCH_CHECK("transpose_coef_kernel");
        transpose_coef_kernel<double2><<<(unsigned)(((long)K * NMr + 255) / 256), 256, 0, bs>>>(
            (const double2 *)d_C_vu_r, K, NMr, CrT);
        DEISM_LAUNCH_CHECK("transpose_coef_kernel");
    } else {
        transpose_coef_kernel<float2><<<(unsigned)(((long)K * NMs + 255) / 256), 256, 0, bs>>>(
            (const float2 *)d_C_nm_s, K, NMs, CsT);
        DEISM_LAUNCH_CHECK("transpose_coef_kernel");
        transpose_coef_kernel<float2><<<(unsigned)(((long)K * NMr + 255) / 256), 256, 0, bs>>>(
            (const float2 *)d_C_vu_r, K, NMr, CrT);
        DEISM_LAUNCH_CHECK("transpose_coef_kernel");
    }
    BtableArgs ba;
    ba.K = K; ba.n_ftiles = n_ftiles; ba.nseg = nseg; ba.R_tot = R_tot;
    for (int i = 0; i < 2; ++i) { ba.ent[i] = d_ent[i]; ba.seg[i] = d_seg[i]; }
    ba.CsT = CsT; ba.CrT = CrT; ba.k = d_k; ba.Bp = Bp;
    {
        dim3 grid((unsigned)((K + 127) / 128), (unsigned)(Lmax * (Lmax + 1) / 2), 2);
        prof_begin("org_btable", bs);
        org_btable_kernel<<<grid, 128, 0, bs>>>(ba);
        prof_end("org_btable", bs);
        DEISM_LAUNCH_CHECK("org_btable_kernel");
    }
    }
    joiner.record();

    const long n_pad_cap = (n_images + ORG_TI - 1) / ORG_TI * ORG_TI + 8 * ORG_TI;
    const long n_tiles_cap = n_pad_cap / ORG_TI;
    long *perm = (long *)workspace(WS_ARG_IDX, (size_t)(n_pad_cap + 32) * sizeof(long));
    OrgImage *rec = (OrgImage *)workspace(WS_ORG_IMG, (size_t)n_pad_cap * sizeof(OrgImage));
    double *Yp = (double *)workspace(WS_ORG_Y, (size_t)n_tiles_cap * y_rows * sizeof(double));
    if (!perm || !rec || !Yp) return DEISM_ERR_NOMEM;
    DEISM_CUDA_TRY(cudaMemsetAsync(Yp, 0, (size_t)n_tiles_cap * y_rows * sizeof(double), st));
    long *meta = perm + n_pad_cap;
    {
        const size_t par_bytes = (size_t)n_images <= 200000 ? ((size_t)n_images + 15) & ~(size_t)15 : 0;
        if (par_bytes > 48 * 1024)
            DEISM_CUDA_TRY(cudaFuncSetAttribute(org_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)par_bytes));
        org_sort_kernel<<<1, SORT_THREADS, par_bytes, st>>>(d_A, n_images, perm, meta, n_pad_cap, par_bytes ? 1 : 0);
    }
    DEISM_LAUNCH_CHECK("org_sort_kernel");
    OrgPrepArgs pa;
    pa.n_images = n_images; pa.n_pad_cap = n_pad_cap; pa.Lmax = Lmax; pa.perm = perm; pa.meta = meta;
    pa.R_sI_r = d_R_sI_r; pa.component_major = 0; pa.mode = atten->mode; pa.angdep = atten->ang_dep_flag;
    pa.A = d_A; pa.Z = (const cplx *)atten->d_Z_S; pa.K = K;
    pa.Yp = Yp; pa.R_tot = R_tot; pa.flags = flags;
    fill_room(pa.room, atten);
    org_prep_kernel<<<(unsigned)((n_pad_cap + 31) / 32), dim3(32, ORG_PREP_Y), 0, st>>>(pa, rec, nullptr);
    DEISM_LAUNCH_CHECK("org_prep_kernel");

    if (n_images <= ORG_SMALL_MAX_IMAGES) {
        // few images: the scalar kernel (no class-tile padding), one warp per image and frequency tile
        long n_chunks_s = (n_images + ORG_SMALL_LANES - 1) / ORG_SMALL_LANES;
        while (n_chunks_s > 1 && n_chunks_s * n_ftiles > 64L * sm_count()) n_chunks_s = (n_chunks_s + 1) / 2;
        cplx *partial_s = (cplx *)workspace(WS_ORG_PARTIAL, (size_t)n_chunks_s * K * sizeof(cplx));
        if (!partial_s) return DEISM_ERR_NOMEM;
        OrgSmallArgs sa;
        sa.K = K; sa.n_ftiles = n_ftiles; sa.n_images = n_images; sa.Lmax = Lmax; sa.R_tot = R_tot; sa.Bp = Bp;
        sa.Yp = Yp; sa.rec = rec; sa.meta = meta; sa.k = d_k; sa.mode = atten->mode; sa.atten = atten->d_atten;
        sa.Z = (const cplx *)atten->d_Z_S; sa.flags = flags; sa.partial = partial_s;
        joiner.join();
        dim3 grid_s((unsigned)n_ftiles, (unsigned)n_chunks_s);
        prof_begin("org_consume", st);
        org_small_kernel<<<grid_s, 32 * ORG_SMALL_LANES, 0, st>>>(sa);
        prof_end("org_consume", st);
        DEISM_LAUNCH_CHECK("org_small_kernel");
        return launch_reduce_partials(partial_s, n_chunks_s, K, d_out, accumulate, st);
    }
    // ring depth: as many stages as fit next to the fixed part (one CTA per SM)
    const size_t stage_bytes = (size_t)rc_rows * (2 * ORG_BS + ORG_OS) * 8;
    int nst = (int)(((size_t)smem_optin_bytes() / ORG_CTAS_PER_SM - (ORG_CTAS_PER_SM > 1 ? 1024 : 0) -
                     sizeof(OrgSmem)) / stage_bytes);
    if (nst > ORG_MAXST) nst = ORG_MAXST;
    if (nst < 2)
        return set_error(DEISM_ERR_INVALID, "ORG consume kernel: two stages of %zu B do not fit in shared memory",
                         stage_bytes);
    // Work split: CTA = (frequency tile, chunk of image tiles).  The tile count is known on the
    // device only (class padding); its upper bound is close.  Among chunk sizes that keep >= 6
    // tiles per CTA pick the one whose CTA count fills whole waves of one CTA per SM best.
    const long n_tiles_est = (n_images + ORG_TI - 1) / ORG_TI + 4;
    long tiles_per_chunk = 1;
    {
        const long sms = sm_count();
        double best = -1.0;
        for (long tpc = 6; tpc <= 24; ++tpc) {
            const long chunks = (n_tiles_est + tpc - 1) / tpc;
            const double waves = (double)(chunks * n_ftiles) / (double)sms;
            const double whole = (double)(long)(waves + 0.999999);
            // time ~ whole waves of tpc tiles; useful work ~ n_tiles_est * n_ftiles / sms tiles
            const double eff = ((double)n_tiles_est * n_ftiles / sms) / (whole * tpc);
            if (eff > best + 1e-9) { best = eff; tiles_per_chunk = tpc; }
        }
        if (n_tiles_est * n_ftiles <= 6 * sms) tiles_per_chunk = (n_tiles_est * n_ftiles + sms - 1) / sms;
        if (tiles_per_chunk < 1) tiles_per_chunk = 1;
    }
    long n_chunks = (n_tiles_cap + tiles_per_chunk - 1) / tiles_per_chunk;
    if (n_chunks > 65535) {
        n_chunks = 65535;
        tiles_per_chunk = (n_tiles_cap + n_chunks - 1) / n_chunks;
        n_chunks = (n_tiles_cap + tiles_per_chunk - 1) / tiles_per_chunk;
    }
    // chunks of decreasing size when the simulated hand-out of CTAs says they finish earlier (the
    // tile count is the host's estimate; the last chunk runs to the capacity bound and the kernel
    // clips it to the device-side count)
    // (the plan arrives with the second call of a shape; the partial sums are sized for it from the
    // first call on, so that call — possibly a graph capture — allocates nothing)
    GuidedPlan gplan;
    const long uni_chunks_est = (n_tiles_est + tiles_per_chunk - 1) / tiles_per_chunk;
    const bool may_guide = guided_chunks_possible(n_tiles_est, n_ftiles, uni_chunks_est);
    const bool guided = may_guide && guided_chunks(n_tiles_est, n_ftiles, sm_count(), uni_chunks_est,
                                                   tiles_per_chunk, (1L << 24), gplan);
    if (guided) {
        gplan.lo[gplan.n] = (int)(n_tiles_cap > gplan.lo[gplan.n] ? n_tiles_cap : gplan.lo[gplan.n]);
        n_chunks = gplan.n;
    }
    const long partial_rows = may_guide && n_chunks < DEISM_MAXCHUNKS ? DEISM_MAXCHUNKS : n_chunks;
    cplx *partial = (cplx *)workspace(WS_ORG_PARTIAL, (size_t)partial_rows * K * sizeof(cplx));
    if (!partial) return DEISM_ERR_NOMEM;

    OrgConsumeArgs ca;
    ca.K = K; ca.n_ftiles = n_ftiles; ca.Lmax = Lmax; ca.R_tot = R_tot; ca.ncpt = ncpt; ca.rc = rc_rows;
    ca.nst = nst;
    ca.Bp = Bp; ca.Yp = Yp; ca.rec = rec; ca.meta = meta; ca.k = d_k; ca.mode = atten->mode;
    ca.atten = atten->d_atten; ca.Z = (const cplx *)atten->d_Z_S; ca.flags = flags;
    ca.tiles_per_chunk = tiles_per_chunk; ca.partial = partial;
    ca.n_guided = 0;
    memset(ca.chunk_lo, 0, sizeof(ca.chunk_lo));
    if (guided) {
        ca.n_guided = gplan.n;
        memcpy(ca.chunk_lo, gplan.lo, sizeof(int) * (gplan.n + 1));
    }
    const size_t consume_smem = sizeof(OrgSmem) + (size_t)nst * stage_bytes;
    DEISM_CUDA_TRY(cudaFuncSetAttribute(org_consume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)consume_smem));
    joiner.join();
    dim3 grid((unsigned)n_ftiles, (unsigned)n_chunks);
    prof_begin("org_consume", st);
    org_consume_kernel<<<grid, ORG_THREADS, consume_smem, st>>>(ca);
    prof_end("org_consume", st);
    DEISM_LAUNCH_CHECK("org_consume_kernel");
    return launch_reduce_partials(partial, n_chunks, K, d_out, accumulate, st);
}

extern "C" int deism_org_shoebox(int64_t n_images, int64_t K, int N, int V, const float *d_C_nm_s,
                                 const float *d_C_vu_r, const float *h_W1, const float *h_W2,
                                 const int32_t *d_A, const float *d_R_sI_r,
                                 const deism_shoebox_atten *atten, const double *d_k, double *d_out,
                                 int accumulate, int algo, const deism_org_plan *plan, void *stream) {
    return org_shoebox_impl(n_images, K, N, V, d_C_nm_s, d_C_vu_r, 0, h_W1, h_W2, d_A, d_R_sI_r, atten, d_k,
                            d_out, accumulate, algo, plan, stream);
}

extern "C" int deism_org_shoebox_c128(int64_t n_images, int64_t K, int N, int V, const double *d_C_nm_s,
                                      const double *d_C_vu_r, const float *h_W1, const float *h_W2,
                                      const int32_t *d_A, const float *d_R_sI_r,
                                      const deism_shoebox_atten *atten, const double *d_k, double *d_out,
                                      int accumulate, const deism_org_plan *plan, void *stream) {
    return org_shoebox_impl(n_images, K, N, V, d_C_nm_s, d_C_vu_r, 1, h_W1, h_W2, d_A, d_R_sI_r, atten, d_k,
                            d_out, accumulate, 1, plan, stream);
}

extern "C" int deism_org_arg(int64_t n_images, int64_t K, int N, int V, const float *d_C_nm_s_ARG,
                             const float *d_C_vu_r, const float *h_W1, const float *h_W2,
                             const float *d_R_sI_r, const float *d_atten, const double *d_k,
                             const int64_t *h_image_index, int64_t n_sel, double *d_out, int accumulate,
                             const deism_org_plan *plan, void *stream) {
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
    if (N < 0 || V < 0 || N > 20 || V > 20) return set_error(DEISM_ERR_INVALID, "orders must be 0..20");
    if (!d_C_nm_s_ARG || !d_C_vu_r || !d_R_sI_r || !d_atten)
        return set_error(DEISM_ERR_INVALID, "null input pointer");
    const int Lmax = N + V + 1, nseg = Lmax * Lmax;
    const int NMs = (N + 1) * (2 * N + 1), NMr = (V + 1) * (2 * V + 1);
    PlanGuard guard;
    rc = guard.get(plan, N, V, h_W1, h_W2);
    if (rc) return rc;
    const OrgEntry *d_ent[2] = {guard.use->d_ent[0], guard.use->d_ent[1]};
    const int *d_seg[2] = {guard.use->d_seg[0], guard.use->d_seg[1]};

    OrgImage *rec = (OrgImage *)workspace(WS_ORG_IMG, (size_t)n_images * sizeof(OrgImage));
    cplx *Y = (cplx *)workspace(WS_ORG_Y, (size_t)nseg * n_images * sizeof(cplx));
    long nblk = (n_sel + 127) / 128;
    cplx *partial = (cplx *)workspace(WS_ORG_PARTIAL, (size_t)nblk * K * sizeof(cplx));
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
        DEISM_CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (!rec || !Y || !partial) return DEISM_ERR_NOMEM;
    OrgPrepArgs pa;
    pa.n_images = n_images; pa.n_pad_cap = n_images; pa.Lmax = Lmax; pa.perm = nullptr; pa.meta = nullptr;
    pa.R_sI_r = d_R_sI_r; pa.component_major = 1; pa.mode = DEISM_ATTEN_C64; pa.angdep = 0;
    pa.A = nullptr; pa.Z = nullptr; pa.K = K;
    pa.Yp = nullptr; pa.R_tot = 0; pa.flags = nullptr;
    for (int i = 0; i < 3; ++i) pa.room.LL[i] = pa.room.xs[i] = pa.room.xr[i] = 0.0;
    org_prep_kernel<<<(unsigned)((n_images + 127) / 128), 128, 0, st>>>(pa, rec, Y);
    DEISM_LAUNCH_CHECK("org_prep_kernel");
    OrgDirectArgs da;
    da.n_sel = n_sel; da.n_images = n_images; da.K = K; da.Lmax = Lmax; da.nseg = nseg; da.NMr = NMr;
    for (int i = 0; i < 2; ++i) { da.ent[i] = d_ent[i]; da.seg[i] = d_seg[i]; }
    da.Cs = (const float2 *)d_C_nm_s_ARG; da.sk = (long)NMs * n_images; da.snm = n_images; da.si = 1;
    da.Cr = (const float2 *)d_C_vu_r; da.Y = Y; da.rec = rec; da.sel = sel; da.k = d_k;
    da.mode = -1; da.atten = d_atten; da.Z = nullptr; da.shoebox = 0; da.partial = partial;
    dim3 grid((unsigned)nblk, (unsigned)(K < 65535 ? K : 65535));
    prof_begin("org_direct", st);
    org_direct_kernel<<<grid, 128, 0, st>>>(da);
    prof_end("org_direct", st);
    DEISM_LAUNCH_CHECK("org_direct_kernel");
    return launch_reduce_partials(partial, nblk, K, d_out, accumulate, st);
}
