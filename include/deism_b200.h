/*
 * deism_b200.h — C ABI of libdeism_b200.so: the B200-native (sm_100a) engine for the
 * image x frequency hot path behind deism.core_deism.DEISM.run_DEISM() of audiolabs/DEISM.
 *
 * Every entry point names the reference interface it replaces ("pb" =
 * deism/parallel_backends.py, "cd" = deism/core_deism.py, "room.cpp" etc. =
 * deism/libroom_src/*), so a maintainer can bind it with ctypes / cffi / pybind11 at that call
 * site (INTEGRATION.md shows the stubs).  Signatures use only plain pointers, sizes and PODs.
 *
 * Conventions
 *   - complex128 = interleaved (re, im) doubles; complex64 = interleaved floats (NumPy layout).
 *   - all arrays are C-contiguous with the reference's own shapes and dtypes, so the
 *     reference's arrays can be handed over without transposition;
 *   - "d_" pointers are DEVICE pointers on the current CUDA device, "h_" pointers are HOST
 *     pointers (small tables and scalars only).  Bulk host<->device staging is the caller's
 *     job; deism_b200/backend.py does it for NumPy callers (asynchronous copies from the
 *     caller's arrays as they are, pageable or pinned);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *     device-pointer entry points are asynchronous with respect to the host unless noted;
 *   - return value: DEISM_OK (0) or a DEISM_ERR_* code; deism_last_error() gives the message
 *     (thread-local).  There is no CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with DEISM_ERR_CUDA;
 *   - the caller owns every buffer it passes; the library keeps no pointer after a call
 *     returns (device scratch is library-owned, cached per device, freed by
 *     deism_release_workspace()).
 *
 * Threading (SURVEY 8b "Threading": the reference's kernels are driven by one Python thread)
 *   - every entry point may be called from any host thread, on any stream, for any device;
 *   - the library keeps per-device state that a call's kernels read while they run (cached
 *     scratch, the constant-memory tables of the (n,m) bases, row geometry and walls, one side
 *     stream), so it admits ONE call in flight per device and enforces that itself: a per-device
 *     mutex serialises the host side of the entry points, and when two consecutive calls on a
 *     device use DIFFERENT streams the second call's stream is made to wait (cudaStreamWaitEvent)
 *     for the first call's kernels.  Calls on one stream queue back to back without any wait;
 *     calls on different devices never wait for each other;
 *   - deism_convex_images_count + deism_convex_images_fill are one two-phase operation whose
 *     intermediate result lives in that per-device state: issue them from one thread with no
 *     other convex call on the device in between;
 *   - during CUDA-graph stream capture the cross-stream ordering is the capturer's job (no
 *     event is recorded or waited on while `stream` is capturing);
 *   - deism_last_error() is thread-local; deism_profile_* and deism_release_workspace() are
 *     control calls for a quiescent library (no compute call in flight).
 */
#ifndef DEISM_B200_H
#define DEISM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DEISM_API __attribute__((visibility("default")))
#else
#define DEISM_API
#endif

#define DEISM_OK 0
#define DEISM_ERR_INVALID 1 /* bad argument (the ValueError cases of pb:922,935) */
#define DEISM_ERR_CUDA 2    /* CUDA runtime / launch failure, or no sm_100 device */
#define DEISM_ERR_NOMEM 3   /* device or host allocation failed */

/* ------------------------------------------------------------------------------------ */
/* library / device                                                                      */
/* ------------------------------------------------------------------------------------ */
DEISM_API int deism_b200_version(void);          /* 10000*major + 100*minor + patch */
DEISM_API const char *deism_last_error(void);    /* message of the last failing call */
/* First 16 hex digits of the SHA-256 of the sources (csrc/*.cu, common.cuh, this header) the
 * binary was built from; deism_b200/_lib.py refuses (rebuilds) a library whose fingerprint does
 * not match the sources next to it. */
DEISM_API const char *deism_b200_source_hash(void);
/* Fills SM count, HBM bytes and compute capability of `device` (or the current one if <0). */
DEISM_API int deism_device_info(int device, int *sm_count, size_t *hbm_bytes, int *cc_major,
                                int *cc_minor);
DEISM_API int deism_release_workspace(void);     /* frees cached device scratch */
/* CUDA-graph support.  After a first (warm-up) call with given shapes the shoebox entry points
 * neither allocate nor synchronise, so a caller may capture them into a CUDA graph and replay it
 * (deism_b200/backend.py does).  A captured graph bakes in the library's scratch pointers and the
 * contents of its constant-memory tables; this counter changes whenever either changes: replay a
 * graph only while it still has the value it had at capture. */
DEISM_API uint64_t deism_state_generation(void);
/* To be called right after launching a captured graph of entry-point calls on `stream`: adds the
 * graph's kernel count to deism_kernel_launch_count() and lets the library order its next call
 * behind the replayed work. */
DEISM_API int deism_graph_replayed(int n_kernels, void *stream);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
DEISM_API uint64_t deism_kernel_launch_count(void);
/* Host-only view of the image-chunk schedule the two tensor-core kernels use (no GPU needed; for
 * tests and tuning): n_tiles image tiles per frequency tile, n_ftiles frequency tiles, `slots` CTAs
 * resident at a time, against `uni_chunks` equal chunks of `uni_tpc` tiles.  Returns 1 and fills
 * lo[0..*n] (chunk c = tiles [lo[c], lo[c+1]), at most 96 chunks, sizes non-increasing) when a
 * schedule of shrinking chunks finishes earlier under the block scheduler's in-order hand-out,
 * else 0 (the uniform cut stays).  Like the kernels' own calls it only records a shape the first time
 * it sees it (returns 0) and plans it when the shape comes again: the simulations cost 0.3-0.7 ms of
 * host time, more than the schedule saves on a single call. */
DEISM_API int deism_chunk_plan(int64_t n_tiles, int64_t n_ftiles, int64_t slots, int64_t uni_chunks,
                               int64_t uni_tpc, int64_t max_tiles_per_chunk, int32_t *lo, int32_t *n);
/* Per-kernel device timing for the roofline report: when enabled, the accumulation entry points
 * bracket their dominant kernels with CUDA events on the launching stream.
 * deism_profile_get("lc_main" | "org_consume" | "org_btable" | "org_direct" | "arg_lc" |
 * "shoebox_fill" | "convex_expand", ...) synchronises those events and returns the summed
 * milliseconds and the number of launches since the last deism_profile_reset(). */
DEISM_API int deism_profile_enable(int on);
DEISM_API int deism_profile_reset(void);
DEISM_API int deism_profile_get(const char *kernel, double *ms_total, int *launches);

/* ------------------------------------------------------------------------------------ */
/* Shoebox wall attenuation source (what the reference calls images["atten_all"])        */
/* ------------------------------------------------------------------------------------ */
/*
 * The reference materialises atten[I,K] as complex64 while it enumerates images
 * (cd:2900-2928; 8 bytes per term of HBM/host traffic) or, in "compact" storage, rebuilds it
 * per batch from the f32-rounded angles (pb:140-186).  This engine can also FUSE the factor
 * into the accumulation kernels, reproducing either numerics:
 */
#define DEISM_ATTEN_C64 0        /* d_atten: complex64[I,K]  (reference default storage)      */
#define DEISM_ATTEN_C128 1       /* d_atten: complex128[I,K]                                   */
#define DEISM_ATTEN_FUSED_ROOM 2 /* in-kernel from A + room geometry in FP64, rounded to c64:
                                    bit-for-bit the values of cd:2847-2928 (default storage)   */
#define DEISM_ATTEN_FUSED_ANGLES 3 /* in-kernel from A + f32 angles R_sI_r, NOT rounded:
                                    the values of pb:140-186 (compact storage)                 */

typedef struct deism_shoebox_atten {
    int32_t mode;            /* DEISM_ATTEN_*                                                  */
    int32_t ang_dep_flag;    /* params["angDepFlag"] (1 = angle dependent)                     */
    const void *d_atten;     /* modes 0/1: [I,K]                                               */
    const int32_t *d_A;      /* modes 2/3: int32[I,6] = (q_x,q_y,q_z,p_x,p_y,p_z)              */
    const float *d_R_sI_r;   /* mode 3: float32[I,3] = (phi, theta, r)                         */
    const double *d_Z_S;     /* modes 2/3: complex128[6,K], wall rows x1,x2,y1,y2,z1,z2        */
    double room_size[3];     /* mode 2: params["roomSize"]                                     */
    double pos_source[3];    /* mode 2: params["posSource"]                                    */
    double pos_receiver[3];  /* mode 2: params["posReceiver"]                                  */
} deism_shoebox_atten;

/* ------------------------------------------------------------------------------------ */
/* (1) Shoebox image lattice   — replaces cd:2649-2728 (_count_shoebox_images_nofs_v2_numba),
 *     cd:2731-2928 (_fill_...), the prefix sum of cd:3390-3394, the direct-path removal of
 *     cd:3442-3451, and deism/count_reflections.cpp                                         */
/* ------------------------------------------------------------------------------------ */
/* Counts images per task t = (4p_x+2p_y+p_z)*(N_o+1) + order.  d_early/d_late: int64[8*(N_o+1)].
 * remove_direct != 0 drops the (q=0,p=0) image (params["ifRemoveDirectPath"]). */
DEISM_API int deism_shoebox_count_images(const double h_room_size[3], const double h_pos_source[3],
                                         const double h_pos_receiver[3], double max_dist_sq,
                                         int N_o, int N_o_ORG, int remove_direct,
                                         int64_t *d_early_counts, int64_t *d_late_counts,
                                         void *stream);
/* Fills the compacted lists in the reference order.  Offsets are the exclusive
 * prefix sums of the counts (computed on the device).  Any output pointer may be NULL. */
DEISM_API int deism_shoebox_fill_images(const double h_room_size[3], const double h_pos_source[3],
                                        const double h_pos_receiver[3], double max_dist_sq,
                                        int N_o, int N_o_ORG, int remove_direct,
                                        const int64_t *d_early_counts, const int64_t *d_late_counts,
                                        int32_t *d_A_early, float *d_R_sI_r_early,
                                        float *d_R_s_rI_early, float *d_R_r_sI_early,
                                        int32_t *d_A_late, float *d_R_sI_r_late,
                                        float *d_R_s_rI_late, float *d_R_r_sI_late, void *stream);
/* Materialises atten[I,K] (complex64 if out_c128 == 0, else complex128) from a FUSED_* spec —
 * the array the reference stores as images["atten_all_*"] (cd:2900-2928) or rebuilds per
 * batch (pb:140-186). */
DEISM_API int deism_shoebox_attenuation(int64_t n_images, int64_t K, const deism_shoebox_atten *spec,
                                        void *d_out, int out_c128, void *stream);

/* ------------------------------------------------------------------------------------ */
/* (2) Accumulation kernels — replace the Numba kernels of pb:189-460                     */
/* ------------------------------------------------------------------------------------ */
/* Shoebox DEISM-LC  (pb:264-325 _numba_LC_matrix_batch + batching of pb:594-654).
 * h_n_all..h_u_all: HOST int64[Js]/[Jr] (params["n_all"] ...); d_C_s/d_C_r: complex64[K,Js],
 * [K,Jr] (kept complex64 as at pb:619-620); d_R_s_rI/d_R_r_sI: float32[I,3]; d_k: float64[K];
 * d_out: complex128[K]; accumulate != 0 adds to d_out instead of overwriting. */
DEISM_API int deism_lc_shoebox(int64_t n_images, int64_t K, int Js, int Jr,
                               const int64_t *h_n_all, const int64_t *h_m_all,
                               const int64_t *h_v_all, const int64_t *h_u_all,
                               const float *d_C_s_vec, const float *d_C_r_vec,
                               const float *d_R_s_rI, const float *d_R_r_sI,
                               const deism_shoebox_atten *atten, const double *d_k,
                               double *d_out, int accumulate, void *stream);

/* Coupling plan: the non-zero quintuples (n,m,v,u,l) of the Wigner tables with their real
 * coefficients, sorted by (l, mu) and resident on the device.  Building it costs a host pass over
 * W_2 (1.1 M entries at N=V=10); callers that run the same directivity orders repeatedly create it
 * once (like an FFT plan) and pass it to deism_org_shoebox / deism_org_arg instead of the tables. */
typedef struct deism_org_plan deism_org_plan;
DEISM_API int deism_org_plan_create(int N, int V, const float *h_W1, const float *h_W2,
                                    deism_org_plan **plan);
DEISM_API int deism_org_plan_destroy(deism_org_plan *plan);

/* Shoebox DEISM-ORG (pb:189-261 _numba_ORG_batch + batching of pb:543-591).
 * d_C_nm_s: complex64[K,N+1,2N+1], d_C_vu_r: complex64[K,V+1,2V+1] (signed m/u wrap as in
 * NumPy); h_W1: HOST complex64[N+1,V+1,N+V+1]; h_W2: HOST complex64[N+1,V+1,N+V+1,2N+1,2V+1]
 * (params["Wigner"]); d_A: int32[I,6]; d_R_sI_r: float32[I,3].  plan: optional (NULL = build the
 * coupling lists from h_W1/h_W2 inside the call; non-NULL = h_W1/h_W2 are ignored).
 * algo: 0 = auto, 1 = factorised B[parity,k,l,mu] table, 2 = direct quintuple sum. */
DEISM_API int deism_org_shoebox(int64_t n_images, int64_t K, int N, int V,
                                const float *d_C_nm_s, const float *d_C_vu_r,
                                const float *h_W1, const float *h_W2, const int32_t *d_A,
                                const float *d_R_sI_r, const deism_shoebox_atten *atten,
                                const double *d_k, double *d_out, int accumulate, int algo,
                                const deism_org_plan *plan, void *stream);

/* The same with complex128 coefficient tables (d_C_nm_s: complex128[K,N+1,2N+1], d_C_vu_r:
 * complex128[K,V+1,2V+1]).  The reference's own pipeline stores complex64 (cd:1184,1236) and its
 * ORG path up-casts whatever it is handed (pb:556-557), so a caller's own complex128 tables are
 * used at full precision there; this entry point does the same (factorised algorithm). */
DEISM_API int deism_org_shoebox_c128(int64_t n_images, int64_t K, int N, int V,
                                     const double *d_C_nm_s, const double *d_C_vu_r,
                                     const float *h_W1, const float *h_W2, const int32_t *d_A,
                                     const float *d_R_sI_r, const deism_shoebox_atten *atten,
                                     const double *d_k, double *d_out, int accumulate,
                                     const deism_org_plan *plan, void *stream);

/* Convex DEISM-ARG LC (pb:328-385).  d_C_s_ARG_vec: complex64[K,Js,I] (image index innermost,
 * cd:1193-1217), d_C_r_vec: complex64[K,Jr], d_R_sI_r: float32[3,I] (component-major),
 * d_atten: complex64[K,I].  h_image_index (HOST int64[n_sel] or NULL = all) selects the columns
 * processed (the late_indices of pb:886-899). */
DEISM_API int deism_lc_arg(int64_t n_images, int64_t K, int Js, int Jr, const int64_t *h_n_all,
                           const int64_t *h_m_all, const int64_t *h_v_all, const int64_t *h_u_all,
                           const float *d_C_s_ARG_vec, const float *d_C_r_vec,
                           const float *d_R_sI_r, const float *d_atten, const double *d_k,
                           const int64_t *h_image_index, int64_t n_sel, double *d_out,
                           int accumulate, void *stream);

/* Convex DEISM-ARG ORG (pb:388-460).  d_C_nm_s_ARG: complex64[K,N+1,2N+1,I]. */
DEISM_API int deism_org_arg(int64_t n_images, int64_t K, int N, int V, const float *d_C_nm_s_ARG,
                            const float *d_C_vu_r, const float *h_W1, const float *h_W2,
                            const float *d_R_sI_r, const float *d_atten, const double *d_k,
                            const int64_t *h_image_index, int64_t n_sel, double *d_out,
                            int accumulate, const deism_org_plan *plan, void *stream);

/* ------------------------------------------------------------------------------------ */
/* (3) Precompute — replaces cd:1842-1913 pre_calc_Wigner (sympy, 169 s at N=V=10)        */
/* ------------------------------------------------------------------------------------ */
/* Exact (big-rational) Wigner 3j tables rounded to float32 exactly like the reference
 * (float64 -> complex64).  Host buffers: h_W1 complex64[N+1,V+1,N+V+1],
 * h_W2 complex64[N+1,V+1,N+V+1,2N+1,2V+1], W2[n,v,l,m,u] = (n v l; -m u m-u), negative m/u
 * stored with NumPy wrap-around.  CPU code (setup, not the hot path). */
DEISM_API int deism_wigner_tables(int N, int V, float *h_W1, float *h_W2);

/* ------------------------------------------------------------------------------------ */
/* (4) Convex-room reflection tree — replaces Room_deism<3>::image_source_model            */
/*     (room.cpp:271-511) behind libroom_deism (libroom.cpp:58-127)                        */
/* ------------------------------------------------------------------------------------ */
typedef struct deism_convex_walls {
    int32_t n_walls;
    int32_t max_corners;        /* row stride of h_flat_corners                             */
    const float *h_normal;      /* float32[n_walls,3]   Wall_deism.normal                   */
    const float *h_origin;      /* float32[n_walls,3]   Wall_deism.origin                   */
    const float *h_basis;       /* float32[n_walls,3,2] Wall_deism.basis (row-major 3x2)    */
    const float *h_flat_corners;/* float32[n_walls,2,max_corners] Wall_deism.flat_corners   */
    const int32_t *h_n_corners; /* int32[n_walls]                                           */
    const float *h_reflection;  /* float32[n_walls,3,3] Wall_deism.reflection_matrix        */
    const float *h_impedance;   /* float32[n_walls,K]  real impedance per band (wall.cpp:428) */
} deism_convex_walls;

/* Phase 1: enumerate the visible images (pre-order DFS order of room.cpp:351-420).
 * Returns the count in *h_n_images; results stay in library scratch until phase 2. */
DEISM_API int deism_convex_images_count(const deism_convex_walls *walls, const float h_source[3],
                                        const float h_mic[3], int ism_order, int64_t *h_n_images,
                                        void *stream);
/* Phase 2: export.  d_sources float32[3,I], d_orders int32[I], d_gen_walls int32[I],
 * d_attenuations float32[K,I], d_reflection float32[I,3,3].  Any pointer may be NULL. */
DEISM_API int deism_convex_images_fill(const deism_convex_walls *walls, int64_t K, float *d_sources,
                                       int32_t *d_orders, int32_t *d_gen_walls,
                                       float *d_attenuations, float *d_reflection, void *stream);

/* ------------------------------------------------------------------------------------ */
/* (4b) DEISM-ARG source directivity of every image source                    */
/* ------------------------------------------------------------------------------------ */
/* Replaces cal_C_nm_s_new (deism/core_deism.py:1545-1605), its per-image least-squares fit
 * SHCs_from_pressure_LS (core_deism.py:1481-1509) and vectorize_C_nm_s_ARG
 * (core_deism.py:1193-1217).  For every image the sampling directions are mirrored by the
 * image's reflection matrix, order-N spherical-harmonic coefficients are fitted to the sampled
 * pressures and divided by h_n^(2)(k r0).
 *   d_reflection : float32[I,3,3]  (deism_convex_images_fill's d_reflection)
 *   d_coords     : float64[3,n_dir] Cartesian sampling directions (already rotated, cd:1676-1699)
 *   d_Psh        : complex128[K,n_dir] sampled pressures
 *   d_C_vec      : out, complex64[K,(N+1)^2,I], image index innermost (= C_nm_s_ARG_vec)
 * N <= 11 (the normal matrix of the fit lives in shared memory).  DEISM_ERR_INVALID if the
 * directions do not determine the fit. */
DEISM_API int deism_arg_refit(int64_t n_images, int64_t K, int N, int64_t n_dir,
                              const float *d_reflection, const double *d_coords,
                              const double *d_Psh, const double *d_k, double r0, float *d_C_vec,
                              void *stream);

/* ------------------------------------------------------------------------------------ */
/* (4c) Wall material data on the dense frequency grid                        */
/* ------------------------------------------------------------------------------------ */
/* Replaces interpolate_functions (deism/core_deism.py:1126-1163): PCHIP interpolation of the real
 * and imaginary parts of d_data complex128[rows, n_bands] (sampled at d_sparse_freqs, strictly
 * increasing) at d_dense_freqs[K], end values held outside the band, a single band broadcast.
 * d_out: complex128[rows, K] — e.g. the Z_S[6, K] of deism_shoebox_atten.  n_bands <= 64. */
DEISM_API int deism_pchip_interp(int rows, int n_bands, int64_t K, const double *d_sparse_freqs,
                                 const double *d_data, const double *d_dense_freqs, double *d_out,
                                 void *stream);

/* ------------------------------------------------------------------------------------ */
/* (5) Micro-benchmarks used by bench.py to measure the FP64 roofline denominators         */
/* ------------------------------------------------------------------------------------ */
/* Runs a register-resident DFMA (kind 0) or DMMA m8n8k4 (kind 1) loop on every SM and returns
 * the achieved TFLOP/s (2 flops per FMA) in *tflops, timed with CUDA events.  kind 2: half the
 * warps DMMA, half DFMA (the two share one datapath); kinds 3/4/5: DMMA fed by shared-memory
 * fragment loads at 16 / 4 / 8 warps per SM. */
DEISM_API int deism_fp64_peak(int kind, int iters, double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* DEISM_B200_H */
