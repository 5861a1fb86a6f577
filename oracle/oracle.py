"""oracle/oracle.py — TEST INFRASTRUCTURE ONLY (the checker, never the product).

ctypes front-end of oracle/liboracle.so (deism_oracle.c) plus a restatement of the Python
dispatch layer of the reference's Numba backend:

  run_DEISM(params)       ~ deism/parallel_backends.py:912-922  (run_DEISM_numba)
  run_DEISM_ARG(params)   ~ deism/parallel_backends.py:925-935  (run_DEISM_ARG_numba)
  shoebox_images(params)  ~ deism/core_deism.py:3326-3473       (v2-numba image generator)
  merge_images(images)    ~ deism/core_deism.py:4266-4295
  wigner_tables(N, V)     ~ deism/core_deism.py:1842-1913       (sympy, exact)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Pinned against reference outputs by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc + OpenMP) if missing."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_ORG_batch.restype = C.c_int
        _lib.oracle_ARG_ORG_batch.restype = C.c_int
        _lib.oracle_LC_matrix_batch.restype = C.c_int
        _lib.oracle_ARG_LC_batch.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(C.c_int(int(n)))


def _p(arr, typ):
    return arr.ctypes.data_as(typ)


def _c(arr, dtype):
    return np.ascontiguousarray(np.asarray(arr, dtype=dtype))


# --------------------------------------------------------------------------------------
# scalar helpers (pb:40-118)
# --------------------------------------------------------------------------------------
def sph_harm(m, n, phi, theta) -> complex:
    out = np.zeros(2)
    lib().oracle_sph_harm(C.c_long(m), C.c_long(n), C.c_double(phi), C.c_double(theta),
                          _p(out, _f64p))
    return complex(out[0], out[1])


def sphankel2(n, kr) -> complex:
    out = np.zeros(2)
    lib().oracle_sphankel2(C.c_long(n), C.c_double(kr), _p(out, _f64p))
    return complex(out[0], out[1])


# --------------------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------------------
def build_shoebox_attenuation_batch(A, R_sI_r, Z_S, ang_dep_flag):
    """pb:140-186 (compact storage: cosines from the f32-rounded angles)."""
    A = _c(A, np.int64)
    R = _c(R_sI_r, np.float64)
    Z = _c(Z_S, np.complex128)
    n, K = A.shape[0], Z.shape[1]
    out = np.empty((n, K), dtype=np.complex128)
    lib().oracle_build_shoebox_attenuation_batch(
        _p(A, _i64p), _p(R, _f64p), _p(Z, _f64p), C.c_int(int(ang_dep_flag)),
        C.c_long(n), C.c_long(K), _p(out, _f64p))
    return out


def ORG_batch(N, V, C_nm_s, C_vu_r, A, atten, x0, W_1, W_2, k):
    """pb:189-261."""
    C_nm_s = _c(C_nm_s, np.complex128)
    C_vu_r = _c(C_vu_r, np.complex128)
    A = _c(A, np.int64)
    atten = _c(atten, np.complex128)
    x0 = _c(x0, np.float64)
    W_1 = _c(W_1, np.complex128)
    W_2 = _c(W_2, np.complex128)
    k = _c(k, np.float64)
    K, n = k.shape[0], A.shape[0]
    assert C_nm_s.shape == (K, N + 1, 2 * N + 1) and C_vu_r.shape == (K, V + 1, 2 * V + 1)
    assert atten.shape == (n, K) and x0.shape == (n, 3)
    assert W_2.shape == (N + 1, V + 1, N + V + 1, 2 * N + 1, 2 * V + 1)
    out = np.zeros(K, dtype=np.complex128)
    rc = lib().oracle_ORG_batch(
        C.c_long(N), C.c_long(V), _p(C_nm_s, _f64p), _p(C_vu_r, _f64p), _p(A, _i64p),
        _p(atten, _f64p), _p(x0, _f64p), _p(W_1, _f64p), _p(W_2, _f64p), _p(k, _f64p),
        C.c_long(n), C.c_long(K), _p(out, _f64p))
    if rc != 0:
        raise MemoryError("oracle_ORG_batch")
    return out


def ARG_ORG_batch(N, V, C_nm_s_ARG, C_vu_r, atten, R_sI_r, W_1, W_2, k):
    """pb:388-460."""
    C_nm_s_ARG = _c(C_nm_s_ARG, np.complex128)
    C_vu_r = _c(C_vu_r, np.complex128)
    atten = _c(atten, np.complex128)
    R = _c(R_sI_r, np.float64)
    W_1 = _c(W_1, np.complex128)
    W_2 = _c(W_2, np.complex128)
    k = _c(k, np.float64)
    K, n = k.shape[0], R.shape[1]
    assert C_nm_s_ARG.shape == (K, N + 1, 2 * N + 1, n) and atten.shape == (K, n)
    out = np.zeros(K, dtype=np.complex128)
    rc = lib().oracle_ARG_ORG_batch(
        C.c_long(N), C.c_long(V), _p(C_nm_s_ARG, _f64p), _p(C_vu_r, _f64p),
        _p(atten, _f64p), _p(R, _f64p), _p(W_1, _f64p), _p(W_2, _f64p), _p(k, _f64p),
        C.c_long(n), C.c_long(K), _p(out, _f64p))
    if rc != 0:
        raise MemoryError("oracle_ARG_ORG_batch")
    return out


def LC_matrix_batch(n_all, m_all, v_all, u_all, C_nm_s_vec, C_vu_r_vec, R_s_rI, R_r_sI,
                    atten, k):
    """pb:264-325.  C_*_vec stay complex64 (pb:619-620)."""
    n_all, m_all = _c(n_all, np.int64), _c(m_all, np.int64)
    v_all, u_all = _c(v_all, np.int64), _c(u_all, np.int64)
    Cs = _c(C_nm_s_vec, np.complex64)
    Cr = _c(C_vu_r_vec, np.complex64)
    Rs, Rr = _c(R_s_rI, np.float64), _c(R_r_sI, np.float64)
    atten = _c(atten, np.complex128)
    k = _c(k, np.float64)
    K, n = k.shape[0], Rs.shape[0]
    assert Cs.shape == (K, n_all.shape[0]) and Cr.shape == (K, v_all.shape[0])
    assert atten.shape == (n, K)
    out = np.zeros(K, dtype=np.complex128)
    rc = lib().oracle_LC_matrix_batch(
        _p(n_all, _i64p), _p(m_all, _i64p), _p(v_all, _i64p), _p(u_all, _i64p),
        C.c_long(n_all.shape[0]), C.c_long(v_all.shape[0]), _p(Cs, _f32p), _p(Cr, _f32p),
        _p(Rs, _f64p), _p(Rr, _f64p), _p(atten, _f64p), _p(k, _f64p),
        C.c_long(n), C.c_long(K), _p(out, _f64p))
    if rc != 0:
        raise MemoryError("oracle_LC_matrix_batch")
    return out


def ARG_LC_batch(n_all, m_all, v_all, u_all, C_nm_s_ARG_vec, C_vu_r_vec, R_sI_r, atten, k):
    """pb:328-385."""
    n_all, m_all = _c(n_all, np.int64), _c(m_all, np.int64)
    v_all, u_all = _c(v_all, np.int64), _c(u_all, np.int64)
    Cs = _c(C_nm_s_ARG_vec, np.complex128)
    Cr = _c(C_vu_r_vec, np.complex128)
    R = _c(R_sI_r, np.float64)
    atten = _c(atten, np.complex128)
    k = _c(k, np.float64)
    K, n = k.shape[0], R.shape[1]
    assert Cs.shape == (K, n_all.shape[0], n) and atten.shape == (K, n)
    out = np.zeros(K, dtype=np.complex128)
    rc = lib().oracle_ARG_LC_batch(
        _p(n_all, _i64p), _p(m_all, _i64p), _p(v_all, _i64p), _p(u_all, _i64p),
        C.c_long(n_all.shape[0]), C.c_long(v_all.shape[0]), _p(Cs, _f64p), _p(Cr, _f64p),
        _p(R, _f64p), _p(atten, _f64p), _p(k, _f64p), C.c_long(n), C.c_long(K),
        _p(out, _f64p))
    if rc != 0:
        raise MemoryError("oracle_ARG_LC_batch")
    return out


# --------------------------------------------------------------------------------------
# dispatch layer (pb:471-935)
# --------------------------------------------------------------------------------------
_TEMP_TARGET_MB = 256  # pb:34-35


def _batch_size(params, n_images, n_freqs, key):
    """pb:471-522."""
    if n_images <= 0:
        return 1
    requested = params.get("numParaImages")
    if requested is None:
        requested = n_images
    requested = max(1, min(int(requested), int(n_images)))
    target = max(1, int(params.get(key, _TEMP_TARGET_MB))) * 1024 * 1024
    auto = max(1, target // max(1, int(n_freqs) * 32))
    return max(1, min(requested, auto))


def _image_freq_layout(arr, n_images):
    arr2 = np.atleast_2d(np.asarray(arr))  # pb:525-529
    if arr2.shape[0] != n_images and arr2.shape[1] == n_images:
        arr2 = arr2.T
    return arr2


def _org_in_batches(params, A, R_sI_r, atten, Wigner):
    """pb:543-591."""
    n = len(A)
    k = _c(params["waveNumbers"], np.float64)
    if n == 0:
        return np.zeros(k.shape[0], dtype=np.complex128)
    A = _c(A, np.int64)
    x0 = _c(R_sI_r, np.float64)
    att = None if atten is None else _image_freq_layout(atten, n)
    Cs = np.ascontiguousarray(params["C_nm_s"].astype(np.complex128))
    Cr = np.ascontiguousarray(params["C_vu_r"].astype(np.complex128))
    W1 = np.ascontiguousarray(Wigner["W_1_all"].astype(np.complex128))
    W2 = np.ascontiguousarray(Wigner["W_2_all"].astype(np.complex128))
    bs = _batch_size(params, n, k.shape[0], "numbaORGTempTargetMB")
    P = np.zeros(k.shape[0], dtype=np.complex128)
    for s in range(0, n, bs):
        e = min(s + bs, n)
        if att is None:
            ab = build_shoebox_attenuation_batch(A[s:e], x0[s:e], params["impedance"],
                                                 int(params["angDepFlag"]))
        else:
            ab = np.ascontiguousarray(np.asarray(att[s:e], dtype=np.complex128))
        P += ORG_batch(params["sourceOrder"], params["receiverOrder"], Cs, Cr, A[s:e], ab,
                       x0[s:e], W1, W2, k)
    return P


def _lc_in_batches(params, A, R_sI_r, R_s_rI, R_r_sI, atten):
    """pb:594-654."""
    n = np.asarray(R_s_rI).shape[0]
    k = _c(params["waveNumbers"], np.float64)
    if n == 0:
        return np.zeros(k.shape[0], dtype=np.complex128)
    A = _c(A, np.int64)
    x0 = _c(R_sI_r, np.float64)
    Rs, Rr = _c(R_s_rI, np.float64), _c(R_r_sI, np.float64)
    att = None if atten is None else _image_freq_layout(atten, n)
    bs = _batch_size(params, n, k.shape[0], "numbaLCTempTargetMB")
    P = np.zeros(k.shape[0], dtype=np.complex128)
    for s in range(0, n, bs):
        e = min(s + bs, n)
        if att is None:
            ab = build_shoebox_attenuation_batch(A[s:e], x0[s:e], params["impedance"],
                                                 int(params["angDepFlag"]))
        else:
            ab = np.ascontiguousarray(np.asarray(att[s:e], dtype=np.complex128))
        P += LC_matrix_batch(params["n_all"], params["m_all"], params["v_all"],
                             params["u_all"], params["C_nm_s_vec"], params["C_vu_r_vec"],
                             Rs[s:e], Rr[s:e], ab, k)
    return P


def run_DEISM(params):
    """pb:912-922 (shoebox)."""
    method = params["DEISM_method"]
    im = params["images"]
    if method == "ORG":
        return _org_in_batches(params, im["A"], im["R_sI_r_all"], im.get("atten_all"),
                               params["Wigner"])
    if method == "LC":
        return _lc_in_batches(params, im["A"], im["R_sI_r_all"], im["R_s_rI_all"],
                              im["R_r_sI_all"], im.get("atten_all"))
    if method == "MIX":
        P = np.zeros(np.asarray(params["waveNumbers"]).size, dtype=np.complex128)
        if len(im["A_early"]) > 0:
            P += _org_in_batches(params, im["A_early"], im["R_sI_r_all_early"],
                                 im.get("atten_all_early"), params["Wigner"])
        if len(im["A_late"]) > 0:
            P += _lc_in_batches(params, im["A_late"], im["R_sI_r_all_late"],
                                im["R_s_rI_all_late"], im["R_r_sI_all_late"],
                                im.get("atten_all_late"))
        return P
    raise ValueError(f"Unknown DEISM method: {method}")


def run_DEISM_ARG(params):
    """pb:925-935 (convex)."""
    method = params["DEISM_method"]
    im = params["images"]
    k = _c(params["waveNumbers"], np.float64)
    N, V = params["sourceOrder"], params["receiverOrder"]
    if method == "LC":
        return ARG_LC_batch(params["n_all"], params["m_all"], params["v_all"],
                            params["u_all"], params["C_nm_s_ARG_vec"], params["C_vu_r_vec"],
                            im["R_sI_r_all"], im["atten_all"], k)
    if method == "ORG":
        W = params["Wigner"]
        return ARG_ORG_batch(N, V, params["C_nm_s_ARG"], params["C_vu_r"], im["atten_all"],
                             im["R_sI_r_all"], W["W_1_all"], W["W_2_all"], k)
    if method == "MIX":
        W = params["Wigner"]
        early, late = im["early_indices"], im["late_indices"]
        P = np.zeros(k.size, dtype=np.complex128)
        if len(early) > 0:
            P += ARG_ORG_batch(N, V, params["C_nm_s_ARG"][:, :, :, early], params["C_vu_r"],
                               im["atten_all"][:, early], im["R_sI_r_all"][:, early],
                               W["W_1_all"], W["W_2_all"], k)
        if len(late) > 0:
            P += ARG_LC_batch(params["n_all"], params["m_all"], params["v_all"],
                              params["u_all"], params["C_nm_s_ARG_vec"][:, :, late],
                              params["C_vu_r_vec"], im["R_sI_r_all"][:, late],
                              im["atten_all"][:, late], k)
        return P
    raise ValueError(f"Unknown DEISM method: {method}")


# --------------------------------------------------------------------------------------
# shoebox image generator (cd:3326-3473) and merge (cd:4266-4295)
# --------------------------------------------------------------------------------------
def shoebox_images(params):
    LL = _c(params["roomSize"], np.float64)
    x_r = _c(params["posReceiver"], np.float64)
    x_s = _c(params["posSource"], np.float64)
    max_d2 = float((params["soundSpeed"] * params["reverberationTime"]) ** 2)
    angdep = int(params["angDepFlag"])
    N_o = int(params["maxReflOrder"])
    Z = _c(params["impedance"], np.complex128)
    N_o_ORG = int(params["mixEarlyOrder"])
    compact = bool(params.get("shoeboxCompactImages", False))
    if N_o < N_o_ORG:
        N_o_ORG = N_o
    n_tasks = 8 * (N_o + 1)
    ec = np.zeros(n_tasks, dtype=np.int64)
    lc = np.zeros(n_tasks, dtype=np.int64)
    lib().oracle_count_shoebox_images(
        _p(LL, _f64p), _p(x_s, _f64p), _p(x_r, _f64p), C.c_double(max_d2),
        C.c_long(N_o), C.c_long(N_o_ORG), _p(ec, _i64p), _p(lc, _i64p))
    eo = np.zeros(n_tasks, dtype=np.int64)
    lo = np.zeros(n_tasks, dtype=np.int64)
    eo[1:] = np.cumsum(ec)[:-1]
    lo[1:] = np.cumsum(lc)[:-1]
    ne, nl, K = int(ec.sum()), int(lc.sum()), Z.shape[1]

    def geo(n):
        return [np.empty((n, 3), dtype=np.float32) for _ in range(3)]

    Ra_e, Rb_e, Rc_e = geo(ne)
    Ra_l, Rb_l, Rc_l = geo(nl)
    at_e = np.empty((ne, 0 if compact else K), dtype=np.complex64)
    at_l = np.empty((nl, 0 if compact else K), dtype=np.complex64)
    A_e = np.empty((ne, 6), dtype=np.int32)
    A_l = np.empty((nl, 6), dtype=np.int32)
    lib().oracle_fill_shoebox_images(
        _p(LL, _f64p), _p(x_s, _f64p), _p(x_r, _f64p), _p(Z, _f64p), C.c_long(K),
        C.c_double(max_d2), C.c_int(angdep), C.c_long(N_o), C.c_long(N_o_ORG),
        C.c_int(0 if compact else 1), _p(eo, _i64p), _p(lo, _i64p),
        _p(Ra_e, _f32p), _p(Rb_e, _f32p), _p(Rc_e, _f32p), _p(at_e, _f32p), _p(A_e, _i32p),
        _p(Ra_l, _f32p), _p(Rb_l, _f32p), _p(Rc_l, _f32p), _p(at_l, _f32p), _p(A_l, _i32p))
    if params["ifRemoveDirectPath"]:
        mask = np.all(A_e == 0, axis=1)
        if np.any(mask):
            i = np.where(mask)[0][0]
            A_e = np.delete(A_e, i, axis=0)
            Ra_e, Rb_e, Rc_e = (np.delete(x, i, axis=0) for x in (Ra_e, Rb_e, Rc_e))
            if not compact:
                at_e = np.delete(at_e, i, axis=0)
    images = {
        "storage": "compact" if compact else "materialized",
        "R_sI_r_all_early": Ra_e, "R_s_rI_all_early": Rb_e, "R_r_sI_all_early": Rc_e,
        "A_early": A_e,
        "R_sI_r_all_late": Ra_l, "R_s_rI_all_late": Rb_l, "R_r_sI_all_late": Rc_l,
        "A_late": A_l,
    }
    if not compact:
        images["atten_all_early"] = at_e
        images["atten_all_late"] = at_l
    return images


def merge_images(images):
    merged = {}
    if "storage" in images:
        merged["storage"] = images["storage"]
    for key in ("A", "R_sI_r_all", "R_s_rI_all", "R_r_sI_all"):
        merged[key] = np.concatenate([np.asarray(images[key + "_early"]),
                                      np.asarray(images[key + "_late"])], axis=0)
    if "atten_all_early" in images and "atten_all_late" in images:
        merged["atten_all"] = np.concatenate([images["atten_all_early"],
                                              images["atten_all_late"]], axis=0)
    return merged


# --------------------------------------------------------------------------------------
# Wigner tables (cd:1842-1913) — sympy exact arithmetic, as the reference
# --------------------------------------------------------------------------------------
def wigner_tables(N, V):
    from sympy.physics.wigner import wigner_3j

    W1 = np.zeros((N + 1, V + 1, N + V + 1))
    W2 = np.zeros((N + 1, V + 1, N + V + 1, 2 * N + 1, 2 * V + 1))
    for n in range(N + 1):
        for m in range(-n, n + 1):
            for v in range(V + 1):
                for u in range(-v, v + 1):
                    for l in range(abs(n - v), n + v + 1):
                        if abs(u - m) <= l:
                            W1[n, v, l] = float(wigner_3j(n, v, l, 0, 0, 0))
                            W2[n, v, l, m, u] = float(wigner_3j(n, v, l, -m, u, m - u))
    return {"W_1_all": W1.astype(np.complex64), "W_2_all": W2.astype(np.complex64)}


def monopole_coefficients(k):
    """cd:1254-1259 / 1323-1331: C = -i k j_0(0) conj(Y_0^0) = -i k / sqrt(4 pi), complex64,
    shape [K,1,1]."""
    from scipy.special import sph_harm_y, spherical_jn

    C0 = -1j * np.asarray(k) * spherical_jn(0, 0) * np.conj(sph_harm_y(0, 0, 0, 0))
    return C0[..., None, None].astype(np.complex64)


def vectorize(Cnm, order):
    """cd:1169-1242: j = n^2 + n + m (negative m wraps), complex64."""
    K = Cnm.shape[0]
    J = (order + 1) ** 2
    n_all = np.zeros(J, dtype=int)
    m_all = np.zeros(J, dtype=int)
    vec = np.zeros((K, J), dtype=complex)
    for n in range(order + 1):
        for m in range(-n, n + 1):
            n_all[n * n + n + m] = n
            m_all[n * n + n + m] = m
            vec[:, n * n + n + m] = Cnm[:, n, m]
    return n_all, m_all, vec.astype(np.complex64)


# --------------------------------------------------------------------------------------
# convex room (libroom_deism restatement, oracle/convex_oracle.c)
# --------------------------------------------------------------------------------------
def walls_from_libroom(walls, K=None):
    """Pack a list of libroom Wall_deism objects (or anything with .normal/.origin/.basis/
    .flat_corners/.reflection_matrix/.impedence_bands) into the flat float32 tables used by the
    oracle and by the C ABI (struct deism_convex_walls)."""
    W = len(walls)
    mc = max(np.asarray(w.flat_corners).shape[1] for w in walls)
    out = {
        "normal": np.zeros((W, 3), np.float32), "origin": np.zeros((W, 3), np.float32),
        "basis": np.zeros((W, 3, 2), np.float32), "flat_corners": np.zeros((W, 2, mc), np.float32),
        "n_corners": np.zeros(W, np.int32), "reflection": np.zeros((W, 3, 3), np.float32),
    }
    imp = []
    for i, w in enumerate(walls):
        out["normal"][i] = np.asarray(w.normal, np.float32).reshape(3)
        out["origin"][i] = np.asarray(w.origin, np.float32).reshape(3)
        out["basis"][i] = np.asarray(w.basis, np.float32)
        fc = np.asarray(w.flat_corners, np.float32)
        out["flat_corners"][i, :, :fc.shape[1]] = fc
        out["n_corners"][i] = fc.shape[1]
        out["reflection"][i] = np.asarray(w.reflection_matrix, np.float32)
        imp.append(np.real(np.asarray(w.impedence_bands)).astype(np.float32))   # wall.cpp:428
    out["impedance"] = np.ascontiguousarray(np.stack(imp, axis=0))
    return out


def walls_from_golden(g):
    """Same tables from a tests/golden convex fixture (walls/<i>/...)."""
    grp = g.group("walls")
    W = int(grp["n"])

    class _W:
        pass

    ws = []
    for i in range(W):
        w = _W()
        w.normal, w.origin = grp[f"{i}/normal"], grp[f"{i}/origin"]
        w.basis, w.flat_corners = grp[f"{i}/basis"], grp[f"{i}/flat_corners"]
        w.reflection_matrix, w.impedence_bands = grp[f"{i}/reflection_matrix"], grp[f"{i}/impedance"]
        ws.append(w)
    return walls_from_libroom(ws)


def convex_images(wt, source, mic, ism_order):
    """room.cpp:271-511: returns dict(sources[3,I], orders[I], gen_walls[I], attenuations[K,I],
    reflection_matrix[I,3,3], n_nodes)."""
    L = lib()
    L.oracle_convex_run.restype = C.c_void_p
    K = wt["impedance"].shape[1]
    src = _c(source, np.float32)
    m = _c(mic, np.float32)
    n, nodes = C.c_long(0), C.c_long(0)
    h = L.oracle_convex_run(
        C.c_int(wt["normal"].shape[0]), C.c_int(wt["flat_corners"].shape[2]),
        _p(wt["normal"], _f32p), _p(wt["origin"], _f32p), _p(wt["basis"], _f32p),
        _p(wt["flat_corners"], _f32p), _p(wt["n_corners"], _i32p), _p(wt["reflection"], _f32p),
        _p(wt["impedance"], _f32p), C.c_int(K), _p(src, _f32p), _p(m, _f32p), C.c_int(ism_order),
        C.byref(n), C.byref(nodes))
    if not h:
        raise MemoryError("oracle_convex_run")
    I = n.value
    out = {"sources": np.zeros((3, I), np.float32), "orders": np.zeros(I, np.int32),
           "gen_walls": np.zeros(I, np.int32), "attenuations": np.zeros((K, I), np.float32),
           "reflection_matrix": np.zeros((I, 3, 3), np.float32), "n_nodes": nodes.value}
    L.oracle_convex_export(C.c_void_p(h), _p(out["sources"], _f32p), _p(out["orders"], _i32p),
                           _p(out["gen_walls"], _i32p), _p(out["attenuations"], _f32p),
                           _p(out["reflection_matrix"], _f32p))
    L.oracle_convex_free(C.c_void_p(h))
    return out


def acosf(x):
    L = lib()
    L.oracle_acosf.restype = C.c_float
    return float(L.oracle_acosf(C.c_float(x)))


def cosf(x):
    L = lib()
    L.oracle_cosf.restype = C.c_float
    return float(L.oracle_cosf(C.c_float(x)))


# ------------------------------------------------------------------------------------------
# ARG source directivity refit per image (NumPy restatement; test infrastructure)
# ------------------------------------------------------------------------------------------
def SHCs_from_pressure_LS(Psh, Dir_all, order):
    """core_deism.py:1481-1509: fnm = pinv(Y) @ Psh.T with Y[s, n^2+n+m] = Y_n^m(az_s, polar_s).
    Returns fnm [J, K] (the reference reshapes it to [K, n, m+n])."""
    from scipy import special as sps

    J = (order + 1) ** 2
    Y = np.zeros((len(Dir_all), J), dtype=complex)
    for n in range(order + 1):
        for m in range(-n, n + 1):
            Y[:, n * n + n + m] = sps.sph_harm_y(n, m, Dir_all[:, 1], Dir_all[:, 0])
    return np.linalg.pinv(Y) @ np.asarray(Psh).T


def cal_C_nm_s_ARG_vec(reflection_matrix, Psh_source, src_Psh_coords, k, order, r0):
    """core_deism.py:1545-1605 followed by 1193-1217: complex64 [K, J, I].
    reflection_matrix [3, 3, I]; Psh_source [K, n_dir]; src_Psh_coords [3, n_dir]."""
    from scipy import special as sps

    k = np.asarray(k, dtype=np.float64)
    n_img = reflection_matrix.shape[2]
    J = (order + 1) ** 2
    out = np.zeros((k.size, J, n_img), dtype=complex)
    hn = np.stack([sps.spherical_jn(n, k * r0) - 1j * sps.spherical_yn(n, k * r0) for n in range(order + 1)])
    for i in range(n_img):
        xyz = reflection_matrix[:, :, i] @ src_Psh_coords                    # cd:1577
        hxy = np.hypot(xyz[0], xyz[1])
        az, el = np.arctan2(xyz[1], xyz[0]), np.arctan2(xyz[2], hxy)         # utilities.py:95-102
        fnm = SHCs_from_pressure_LS(Psh_source[: k.size], np.stack((az, np.pi / 2 - el), axis=1), order)
        for n in range(order + 1):
            for m in range(-n, n + 1):
                j = n * n + n + m
                out[:, j, i] = fnm[j] / hn[n]                                 # cd:1597-1604
    return out.astype(np.complex64)
