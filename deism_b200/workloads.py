"""BASELINE workloads rebuilt from the scalars frozen in
tests/golden/*.npz (room, positions, T60, impedance, wave numbers — minted from the reference by
oracle/make_golden.py) plus seeded synthetic directivity coefficients.

Used by bench.py, __graft_entry__.smoke() and the full-size parity tests.  No reference code,
no oracle: host-side NumPy only.
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (fixture, method, N, V, description)
WORKLOADS = {
    "c1": ("c1_full", "MIX", 0, 0,
           "C1 shipped configSingleParam_RTF.yml: shoebox RTF MIX order 25, K=12000, monopole"),
    "c2": ("c2_full_N5", "LC", 5, 5,
           "C2 shoebox RIR DEISM-LC order 20, fs 48 kHz (K=8000), synthetic N=V=5"),
    "c2mono": ("c2_full_mono", "LC", 0, 0,
               "C2 shoebox RIR DEISM-LC order 20, fs 48 kHz (K=8000), monopole"),
    "c3": ("c3_sub100", "ORG", 10, 10,
           "C3 shoebox DEISM-ORG order 25, N=V=10, K=2000 (2:2:4000 Hz)"),
    "c5": ("c5_full_N5", "MIX", 5, 5,
           "C5 shoebox RIR DEISM-MIX order 40 (88641 images) x 16000 freqs, synthetic N=V=5"),
}


def synth_coeffs(K, N, V, seed=0):
    """rng = default_rng(seed); source first, then receiver; complex64."""
    rng = np.random.default_rng(seed)
    cs = (rng.standard_normal((K, N + 1, 2 * N + 1))
          + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
    cr = (rng.standard_normal((K, V + 1, 2 * V + 1))
          + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(np.complex64)
    return cs, cr


def monopole_coeffs(k):
    """core_deism.py:1254-1259: C = -i k j_0(0) conj(Y_0^0) = -i k / sqrt(4 pi), complex64."""
    c = (-1j * np.asarray(k, dtype=np.float64) * (1.0 / np.sqrt(4.0 * np.pi)))
    return c[:, None, None].astype(np.complex64)


def vectorize(C, order):
    """core_deism.py:1169-1242: j = n^2 + n + m, complex64."""
    K = C.shape[0]
    J = (order + 1) ** 2
    n_all = np.zeros(J, dtype=np.int64)
    m_all = np.zeros(J, dtype=np.int64)
    vec = np.zeros((K, J), dtype=np.complex64)
    for n in range(order + 1):
        for m in range(-n, n + 1):
            j = n * n + n + m
            n_all[j], m_all[j] = n, m
            vec[:, j] = C[:, n, m]
    return n_all, m_all, vec


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    p = {}
    for key in z.files:
        if key.startswith("params/"):
            v = z[key]
            p[key[7:]] = v.item() if v.ndim == 0 else v
    if "impedance" not in p:
        K = int(p["waveNumbers"].shape[0])
        p["impedance"] = np.ascontiguousarray(
            np.repeat(np.asarray(p["impedance_col0"]).reshape(-1, 1), K, axis=1))
    p.pop("impedance_col0", None)
    p["silentMode"] = 1
    p["ifRemoveDirectPath"] = int(p.get("ifRemoveDirectPath", 0))
    return p, z


def c3_full_wavenumbers(p):
    """C3 is frozen on every 20th frequency; the bench runs the full 2:2:4000 Hz grid."""
    freqs = np.arange(2.0, 4000.0 + 2.0, 2.0)
    return 2 * np.pi * freqs / float(p["soundSpeed"])


def build(name, full_c3=True, seed=0):
    """Returns (params without images, fixture npz, description)."""
    from . import wigner

    fixture, method, N, V, desc = WORKLOADS[name]
    if name == "c3" and full_c3 and os.path.exists(os.path.join(GOLDEN, "c3_full.npz")):
        fixture, full_c3 = "c3_full", False      # the reference's run on the whole 2:2:4000 Hz grid
    p, z = load_fixture(fixture)
    if name == "c3" and full_c3:
        k = c3_full_wavenumbers(p)
        p["waveNumbers"] = k
        p["impedance"] = np.ascontiguousarray(np.repeat(p["impedance"][:, :1], len(k), axis=1))
    k = np.asarray(p["waveNumbers"], dtype=np.float64)
    K = len(k)
    p["DEISM_method"] = method
    p["sourceOrder"], p["receiverOrder"] = N, V
    if N == 0 and V == 0:
        p["C_nm_s"], p["C_vu_r"] = monopole_coeffs(k), monopole_coeffs(k)
    else:
        p["C_nm_s"], p["C_vu_r"] = synth_coeffs(K, N, V, seed)
    if method in ("LC", "MIX"):
        p["n_all"], p["m_all"], p["C_nm_s_vec"] = vectorize(p["C_nm_s"], N)
        p["v_all"], p["u_all"], p["C_vu_r_vec"] = vectorize(p["C_vu_r"], V)
    if method in ("ORG", "MIX"):
        p = wigner.pre_calc_Wigner(p)
    return p, z, desc


def slice_frequencies(p, lo, hi):
    """Contiguous frequency slab [lo, hi) of a workload (multi-GPU sharding)."""
    q = dict(p)
    for key in ("waveNumbers", "freqs"):
        if key in q and q[key] is not None:
            q[key] = np.ascontiguousarray(np.asarray(q[key])[lo:hi])
    q["impedance"] = np.ascontiguousarray(np.asarray(p["impedance"])[:, lo:hi])
    for key in ("C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
        if key in q:
            q[key] = np.ascontiguousarray(np.asarray(q[key])[lo:hi])
    return q


def n_images(images):
    if "A" in images:
        return len(images["A"])
    return len(images["A_early"]) + len(images["A_late"])
