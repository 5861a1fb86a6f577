"""Host-side precompute consumed by the kernels, same names and ``params``
contract as the reference:

    vectorize_C_nm_s / vectorize_C_nm_s_ARG / vectorize_C_vu_r   ~ core_deism.py:1169-1242
    monopole branches of init_{source,receiver}_directivities[_ARG] ~ core_deism.py:1254-1263,
                                                       1323-1338, 1623-1640, 1730-1745

Sampled (non-monopole) directivities need the .mat data sets that are absent from the reference
tree (.MISSING_LARGE_BLOBS); callers inject ``C_nm_s`` / ``C_vu_r`` (complex64 [K, N+1, 2N+1],
signed m indexed NumPy-style) and call the vectorisers.
"""
from __future__ import annotations

import numpy as np

_Y00 = 0.28209479177387814      # conj(Y_0^0(0, 0)) = 1 / sqrt(4 pi); spherical_jn(0, 0) = 1


def _track(params, keys, who):
    if params.get("track_updated_where"):
        for key in keys:
            params["updated_where"][key] = [who]


def _flatten(C, order):
    """j = n^2 + n + m over the first axes [K, n, m(, I)]; negative m wraps like NumPy."""
    J = (order + 1) ** 2
    n_all = np.zeros(J, dtype="int")
    m_all = np.zeros(J, dtype="int")
    vec = np.zeros((C.shape[0], J) + C.shape[3:], dtype="complex")
    for n in range(order + 1):
        for m in range(-n, n + 1):
            j = n * n + n + m
            n_all[j], m_all[j] = n, m
            vec[:, j] = C[:, n, m]
    return n_all, m_all, vec.astype(np.complex64)


def vectorize_C_nm_s(params):
    params["n_all"], params["m_all"], params["C_nm_s_vec"] = _flatten(params["C_nm_s"], params["sourceOrder"])
    _track(params, ("C_nm_s_vec", "n_all", "m_all"), "vectorize_C_nm_s")
    return params


def vectorize_C_nm_s_ARG(params):
    params["n_all"], params["m_all"], params["C_nm_s_ARG_vec"] = _flatten(params["C_nm_s_ARG"],
                                                                        params["sourceOrder"])
    _track(params, ("C_nm_s_ARG_vec", "n_all", "m_all"), "vectorize_C_nm_s_ARG")
    return params


def vectorize_C_vu_r(params):
    params["v_all"], params["u_all"], params["C_vu_r_vec"] = _flatten(params["C_vu_r"], params["receiverOrder"])
    _track(params, ("C_vu_r_vec", "v_all", "u_all"), "vectorize_C_vu_r")
    return params


def _monopole(k):
    return (-1j * np.asarray(k, dtype=np.float64) * 1.0 * _Y00)


def init_source_directivities(params):
    if params["sourceType"] != "monopole":
        raise NotImplementedError("sampled source directivities need the absent .mat data; inject "
                                  "params['C_nm_s'] instead")
    params["C_nm_s"] = _monopole(params["waveNumbers"])[..., None, None].astype(np.complex64)
    params["sourceOrder"] = 0
    _track(params, ("C_nm_s", "sourceOrder"), "init_source_directivities")
    return params


def init_receiver_directivities(params):
    if params["receiverType"] != "monopole":
        raise NotImplementedError("sampled receiver directivities need the absent .mat data; inject "
                                  "params['C_vu_r'] instead")
    params["C_vu_r"] = _monopole(params["waveNumbers"])[..., None, None].astype(np.complex64)
    params["receiverOrder"] = 0
    params["ifReceiverNormalize"] = 0
    _track(params, ("C_vu_r", "receiverOrder", "ifReceiverNormalize"), "init_receiver_directivities")
    return params


def init_source_directivities_ARG(params):
    if params["sourceType"] != "monopole":
        raise NotImplementedError("sampled source directivities need the absent .mat data")
    n_img = params["reflection_matrix"].shape[2]
    c = _monopole(params["waveNumbers"])[..., None, None, None].astype(np.complex64)
    params["C_nm_s_ARG"] = c * np.ones((1, 1, 1, n_img)).astype(np.complex64)
    params["sourceOrder"] = 0
    _track(params, ("C_nm_s_ARG", "sourceOrder"), "init_source_directivities")
    return params


init_receiver_directivities_ARG = init_receiver_directivities
