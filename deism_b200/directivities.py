"""Host-side precompute consumed by the kernels, same names and ``params``
contract as the reference:

    vectorize_C_nm_s / vectorize_C_nm_s_ARG / vectorize_C_vu_r   ~ core_deism.py:1169-1242
    monopole branches of init_{source,receiver}_directivities[_ARG] ~ core_deism.py:1254-1263,
                                                       1323-1338, 1623-1640, 1730-1745

    sampled (non-monopole) init_{source,receiver}_directivities: rotate_directions,
    SHCs_from_pressure_LS, get_directivity_coefs                    ~ core_deism.py:1264-1308, 1340-1397,
                                                       1481-1542 (one least-squares fit: host NumPy)
    cal_C_nm_s_new (+ SHCs_from_pressure_LS, vectorize_C_nm_s_ARG fused)   ~ core_deism.py:1545-1605,
                                                       1481-1509, 1193-1217   [GPU: deism_arg_refit]
    rotation_matrix_ZXZ                                                  ~ shared_utils.py:9-26

Sampled (non-monopole) directivities are read from the ``.mat`` profile ``sourceType`` /
``receiverType`` names (``data_loader.load_directive_pressure``, the reference's lookup rule; the data
sets themselves are absent from the reference tree, .MISSING_LARGE_BLOBS), unless the caller injects
the sampled field as ``params["sampledSource"]`` / ``params["sampledReceiver"]`` = ``{"freqs",
"Psh", "Dir_all", "r0"}`` (what the loader returns), or ready coefficients ``C_nm_s`` / ``C_vu_r``
(complex64 [K, N+1, 2N+1], signed m indexed NumPy-style) followed by the vectorisers.
"""
from __future__ import annotations

import numpy as np

_Y00 = 0.28209479177387814      # conj(Y_0^0(0, 0)) = 1 / sqrt(4 pi); spherical_jn(0, 0) = 1


def _track(params, keys, who):
    if params.get("track_updated_where"):
        for key in keys:
            params["updated_where"][key] = [who]


def _flatten(C, order):
    """j = n^2 + n + m over the first axes [K, n, m(, I)]; negative m wraps like NumPy.  Torch
    tensors (device-resident coefficients) stay torch tensors on their device."""
    J = (order + 1) ** 2
    n_all = np.zeros(J, dtype="int")
    m_all = np.zeros(J, dtype="int")
    for n in range(order + 1):
        for m in range(-n, n + 1):
            n_all[n * n + n + m], m_all[n * n + n + m] = n, m
    if hasattr(C, "device") and not isinstance(C, np.ndarray):
        import torch

        vec = torch.stack([C[:, int(n), int(m)] for n, m in zip(n_all, m_all)], dim=1)
        return n_all, m_all, vec.to(torch.complex64).contiguous()
    if isinstance(C, np.ndarray) and C.ndim == 4 and C.shape[3] > 1 and C.strides[3] == 0:
        # per-image coefficients that are one column broadcast over the images (monopole source,
        # cd:1623-1640): flatten the column and keep the broadcast, no [K, J, I] copy
        _, _, base = _flatten(np.ascontiguousarray(C[..., :1]), order)
        return n_all, m_all, np.broadcast_to(base, base.shape[:2] + (C.shape[3],))
    # (the reference fills a complex128 array and rounds it to complex64; complex64 input gives the
    # same values without the double-width temporary)
    vec = np.zeros((C.shape[0], J) + C.shape[3:], dtype=np.complex64 if C.dtype == np.complex64 else "complex")
    for j in range(J):
        vec[:, j] = C[:, n_all[j], m_all[j]]
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


def rotate_directions(Dir_all, facing):
    """Sampling directions [azimuth, inclination] after rotating +x to the facing direction given as
    ZXZ Euler angles in degrees (cd:1512-1531)."""
    alpha, beta, gamma = np.asarray(facing, dtype=float) * np.pi / 180
    el = np.pi / 2 - Dir_all[:, 1]
    xyz = np.vstack((np.cos(el) * np.cos(Dir_all[:, 0]), np.cos(el) * np.sin(Dir_all[:, 0]), np.sin(el)))
    r = rotation_matrix_ZXZ(alpha, beta, gamma) @ xyz
    hxy = np.hypot(r[0], r[1])
    return np.hstack((np.arctan2(r[1], r[0])[:, None], (np.pi / 2 - np.arctan2(r[2], hxy))[:, None]))


def SHCs_from_pressure_LS(Psh, Dir_all, sph_order_FEM, freqs_all):
    """Least-squares spherical-harmonic coefficients of a field sampled on a sphere (cd:1481-1509):
    [K, N+1, 2N+1] with mode m stored at index m + n."""
    from scipy import special as sps

    N = int(sph_order_FEM)
    Y = np.zeros((len(Dir_all), (N + 1) ** 2), dtype=complex)
    for n in range(N + 1):
        for m in range(-n, n + 1):
            Y[:, n * n + n + m] = sps.sph_harm_y(n, m, Dir_all[:, 1], Dir_all[:, 0])
    fnm = np.linalg.pinv(Y) @ np.asarray(Psh).T
    out = np.zeros((np.asarray(freqs_all).size, N + 1, 2 * N + 1), dtype=complex)
    for n in range(N + 1):
        for m in range(-n, n + 1):
            out[:, n, m + n] = fnm[n * n + n + m, :]
    return out


def get_directivity_coefs(k, maxSHorder, Pmnr0, r0):
    """C[k, n, m] = P_mn(r0) / h_n^(2)(k r0), signed m wrapping NumPy-style (cd:1534-1542)."""
    from scipy import special as sps

    k = np.asarray(k, dtype=float)
    C_ = np.zeros((k.size, maxSHorder + 1, 2 * maxSHorder + 1), dtype=complex)
    for n in range(maxSHorder + 1):
        hn = sps.spherical_jn(n, k * r0) - 1j * sps.spherical_yn(n, k * r0)
        for m in range(-n, n + 1):
            C_[:, n, m] = Pmnr0[:, n, m + n] / hn
    return C_


def _sampled(params, which):
    """(Psh, Dir_all) of the sampled field: ``params['sampledSource' | 'sampledReceiver']`` =
    ``{freqs, Psh, Dir_all, r0}`` when the caller injected it, else the reference's own route —
    ``load_directive_pressure`` (data_loader.py:1099-1145) on the ``.mat`` profile named by
    ``sourceType`` / ``receiverType`` — with the reference's checks (cd:1269-1283): a radius that
    differs from ``radiusSource`` / ``radiusReceiver`` only warns, different frequencies raise."""
    smp = params.get("sampled" + which)
    if smp is None:
        from .data_loader import load_directive_pressure

        freqs, Psh, Dir_all, r0 = load_directive_pressure(params.get("silentMode", 1), which.lower(),
                                                          params[which.lower() + "Type"])
        smp = {"freqs": freqs, "Psh": Psh, "Dir_all": Dir_all, "r0": r0}
    r0 = np.asarray(smp.get("r0", params.get("radius" + which, 0.0)), dtype=float).ravel()
    if r0.size and "radius" + which in params and np.abs(r0[0] - params["radius" + which]) > 1e-3:
        print(f"Warning: The radius of the {which.lower()} is {r0[0]}, not the same as the one defined in "
              f"params['radius{which}']")
    if not np.allclose(smp["freqs"], params["freqs"]):
        raise ValueError("The frequencies in the directivity data are not the same as the ones "
                         "defined in params['freqs']")
    return np.asarray(smp["Psh"]), np.asarray(smp["Dir_all"], dtype=float)


def init_source_directivities(params):
    if params["sourceType"] != "monopole":                                   # cd:1264-1308
        Psh, Dir_all = _sampled(params, "Source")
        P = SHCs_from_pressure_LS(Psh, rotate_directions(Dir_all, params["orientSource"]),
                                  params["sourceOrder"], params["freqs"])
        params["C_nm_s"] = get_directivity_coefs(params["waveNumbers"], params["sourceOrder"], P,
                                                 params["radiusSource"]).astype(np.complex64)
        _track(params, ("C_nm_s",), "init_source_directivities")
        return params
    params["C_nm_s"] = _monopole(params["waveNumbers"])[..., None, None].astype(np.complex64)
    params["sourceOrder"] = 0
    _track(params, ("C_nm_s", "sourceOrder"), "init_source_directivities")
    return params


def init_receiver_directivities(params):
    if params["receiverType"] != "monopole":                                 # cd:1340-1397
        Psh, Dir_all = _sampled(params, "Receiver")
        if params["ifReceiverNormalize"]:
            Psh = Psh / np.asarray(params["pointSrcStrength"])[..., None]
        P = SHCs_from_pressure_LS(Psh, rotate_directions(Dir_all, params["orientReceiver"]),
                                  params["receiverOrder"], params["freqs"])
        params["C_vu_r"] = get_directivity_coefs(params["waveNumbers"], params["receiverOrder"], P,
                                                 params["radiusReceiver"]).astype(np.complex64)
        _track(params, ("C_vu_r",), "init_receiver_directivities")
        return params
    params["C_vu_r"] = _monopole(params["waveNumbers"])[..., None, None].astype(np.complex64)
    params["receiverOrder"] = 0
    params["ifReceiverNormalize"] = 0
    _track(params, ("C_vu_r", "receiverOrder", "ifReceiverNormalize"), "init_receiver_directivities")
    return params


def rotation_matrix_ZXZ(alpha, beta, gamma):
    """COMSOL's Z-X-Z Euler rotation (shared_utils.py:9-26)."""
    ca, sa, cb, sb, cg, sg = np.cos(alpha), np.sin(alpha), np.cos(beta), np.sin(beta), np.cos(gamma), np.sin(gamma)
    return np.array([[ca * cg - sa * cb * sg, -ca * sg - sa * cb * cg, sb * sa],
                     [sa * cg + ca * cb * sg, -sa * sg + ca * cb * cg, -sb * ca],
                     [sb * sg, sb * cg, cb]])


def cal_C_nm_s_ARG_vec_device(reflection_matrix, Psh_source, src_Psh_coords, params, device=None):
    """Reflected source directivity coefficients of every image, vectorised: complex64 torch tensor
    [K, (N+1)^2, I] on the GPU (= ``C_nm_s_ARG_vec`` of cd:1193-1217 fed by cd:1545-1605).

    reflection_matrix: [3, 3, I] (reference layout) or a torch/NumPy [I, 3, 3] float32 array;
    Psh_source: [>=K, n_dir] complex; src_Psh_coords: [3, n_dir] Cartesian sampling points."""
    import ctypes as C

    import torch

    from . import _lib
    from .backend import to_device

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    k = np.asarray(params["waveNumbers"], dtype=np.float64)
    K, N = int(k.size), int(params["sourceOrder"])
    # [3, 3, I] is the reference's layout (also assumed for the ambiguous 3 x 3 x 3); [I, 3, 3] the engine's
    R = reflection_matrix
    engine_layout = tuple(R.shape[-2:]) == (3, 3) and R.shape[0] != 3
    if isinstance(R, torch.Tensor):
        R = R.to(dev, torch.float32)
        R = R if engine_layout else R.permute(2, 0, 1)
    else:
        R = np.asarray(R, dtype=np.float32)
        R = to_device(np.ascontiguousarray(R if engine_layout else np.transpose(R, (2, 0, 1))), torch.float32, dev)
    R = R.contiguous()
    n_img = int(R.shape[0])
    X = to_device(np.ascontiguousarray(np.asarray(src_Psh_coords, dtype=np.float64)), torch.float64, dev)
    n_dir = int(X.shape[1])
    P = np.asarray(Psh_source)[:K, :] if not isinstance(Psh_source, torch.Tensor) else Psh_source[:K, :]
    if tuple(P.shape) != (K, n_dir):
        raise ValueError(f"Psh_source must provide [{K}, {n_dir}] samples, got {tuple(P.shape)}")
    P = to_device(np.ascontiguousarray(P, dtype=np.complex128) if not isinstance(P, torch.Tensor) else P,
                  torch.complex128, dev).contiguous()
    kd = to_device(k, torch.float64, dev)
    out = torch.empty((K, (N + 1) ** 2, n_img), dtype=torch.complex64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().deism_arg_refit(n_img, K, N, n_dir, R.data_ptr(), X.data_ptr(), P.data_ptr(),
                                             kd.data_ptr(), float(params["radiusSource"]), out.data_ptr(),
                                             C.c_void_p(st)), "deism_arg_refit")
    return out


def cal_C_nm_s_new(reflection_matrix, Psh_source, src_Psh_coords, params):
    """cd:1545-1605 with the reference's output layout [K, N+1, 2N+1, I] (signed m wraps NumPy-style).
    Values are the complex64 the reference rounds to right after the call (cd:1708)."""
    vec = cal_C_nm_s_ARG_vec_device(reflection_matrix, Psh_source, src_Psh_coords, params).cpu().numpy()
    N = int(params["sourceOrder"])
    out = np.zeros((vec.shape[0], N + 1, 2 * N + 1, vec.shape[2]), dtype=np.complex64)
    for n in range(N + 1):
        for m in range(-n, n + 1):
            out[:, n, m, :] = vec[:, n * n + n + m, :]
    return out


def init_source_directivities_ARG(params):
    if params["sourceType"] != "monopole":
        # cd:1641-1712: sampled field -> rotated sampling points -> per-image refit (on the GPU)
        Psh_src, Dir = _sampled(params, "Source")
        el = np.pi / 2 - Dir[:, 1]
        xyz = np.vstack((np.cos(el) * np.cos(Dir[:, 0]), np.cos(el) * np.sin(Dir[:, 0]), np.sin(el)))
        Rm = rotation_matrix_ZXZ(*[float(a) for a in params["orientSource"][:3]])
        if params.get("ifRotateRoom") == 1:
            rr = np.asarray(params["roomRotation"], dtype=np.float64) * np.pi / 180
            Rm = rotation_matrix_ZXZ(rr[0], rr[1], rr[2]) @ Rm
        params["C_nm_s_ARG"] = cal_C_nm_s_new(params["reflection_matrix"], Psh_src, Rm @ xyz, params)
        _track(params, ("C_nm_s_ARG",), "init_source_directivities")
        return params
    refl = params["reflection_matrix"]
    c = _monopole(params["waveNumbers"])[..., None, None, None].astype(np.complex64)
    if hasattr(refl, "device") and not isinstance(refl, np.ndarray):
        # device-resident pipeline (convex.get_ref_paths_ARG_device): [I, 3, 3] torch tensor
        import torch

        n_img = int(refl.shape[0])
        params["C_nm_s_ARG"] = torch.from_numpy(c).to(refl.device).expand(-1, 1, 1, n_img).contiguous()
    else:
        # the reference multiplies by ones((1, 1, 1, I)) (cd:1636-1639): the same values as a
        # read-only broadcast view, without K x I copies of one column (54 MB at C4)
        n_img = refl.shape[2]
        params["C_nm_s_ARG"] = np.broadcast_to(c, (c.shape[0], 1, 1, n_img))
    params["sourceOrder"] = 0
    _track(params, ("C_nm_s_ARG", "sourceOrder"), "init_source_directivities")
    return params


init_receiver_directivities_ARG = init_receiver_directivities
