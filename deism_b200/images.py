"""Shoebox image enumeration on the GPU: the drop-in for the reference's v2-numba generator.

    pre_calc_images_src_rec_optimized_nofs_v2_numba(params)  ~ core_deism.py:3326-3473
    merge_images(images)                                     ~ core_deism.py:4266-4295

Returns the same dict (keys ``storage``, ``A_{early,late}`` int32[.,6],
``R_sI_r_all_*``/``R_s_rI_all_*``/``R_r_sI_all_*`` float32[.,3]) in the reference's order
(bit-equal lists).  By default the [I,K] complex64 attenuation is NOT materialised
(``storage == "fused"``): the accumulation kernels rebuild it in registers with the same
numerics.  ``params["shoeboxMaterializeAtten"] = True`` (or ``materialize=True``) adds the
reference's ``atten_all_{early,late}`` arrays; ``params["shoeboxCompactImages"] = True`` gives
the reference's "compact" storage.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import ATTEN_FUSED_ROOM, ShoeboxAtten, check, lib
from .backend import _device, _ptr, _stream_ptr, _vec3, to_device


def _f64x3(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(3))


def shoebox_images_device(params, dev=None):
    """Runs count + fill on the device; returns (dict of device tensors, n_early, n_late)."""
    dev = dev or _device()
    LL, x_s, x_r = _f64x3(params["roomSize"]), _f64x3(params["posSource"]), _f64x3(params["posReceiver"])
    max_d2 = float((params["soundSpeed"] * params["reverberationTime"]) ** 2)   # cd:3348
    N_o = int(params["maxReflOrder"])
    N_o_ORG = min(int(params["mixEarlyOrder"]), N_o)                            # cd:3360-3361
    remove_direct = 1 if params["ifRemoveDirectPath"] else 0
    n_tasks = 8 * (N_o + 1)
    counts = torch.zeros((2, n_tasks), dtype=torch.int64, device=dev)
    st = _stream_ptr()
    check(lib().deism_shoebox_count_images(LL.ctypes.data, x_s.ctypes.data, x_r.ctypes.data, max_d2,
                                           N_o, N_o_ORG, remove_direct, _ptr(counts[0]),
                                           _ptr(counts[1]), st), "deism_shoebox_count_images")
    totals = counts.sum(dim=1).cpu()          # the one host sync: exact-size outputs
    n_e, n_l = int(totals[0]), int(totals[1])
    n = n_e + n_l
    # early and late are written back to back, so the ORG/LC "merged" lists (cd:161-162) are
    # the same buffers without a copy.
    A = torch.empty((n, 6), dtype=torch.int32, device=dev)
    Ra = torch.empty((n, 3), dtype=torch.float32, device=dev)
    Rb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    Rc = torch.empty((n, 3), dtype=torch.float32, device=dev)
    check(lib().deism_shoebox_fill_images(
        LL.ctypes.data, x_s.ctypes.data, x_r.ctypes.data, max_d2, N_o, N_o_ORG, remove_direct,
        _ptr(counts[0]), _ptr(counts[1]),
        _ptr(A[:n_e]), _ptr(Ra[:n_e]), _ptr(Rb[:n_e]), _ptr(Rc[:n_e]),
        _ptr(A[n_e:]), _ptr(Ra[n_e:]), _ptr(Rb[n_e:]), _ptr(Rc[n_e:]), st),
        "deism_shoebox_fill_images")
    return {"A": A, "R_sI_r_all": Ra, "R_s_rI_all": Rb, "R_r_sI_all": Rc}, n_e, n_l


def shoebox_attenuation_device(params, A, R_sI_r=None, from_angles=False, c128=False, dev=None):
    """atten[I,K] on the device — the reference's images['atten_all*'] (cd:2900-2928) or its
    compact-storage rebuild (pb:140-186) when ``from_angles``."""
    dev = dev or _device()
    A = to_device(A, torch.int32, dev)
    Z = to_device(np.asarray(params["impedance"], dtype=np.complex128), torch.complex128, dev)
    n, K = int(A.shape[0]), int(Z.shape[1])
    s = ShoeboxAtten()
    s.mode = 3 if from_angles else ATTEN_FUSED_ROOM
    s.ang_dep_flag = int(params["angDepFlag"])
    s.d_A, s.d_Z_S = A.data_ptr(), Z.data_ptr()
    R = None
    if from_angles:
        R = to_device(R_sI_r, torch.float32, dev)
        s.d_R_sI_r = R.data_ptr()
    s.room_size, s.pos_source = _vec3(params["roomSize"]), _vec3(params["posSource"])
    s.pos_receiver = _vec3(params["posReceiver"])
    out = torch.empty((n, K), dtype=torch.complex128 if c128 else torch.complex64, device=dev)
    check(lib().deism_shoebox_attenuation(n, K, C.byref(s), _ptr(out), 1 if c128 else 0,
                                          _stream_ptr()), "deism_shoebox_attenuation")
    return out


def pre_calc_images_src_rec_optimized_nofs_v2_numba(params, materialize=None):
    """Drop-in for core_deism.py:3326 (same name on purpose).  NumPy arrays out."""
    dev_im, n_e, n_l = shoebox_images_device(params)
    compact = bool(params.get("shoeboxCompactImages", False))               # cd:2931-2942
    if materialize is None:
        materialize = bool(params.get("shoeboxMaterializeAtten", False))
    host = {k: v.cpu().numpy() for k, v in dev_im.items()}
    images = {"storage": "compact" if compact else ("materialized" if materialize else "fused")}
    for key in ("R_sI_r_all", "R_s_rI_all", "R_r_sI_all", "A"):
        images[key + "_early"] = host[key][:n_e]
        images[key + "_late"] = host[key][n_e:]
    if materialize and not compact:
        att = shoebox_attenuation_device(params, dev_im["A"]).cpu().numpy()
        images["atten_all_early"] = att[:n_e]
        images["atten_all_late"] = att[n_e:]
    return images


shoebox_images = pre_calc_images_src_rec_optimized_nofs_v2_numba


def merge_images(images):
    """core_deism.py:4266-4295: early || late for the ORG / LC methods."""
    merged = {}
    if "storage" in images:
        merged["storage"] = images["storage"]
    for key in ("A", "R_sI_r_all", "R_s_rI_all", "R_r_sI_all"):
        early, late = images[key + "_early"], images[key + "_late"]
        if isinstance(early, torch.Tensor):          # device-resident lists stay on the device
            merged[key] = torch.cat([early, late], dim=0)
            continue
        merged[key] = np.concatenate([np.asarray(early), np.asarray(late)], axis=0)
    if "atten_all_early" in images and "atten_all_late" in images:
        merged["atten_all"] = np.concatenate([images["atten_all_early"],
                                              images["atten_all_late"]], axis=0)
    return merged
