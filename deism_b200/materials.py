"""Wall-material pipeline feeding ``Z_S[walls, K]``, same names and argument meaning
as the reference:

    convert_imp_abs_t60_shoebox   ~ core_deism.py:807-848
    convert_imp_to_abs            ~ core_deism.py:1099-1123   (Paris formula, Kuttruff Eq. 2.54)
    convert_imp_to_t60            ~ core_deism.py:1060-1096   (Badeau 2024, Eq. 124)
    convert_abs_to_imp            ~ core_deism.py:851-909     (inverse of the 200-point Paris integral)
    convert_t60_to_imp            ~ core_deism.py:964-1025    (inverse of Badeau's T60)
    interpolate_functions         ~ core_deism.py:1126-1163   (PCHIP on the K grid, endpoints held)
    load_format_materials_checks  ~ data_loader.py:178-208
    update_wall_materials / update_freqs (function form on a params dict) ~ core_deism.py:257-466
    read_choras_json / write_choras_result ~ examples/DeismInterface.py:113-260 (wire format only)

The two inverse conversions are scalar bounded least-squares fits in the reference
(``scipy.optimize.least_squares`` on a percentage error); the same solver, objective, start values
and bounds are used here so that the fitted impedance is the reference's to the last bit — a
different root finder would land ~1e-7 away and move every reflection factor.

``interpolate_functions_device`` evaluates the same PCHIP interpolant with a CUDA kernel
(``deism_pchip_interp``) straight into the device-resident complex128 ``Z_S[walls, K]`` the
accumulation kernels read.
"""
from __future__ import annotations

import json

import numpy as np

_LN10_24 = 24.0 * np.log(10.0)


# ----------------------------------------------------------------------------------------------
# forward conversions (closed form)
# ----------------------------------------------------------------------------------------------
def convert_imp_to_abs(zeta):
    """Random-incidence absorption coefficient of a locally reacting wall (cd:1099-1123)."""
    z = np.asarray(zeta) + 1e-16j
    zr, zi = np.real(z), np.imag(z)
    m2 = np.abs(z) ** 2
    bracket = (1 + (zr ** 2 - zi ** 2) / (zi * m2) * np.arctan(zi / (1 + zr))
               - zr / m2 * np.log(1 + 2 * zr + m2))
    return 8 * zr / m2 * bracket


def _badeau_d(beta):
    """Wall term d_s of Badeau's reverberation formula; beta = 1/zeta (cd:1038-1044, 1075-1083)."""
    r1 = (1 + beta) / (1 - beta)
    r2 = (beta + 1) / (beta - 1)
    return np.log(np.abs(r1) ** 2) + 2 * np.real(beta * (2 - beta * np.log(np.abs(r2))))


def convert_imp_to_t60(Volumn, Areas, c, zeta):
    """T60[num_freqs] from zeta[6, num_freqs] (cd:1060-1096)."""
    zeta = np.asarray(zeta) + 1e-16j
    d = _badeau_d(1 / zeta)
    d = np.where(np.isfinite(d), d, 1e-6)
    d = np.where(d > 0, d, 1e-6)
    integrals = np.zeros((zeta.shape[1],))
    for w in range(zeta.shape[0]):                      # same accumulation order as the reference
        integrals += d[w, :] * Areas[w]
    return _LN10_24 * Volumn / c / integrals


# ----------------------------------------------------------------------------------------------
# inverse conversions (bounded scalar fits, solver settings of cd:880-909 / 994-1025)
# ----------------------------------------------------------------------------------------------
_THETA = np.linspace(0, np.pi / 2, 200)


def _paris_error(z, target):
    """Percentage error of the 200-point Paris integral against `target` (cd:943-961)."""
    from scipy.integrate import trapezoid

    ct = np.cos(_THETA)
    est = trapezoid(4 * z * ct * np.sin(2 * _THETA) / (z ** 2 * ct ** 2 + 2 * z * ct + 1), _THETA)
    if not np.isfinite(est) or abs(target) < 1e-10:
        return 1e6
    return np.abs(target - est) / target * 100


def get_imp_abs(z, abs_coeff):
    if np.isscalar(z) and np.isscalar(abs_coeff):
        return _paris_error(z, abs_coeff)
    z, abs_coeff = np.asarray(z), np.asarray(abs_coeff)
    if z.shape != abs_coeff.shape:
        raise ValueError("z and abs_coeff must have the same shape")
    out = np.zeros_like(z)
    for idx in np.ndindex(z.shape):
        out[idx] = _paris_error(z[idx], abs_coeff[idx])
    return out


def estimate_imp_t60(z, ref, V, S, c):
    """Percentage error of Badeau's T60 for a uniform wall impedance z (cd:1028-1057)."""
    d = _badeau_d(1 / z)
    if not np.isfinite(d) or d <= 0:
        return 1e6
    return np.abs(ref - _LN10_24 * V / c / S / d) / ref * 100


_paris_grid_cache = {}


def _paris_error_grid(grid, target):
    """_paris_error for a whole vector of impedances at once: the same element-wise operations and the
    same pairwise row sums as the scalar calls (bit-equal values, hence the same argmin).  The Paris
    integral does not depend on the target, so it is evaluated once per grid and kept."""
    from scipy.integrate import trapezoid

    grid = np.asarray(grid, dtype=float)
    key = (grid.shape[0], float(grid[0]), float(grid[-1]))
    est = _paris_grid_cache.get(key)
    if est is None:
        z = grid[:, None]
        ct = np.cos(_THETA)[None, :]
        est = trapezoid(4 * z * ct * np.sin(2 * _THETA)[None, :] / (z ** 2 * ct ** 2 + 2 * z * ct + 1), _THETA, axis=1)
        _paris_grid_cache[key] = est
    if abs(target) < 1e-10:
        return np.full(grid.shape[0], 1e6)
    return np.where(np.isfinite(est), np.abs(target - est) / target * 100, 1e6)


def _fit_impedance(objective, grid_objective=None):
    from scipy.optimize import least_squares

    result = None
    for x0 in (10, 5, 20, 1.5):
        try:
            result = least_squares(objective, x0=x0, bounds=(1, 1e3))
            if result.success:
                break
        except (ValueError, RuntimeError):
            continue
    if result is None or not result.success:
        grid = np.linspace(1, 1000, 1000)
        errors = grid_objective(grid) if grid_objective is not None else [objective(z) for z in grid]
        return grid[int(np.argmin(errors))] + 1e-16j
    return result.x[0] + 1e-16j


def _abs_to_imp_scalar(t):
    """cd:880-909.  least_squares hands the objective a length-1 array, for which get_imp_abs raises
    (shape mismatch with the scalar target) — in the reference as here — so every value ends in the
    1000-point grid search; that search is evaluated in one vectorised pass."""
    return _fit_impedance(lambda z: get_imp_abs(z, t), lambda grid: _paris_error_grid(grid, t))


def convert_abs_to_imp(abs_coeff):
    if np.isscalar(abs_coeff):
        return _abs_to_imp_scalar(abs_coeff)
    a = np.asarray(abs_coeff)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    imp = np.zeros_like(a, dtype=complex)
    done = {}                                  # band tables repeat values across walls
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            key = float(a[i, j])
            if key not in done:
                done[key] = _abs_to_imp_scalar(a[i, j])
            imp[i, j] = done[key]
    return imp


def convert_t60_to_imp(Volumn, Areas, c, t60):
    S = np.sum(Areas)

    def one(t):
        return _fit_impedance(lambda z: estimate_imp_t60(z, t, Volumn, S, c))

    if np.isscalar(t60):
        return one(t60)
    t60 = np.asarray(t60)
    if t60.ndim == 0:
        return one(float(t60))
    imp = np.zeros((6, t60.shape[0]), dtype=complex)
    for i in range(t60.shape[0]):
        imp[:, i] = one(t60[i])
    return imp


def convert_imp_abs_t60_shoebox(Volumn, Areas, c, datain, params_type):
    """(impedance, absorption, max T60) from any one of the three (cd:807-848)."""
    if params_type == "impedance":
        imp = datain
        t60 = convert_imp_to_t60(Volumn, Areas, c, imp)
        abs_coeff = convert_imp_to_abs(imp)
    elif params_type == "absorption":
        abs_coeff = datain
        imp = convert_abs_to_imp(abs_coeff)
        t60 = convert_imp_to_t60(Volumn, Areas, c, imp)
    elif params_type == "reverberationTime":
        t60 = datain
        imp = convert_t60_to_imp(Volumn, Areas, c, t60)
        abs_coeff = convert_imp_to_abs(imp)
        imp = np.full((6, 1), imp)
        abs_coeff = np.full((6, 1), abs_coeff)
    else:
        raise ValueError("The parameter type is not supported")
    return imp, abs_coeff, np.max(t60)


# ----------------------------------------------------------------------------------------------
# interpolation onto the K grid
# ----------------------------------------------------------------------------------------------
def interpolate_functions(datain, sparse_freqs, dense_freqs):
    """PCHIP per row on the dense grid; values held outside the measured band (cd:1126-1163)."""
    from scipy.interpolate import PchipInterpolator

    datain = np.asarray(datain)
    if datain.shape[-1] != len(sparse_freqs):
        raise ValueError("The last dimension of datain is not the same as the length of sparse_freqs")
    if len(sparse_freqs) < 2:
        return np.broadcast_to(datain, datain.shape[:-1] + (len(dense_freqs),)).copy()
    x = np.clip(dense_freqs, sparse_freqs[0], sparse_freqs[-1])
    re = PchipInterpolator(sparse_freqs, datain.real, axis=-1, extrapolate=False)
    im = PchipInterpolator(sparse_freqs, datain.imag, axis=-1, extrapolate=False)
    return re(x) + 1j * im(x)


def interpolate_functions_device(datain, sparse_freqs, dense_freqs, device=None):
    """Same interpolant evaluated on the GPU: returns a complex128 torch tensor [rows, K] (2-D input)
    that can be passed as ``params['impedance']`` without a host round trip."""
    import ctypes as C

    import torch

    from . import _lib
    from .backend import to_device

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    data = np.ascontiguousarray(np.atleast_2d(np.asarray(datain)), dtype=np.complex128)
    xs = np.ascontiguousarray(sparse_freqs, dtype=np.float64)
    if data.shape[-1] != xs.size:
        raise ValueError("The last dimension of datain is not the same as the length of sparse_freqs")
    xd = dense_freqs if isinstance(dense_freqs, torch.Tensor) else np.ascontiguousarray(dense_freqs, dtype=np.float64)
    xd = to_device(xd, torch.float64, dev).contiguous()
    d_data = to_device(data, torch.complex128, dev)
    d_xs = to_device(xs, torch.float64, dev)
    out = torch.empty((data.shape[0], xd.numel()), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().deism_pchip_interp(data.shape[0], xs.size, xd.numel(), d_xs.data_ptr(),
                                                d_data.data_ptr(), xd.data_ptr(), out.data_ptr(),
                                                C.c_void_p(st)), "deism_pchip_interp")
    return out


# ----------------------------------------------------------------------------------------------
# params-dict drivers (function form of DEISM.update_wall_materials / update_freqs)
# ----------------------------------------------------------------------------------------------
def load_format_materials_checks(datain, datatype):
    if datatype in ("impedance", "absorption"):
        if isinstance(datain, (int, float)):
            return np.full((6, 1), datain)
        if isinstance(datain, list):
            return np.array(datain)[:, None]
        raise ValueError(f"Invalid data shape for {datatype}!")
    if datatype == "reverberationTime":
        if isinstance(datain, (int, float)):
            return np.full((1, 1), datain)
        raise ValueError(f"Invalid data shape for {datatype}!")
    return None


def update_wall_materials(params, datain=None, freqs_bands=None, datatype=None, roomtype="shoebox"):
    """cd:257-406 on a params dict: fills impedance / absorption / reverberationTime /
    freqs_bands (and n1, n2, n3 for a shoebox, data_loader.py:211-237)."""
    if datain is not None and datatype is not None:
        if freqs_bands is None:
            freqs_bands = np.array([1000])
        params["givenMaterials"] = [datatype]
        if datatype == "reverberationTime" and roomtype == "convex":
            raise ValueError("T60 input is not supported for convex room yet, please use impedance or "
                             "absorption coefficients instead")
    else:
        datatype = params["givenMaterials"][0]
        freqs_bands = np.array([1000])
        if datatype not in ("impedance", "absorption", "reverberationTime"):
            raise ValueError("The parameter type is not supported")
        if datatype == "reverberationTime" and roomtype == "convex":
            raise ValueError("T60 input is not supported for convex room")
        datain = load_format_materials_checks(params[datatype], datatype)
    imp, abs_coeff, t60 = convert_imp_abs_t60_shoebox(params["roomVolumn"], params["roomAreas"],
                                                      params["soundSpeed"], datain, datatype)
    params["impedance"], params["absorption"], params["reverberationTime"] = imp, abs_coeff, t60
    params["freqs_bands"] = freqs_bands
    if roomtype == "shoebox":
        for key, L in zip(("n1", "n2", "n3"), params["roomSize"]):
            params[key] = int(np.floor(params["soundSpeed"] * params["reverberationTime"] / L / 2))
    return params


def update_freqs(params, device_impedance=False):
    """Frequency grid and wave numbers (cd:407-433) and the impedance interpolated onto it
    (cd:452-455).  device_impedance=True keeps Z_S on the GPU (interpolate_functions_device)."""
    if params["mode"] == "RIR":
        params["nSamples"] = int(params["sampleRate"] * params["RIRLength"])
        n_steps = np.ceil((params["sampleRate"] / 2) / (1 / params["reverberationTime"]))
        f_step = (params["sampleRate"] / 2) / n_steps
        params["freqs"] = np.arange(f_step, params["sampleRate"] / 2 + f_step, f_step)
    elif params["mode"] == "RTF":
        params["freqs"] = np.arange(params["startFreq"], params["endFreq"] + params["freqStep"],
                                    params["freqStep"])
    params["waveNumbers"] = 2 * np.pi * params["freqs"] / params["soundSpeed"]
    if params.get("ifReceiverNormalize") == 1:
        params["pointSrcStrength"] = (1j * params["waveNumbers"] * params["soundSpeed"]
                                      * params["airDensity"] * params["qFlowStrength"])
    interp = interpolate_functions_device if device_impedance else interpolate_functions
    params["impedance"] = interp(np.asarray(params["impedance"]), params["freqs_bands"], params["freqs"])
    return params


# ----------------------------------------------------------------------------------------------
# CHORAS JSON wire format (examples/exampleInput_Deism.json, DeismInterface.py:113-260)
# ----------------------------------------------------------------------------------------------
# DEISM's wall order x1, x2, y1, y2, z1, z2 = wall 1, wall 3, wall 2, wall 4, floor, ceiling
# (DeismInterface.py:201-203); with a .geo file the reference derives the names from the Gmsh
# physical-surface tags [2, 5, 4, 6, 1, 3] (DeismInterface.py:99-111), which needs gmsh.
CHORAS_WALL_ORDER = ("wall1", "wall3", "wall2", "wall4", "floor", "ceiling")


def _csv_floats(value):
    if isinstance(value, str):
        return np.array([float(t) for t in value.replace(";", ",").split(",") if t.strip()])
    return np.atleast_1d(np.asarray(value, dtype=float))


def read_choras_json(path, wall_order=CHORAS_WALL_ORDER):
    """Inputs of one CHORAS job in DEISM's conventions: absorption[6, bands], freq_bands,
    posSource, posReceiver, vertices[N,3], wallCenters[6,3], roomAreas[6,1], roomVolumn, and the
    `should_cancel` flag."""
    with open(path, "r", encoding="utf-8") as fh:
        job = json.load(fh)
    res = job["results"][0]
    geo = job["geometry"][0]
    bands = np.array(res["frequencies"], dtype=float)
    missing = [w for w in wall_order if w not in job["absorption_coefficients"]]
    if missing:
        raise KeyError(f"Missing surfaces required by DEISM: {missing}")
    absorption = np.zeros((6, bands.size))
    centers = np.zeros((6, 3))
    areas = np.zeros((6, 1))
    for i, wall in enumerate(wall_order):
        absorption[i, :] = _csv_floats(job["absorption_coefficients"][wall])
        centers[i, :] = _csv_floats(geo["wall_centers"][wall])
        areas[i, :] = _csv_floats(geo["room_areas"][wall])
    rec = res["responses"][0]
    return {"absorption": absorption, "freqs_bands": bands,
            "posSource": np.array([res["sourceX"], res["sourceY"], res["sourceZ"]], dtype=float),
            "posReceiver": np.array([rec["x"], rec["y"], rec["z"]], dtype=float),
            "vertices": np.array(geo["vertices"], dtype=float), "wallCenters": centers,
            "roomAreas": areas, "roomVolumn": float(geo["room_volume"]),
            "should_cancel": bool(job.get("should_cancel", False))}


def write_choras_result(path, rir):
    """Store the impulse response where the reference does (DeismInterface.py:245-248)."""
    with open(path, "r", encoding="utf-8") as fh:
        job = json.load(fh)
    job["results"][0]["responses"][0]["receiverResults"] = np.asarray(rir).tolist()
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(json.dumps(job, indent=4))
