#!/usr/bin/env python
"""oracle/make_golden.py — TEST INFRASTRUCTURE ONLY.

Mints golden input/output vectors from the UNMODIFIED reference (imported from
/root/reference via oracle/ref_import.py) and freezes them under tests/golden/.  The
reference cannot travel to the GPU box, so these fixtures are what the `-m gpu` parity tests
and the oracle pin (tests/test_oracle_golden.py) compare against.

    python oracle/make_golden.py            # the small cases (seconds each)
    python oracle/make_golden.py --full     # + full-size BASELINE configs C1..C5 (minutes;
                                            #   only RTF + hashes + scalars are stored)

Small cases store every array the reference's Numba kernels receive and the
complex128 RTF that run_DEISM returned.  Full-size cases store the scalars needed to rebuild
the inputs (room, positions, T60, impedance, wave numbers, seeds), the RTF and SHA-256
digests of the image lists.

Synthetic directivity coefficients (the real .mat inputs are absent, .MISSING_LARGE_BLOBS):
rng = default_rng(seed); source drawn first, then receiver; (N(0,1) + i N(0,1)).astype(c64),
shape [K, N+1, 2N+1]   (SURVEY.md §8d).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import ref_import  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def synth_coeffs(K, N, V, seed=0):
    rng = np.random.default_rng(seed)
    cs = (rng.standard_normal((K, N + 1, 2 * N + 1))
          + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
    cr = (rng.standard_normal((K, V + 1, 2 * V + 1))
          + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(np.complex64)
    return cs, cr


def sha(arr) -> str:
    a = np.ascontiguousarray(arr)
    return hashlib.sha256(a.tobytes()).hexdigest()


def inject_directivities(d, N, V, seed=0):
    """Replace update_directivities() by synthetic coefficients."""
    from deism.core_deism import pre_calc_Wigner, vectorize_C_nm_s, vectorize_C_vu_r

    p = d.params
    K = len(p["waveNumbers"])
    cs, cr = synth_coeffs(K, N, V, seed)
    p["sourceOrder"], p["receiverOrder"] = N, V
    p["C_nm_s"], p["C_vu_r"] = cs, cr
    if p["DEISM_method"] in ("LC", "MIX"):
        d.params = vectorize_C_nm_s(d.params)
        d.params = vectorize_C_vu_r(d.params)
    if p["DEISM_method"] in ("ORG", "MIX"):
        d.params = pre_calc_Wigner(d.params)


_SCALARS = ["soundSpeed", "reverberationTime", "angDepFlag", "maxReflOrder", "mixEarlyOrder",
            "ifRemoveDirectPath", "numParaImages", "sourceOrder", "receiverOrder",
            "sampleRate", "RIRLength"]
_ARRAYS = ["roomSize", "posSource", "posReceiver", "impedance", "waveNumbers", "freqs",
           "C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec", "n_all", "m_all", "v_all", "u_all",
           "vertices", "wallCenters", "C_nm_s_ARG", "C_nm_s_ARG_vec", "reflection_matrix"]


def small_inputs(p):
    """Inputs every case keeps (also the full-size ones): geometry, wave numbers, and the
    impedance (one column + a flag when it is frequency-flat)."""
    out = {}
    for key in ("roomSize", "posSource", "posReceiver", "vertices", "wallCenters",
                "waveNumbers", "freqs"):
        if key in p and p[key] is not None:
            out["params/" + key] = np.asarray(p[key])
    Z = np.asarray(p["impedance"])
    flat = bool(Z.ndim == 2 and np.all(Z == Z[:, :1]))
    out["meta/impedance_flat"] = np.array(flat)
    out["params/impedance_col0"] = np.ascontiguousarray(Z[:, 0]) if Z.ndim == 2 else Z
    if not flat:
        out["params/impedance"] = Z
    return out


def pack(d, extra=None, store_images=True, store_inputs=True):
    p = d.params
    out = {"RTF": np.asarray(p["RTF"]), "meta/method": np.array(p["DEISM_method"]),
           "meta/roomtype": np.array(d.roomtype), "meta/mode": np.array(d.mode)}
    for key in _SCALARS:
        if key in p and p[key] is not None:
            out["params/" + key] = np.asarray(p[key])
    out.update(small_inputs(p))
    if store_inputs:
        for key in _ARRAYS:
            if key in p and p[key] is not None:
                out["params/" + key] = np.asarray(p[key])
        if "Wigner" in p:
            out["wigner/W_1_all"] = p["Wigner"]["W_1_all"]
            out["wigner/W_2_all"] = p["Wigner"]["W_2_all"]
    if "images" in p:
        for key, val in p["images"].items():
            if key == "storage":
                out["images/storage"] = np.array(val)
            elif store_images:
                out["images/" + key] = np.asarray(val)
            else:
                a = np.asarray(val)
                out["imagehash/" + key] = np.array(sha(a))
                out["imageshape/" + key] = np.array(a.shape)
    if extra:
        out.update(extra)
    return out


def rir_of(d):
    """DEISM.get_results() (core_deism.py:675-769) without its side effect on params['RTF']."""
    if d.mode != "RIR":
        return None
    keep = np.array(d.params["RTF"], copy=True)
    p = d.get_results()
    d.params["RTF"] = keep
    return np.asarray(p)


def save(name, blob):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"  wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


# --------------------------------------------------------------------------------------
# shoebox
# --------------------------------------------------------------------------------------
def shoebox_case(name, mode, method, order, N=None, V=None, overrides=None, materials=None,
                 store_images=True, store_inputs=True, compact=False, seed=0, room=None):
    t0 = time.time()
    d = ref_import.make_deism(mode, "shoebox")
    p = d.params
    p["DEISM_method"] = method
    p["maxReflOrder"] = order
    for key, val in (overrides or {}).items():
        p[key] = val
    if room is not None:
        d.update_room(roomDimensions=np.asarray(room, dtype=float))
    if compact:
        p["shoeboxCompactImages"] = True
    if materials is None:
        d.update_wall_materials()
    else:
        d.update_wall_materials(*materials)
    d.update_freqs()
    if N is None:
        d.update_directivities()
    else:
        inject_directivities(d, N, V, seed)
    t1 = time.time()
    d.update_source_receiver()
    t2 = time.time()
    d.run_DEISM(if_clean_up=False)
    t3 = time.time()
    K = len(p["waveNumbers"])
    im = p["images"]
    nimg = len(im["A"]) if "A" in im else len(im["A_early"]) + len(im["A_late"])
    print(f"[{name}] {method} order {order}: I={nimg} K={K} T60={p['reverberationTime']:.6f} "
          f"images {t2 - t1:.2f}s run {t3 - t2:.2f}s (cold) setup {t1 - t0:.2f}s")
    extra = {"meta/n_images": np.array(nimg), "meta/seed": np.array(seed),
             "meta/t_images_s": np.array(t2 - t1), "meta/t_run_cold_s": np.array(t3 - t2)}
    rir = rir_of(d)
    if rir is not None:
        extra["RIR"] = rir
    save(name, pack(d, extra, store_images, store_inputs))
    return d


# --------------------------------------------------------------------------------------
# convex
# --------------------------------------------------------------------------------------
def convex_case(name, mode, method, order, N=None, overrides=None, store_inputs=True,
                seed=0):
    ref_import.import_reference()
    from deism.core_deism import (pre_calc_Wigner, vectorize_C_nm_s_ARG, vectorize_C_vu_r)

    t0 = time.time()
    d = ref_import.make_deism(mode, "convex")
    p = d.params
    p["DEISM_method"] = method
    p["maxReflOrder"] = order
    for key, val in (overrides or {}).items():
        p[key] = val
    d.update_wall_materials()
    d.update_freqs()
    t1 = time.time()
    d.update_source_receiver()
    t2 = time.time()
    if N is None:
        d.update_directivities()
    else:
        # synthetic per-image source coefficients C_nm_s_ARG[K, N+1, 2N+1, I] (cd:1608-1836
        # needs directivity .mat data that is absent); receiver shared C_vu_r[K, V+1, 2V+1]
        K = len(p["waveNumbers"])
        nimg = max(p["images"]["R_sI_r_all"].shape)
        rng = np.random.default_rng(seed)
        shp = (K, N + 1, 2 * N + 1, nimg)
        p["C_nm_s_ARG"] = (rng.standard_normal(shp)
                           + 1j * rng.standard_normal(shp)).astype(np.complex64)
        shp = (K, N + 1, 2 * N + 1)
        p["C_vu_r"] = (rng.standard_normal(shp)
                       + 1j * rng.standard_normal(shp)).astype(np.complex64)
        p["sourceOrder"] = p["receiverOrder"] = N
        if method in ("LC", "MIX"):
            d.params = vectorize_C_nm_s_ARG(d.params)
            d.params = vectorize_C_vu_r(d.params)
        if method in ("ORG", "MIX"):
            d.params = pre_calc_Wigner(d.params)
    d.run_DEISM(if_clean_up=False)
    t3 = time.time()
    eng = d.room_convex.room_engine
    walls = d.room_convex.walls
    extra = {
        "engine/sources": np.asarray(eng.sources), "engine/orders": np.asarray(eng.orders),
        "engine/gen_walls": np.asarray(eng.gen_walls),
        "engine/attenuations": np.asarray(eng.attenuations),
        "engine/reflection_matrix": np.asarray(eng.reflection_matrix, dtype=np.float32),
        "engine/visible_mics": np.asarray(eng.visible_mics),
        "walls/n": np.array(len(walls)),
        "meta/n_images": np.array(np.asarray(eng.sources).shape[1]),
        "meta/t_images_s": np.array(t2 - t1), "meta/t_run_cold_s": np.array(t3 - t2),
    }
    for i, w in enumerate(walls):
        lw = w.libroom_walls
        extra[f"walls/{i}/corners"] = np.asarray(lw.corners, dtype=np.float32)
        extra[f"walls/{i}/normal"] = np.asarray(lw.normal, dtype=np.float32)
        extra[f"walls/{i}/origin"] = np.asarray(lw.origin, dtype=np.float32)
        extra[f"walls/{i}/basis"] = np.asarray(lw.basis, dtype=np.float32)
        extra[f"walls/{i}/flat_corners"] = np.asarray(lw.flat_corners, dtype=np.float32)
        extra[f"walls/{i}/reflection_matrix"] = np.asarray(lw.reflection_matrix,
                                                           dtype=np.float32)
        extra[f"walls/{i}/impedance"] = np.asarray(w.impedence_bands)
    rir = rir_of(d)
    if rir is not None:
        extra["RIR"] = rir
    K = len(p["waveNumbers"])
    print(f"[{name}] convex {method} order {order}: I={extra['meta/n_images']} K={K} "
          f"T60={p['reverberationTime']:.6f} images {t2 - t1:.2f}s run {t3 - t2:.2f}s")
    save(name, pack(d, extra, True, store_inputs))
    return d


def small_cases():
    # S1: the reference's own smoke configuration (SURVEY §8c smoke values)
    shoebox_case("s1_shoebox_rtf_mix_o5_mono", "RTF", "MIX", 5, overrides={"endFreq": 200})
    # S2: LC, non-monopole, frequency-dependent complex impedance on 3 bands (PCHIP path)
    Zb = np.array([[18, 20, 8, 5, 30, 19], [12 + 3j, 25 - 2j, 9 + 1j, 6, 28 + 4j, 15 - 1j],
                   [10, 30, 7, 4, 35, 12]]).T
    shoebox_case("s2_shoebox_rtf_lc_o6_N3", "RTF", "LC", 6, N=3, V=3,
                 overrides={"startFreq": 20, "endFreq": 1280, "freqStep": 20},
                 materials=(Zb, np.array([100.0, 500.0, 1000.0]), "impedance"))
    # S3: ORG with asymmetric orders (N=3, V=2), angle-dependent beta
    shoebox_case("s3_shoebox_rtf_org_o4_N3V2", "RTF", "ORG", 4, N=3, V=2,
                 overrides={"startFreq": 50, "endFreq": 2000, "freqStep": 50})
    # S4: RIR mode, T60 input (C5-like), MIX with N=V=2, direct path removed, distance cut
    shoebox_case("s4_shoebox_rir_mix_o9_N2", "RIR", "MIX", 9, N=2, V=2,
                 overrides={"sampleRate": 1500, "ifRemoveDirectPath": 1},
                 materials=(0.08, None, "reverberationTime"))
    # S5: angle-independent beta, LC monopole, odd room
    shoebox_case("s5_shoebox_rtf_lc_o7_mono_noang", "RTF", "LC", 7,
                 overrides={"endFreq": 300, "angDepFlag": 0}, room=[5.3, 3.1, 2.7])
    # S6: compact image storage (beta from f32 angles, pb:140-186)
    shoebox_case("s6_shoebox_rtf_mix_o5_compact", "RTF", "MIX", 5, N=1, V=1,
                 overrides={"endFreq": 120}, compact=True)
    # X1..X3: convex room (shipped ARG RIR yaml, smaller sampling rate)
    convex_case("x1_convex_rir_mix_o4_mono", "RIR", "MIX", 4, overrides={"sampleRate": 4000})
    convex_case("x2_convex_rir_lc_o5_N2", "RIR", "LC", 5, N=2, overrides={"sampleRate": 250})
    convex_case("x3_convex_rir_org_o3_N2", "RIR", "ORG", 3, N=2, overrides={"sampleRate": 500})
    class_cases()


def class_cases():
    """Cases for the standalone DEISM class (deism_b200/core_deism.py) the other fixtures do not
    reach: convex RTF mode (the ARG_RTF yaml), and a shoebox run driven through
    update_source_receiver(source, receiver) / update_room / absorption input."""
    convex_case("x4_convex_rtf_lc_o3_mono", "RTF", "LC", 3, overrides={"endFreq": 400})
    d = ref_import.make_deism("RTF", "shoebox")
    p = d.params
    p["DEISM_method"], p["maxReflOrder"], p["endFreq"] = "MIX", 6, 160
    d.update_room(roomDimensions=np.array([6.0, 4.5, 3.0]))
    d.update_wall_materials([0.1, 0.2, 0.15, 0.1, 0.3, 0.25], None, "absorption")
    d.update_freqs()
    d.update_directivities()
    d.update_source_receiver(np.array([2.0, 1.5, 1.2]), np.array([4.5, 3.0, 1.7]))
    d.run_DEISM(if_clean_up=False)
    im = p["images"]
    save("s7_shoebox_rtf_mix_o6_absorption_moved", pack(d, {"meta/n_images": np.array(len(im["A_early"]) + len(im["A_late"]))}))


def full_cases(which):
    if "c1" in which:
        shoebox_case("c1_full", "RTF", "MIX", 25, store_images=False, store_inputs=False)
    if "c2" in which:
        shoebox_case("c2_full_mono", "RIR", "LC", 20, store_images=False, store_inputs=False,
                     overrides={"sampleRate": 48000},
                     materials=(1.0 / 3.0, None, "reverberationTime"))
        shoebox_case("c2_full_N5", "RIR", "LC", 20, N=5, V=5, store_images=False,
                     store_inputs=False, overrides={"sampleRate": 48000},
                     materials=(1.0 / 3.0, None, "reverberationTime"))
    if "c3" in which:
        # C3 sub-sampled in frequency: the full image list of the order-25 ORG N=V=10 sum on
        # every 20th frequency of 2:2:4000 Hz (the full 2000-frequency reference run costs
        # ~50 CPU-minutes).  Coefficients drawn for K=100.
        shoebox_case("c3_sub100", "RTF", "ORG", 25, N=10, V=10, store_images=False,
                     store_inputs=False,
                     overrides={"startFreq": 40, "endFreq": 4000, "freqStep": 40})
    if "c3full" in which:
        # the whole 2:2:4000 Hz grid of C3 (2 000 frequencies, ~8 core-hours of the reference's
        # Numba kernel): coefficients drawn for K=2000, i.e. what workloads.build("c3") draws.
        shoebox_case("c3_full", "RTF", "ORG", 25, N=10, V=10, store_images=False,
                     store_inputs=False,
                     overrides={"startFreq": 2, "endFreq": 4000, "freqStep": 2})
    if "c4" in which:
        convex_case("c4_full", "RIR", "MIX", 10, store_inputs=False)
    if "c5" in which:
        shoebox_case("c5_full_N5", "RIR", "MIX", 40, N=5, V=5, store_images=False,
                     store_inputs=False, overrides={"sampleRate": 48000},
                     materials=(2.0 / 3.0, None, "reverberationTime"))


def augment_full():
    """Add the small inputs to full-size fixtures minted before small_inputs() existed
    (re-derives them from the reference without re-running the minutes-long solves)."""
    specs = {
        "c3_sub100": ("RTF", {"startFreq": 40, "endFreq": 4000, "freqStep": 40}, None),
        "c5_full_N5": ("RIR", {"sampleRate": 48000}, (2.0 / 3.0, None, "reverberationTime")),
    }
    for name, (mode, overrides, materials) in specs.items():
        path = os.path.join(GOLDEN, name + ".npz")
        if not os.path.exists(path):
            continue
        d = ref_import.make_deism(mode, "shoebox")
        for key, val in overrides.items():
            d.params[key] = val
        if materials is None:
            d.update_wall_materials()
        else:
            d.update_wall_materials(*materials)
        d.update_freqs()
        old = dict(np.load(path))
        assert abs(float(old["params/reverberationTime"]) - d.params["reverberationTime"]) == 0
        old.update(small_inputs(d.params))
        np.savez_compressed(path, **old)
        print("  augmented", path, os.path.getsize(path) // 1024, "KiB")


def scripted_case(name="c1_scripted"):
    """examples/deism_singleparam_example.py:21-38 call for call (BASELINE configs[0] "as scripted":
    RIR mode, room 10 x 8 x 2.5 m, T60 = 1 s -> 24 000 frequencies, reflection order 30)."""
    t0 = time.time()
    d = ref_import.make_deism("RIR", "shoebox")
    d.update_room(roomDimensions=np.array([10.0, 8.0, 2.5]))
    d.update_wall_materials(datain=1, datatype="reverberationTime")
    d.params["sampleRate"] = 48000
    d.params["reverberationTime"] = 1
    d.update_freqs()
    d.update_directivities()
    d.params["maxReflOrder"] = 30
    t1 = time.time()
    d.update_source_receiver()
    t2 = time.time()
    im = d.params["images"]
    nimg = len(im["A_early"]) + len(im["A_late"])
    d.run_DEISM(if_clean_up=True, if_shutdown_ray=True)
    t3 = time.time()
    print(f"[{name}] I={nimg} K={len(d.params['freqs'])} setup {t1 - t0:.2f}s images {t2 - t1:.2f}s run {t3 - t2:.2f}s (cold)")
    save(name, {"RTF": np.asarray(d.params["RTF"]), "meta/n_images": np.array(nimg),
                "meta/t_images_s": np.array(t2 - t1), "meta/t_run_cold_s": np.array(t3 - t2),
                "params/impedance_col0": np.ascontiguousarray(np.asarray(d.params["impedance"])[:, 0])})


def highpass_case(name="h1_rir_highpass"):
    """DEISM.get_results(highpass_filter=True, bandpass_window=False) (cd:754-757, utilities.py:22-92)
    of the reference on the frozen RTF of the s4 fixture: zero-phase and causal."""
    z = np.load(os.path.join(GOLDEN, "s4_shoebox_rir_mix_o9_N2.npz"))
    d = ref_import.make_deism("RIR", "shoebox")
    d.params["sampleRate"] = 1500
    d.update_wall_materials(0.08, None, "reverberationTime")
    d.update_freqs()
    assert np.array_equal(d.params["freqs"], z["params/freqs"])
    out = {}
    for tag, zero_phase, cut in (("zero_phase_30", True, 30.0), ("causal_100", False, 100.0)):
        d.params["RTF"] = np.array(z["RTF"], copy=True)
        out["RIR_" + tag] = np.asarray(d.get_results(highpass_filter=True, bandpass_window=False,
                                                     cut_freq=cut, zero_phase=zero_phase))
    save(name, out)


def choras_job(kind):
    """A CHORAS job file (examples/exampleInput_Deism.json layout) with its geometry block filled."""
    names = ("wall1", "wall3", "wall2", "wall4", "floor", "ceiling")      # DEISM order x1 x2 y1 y2 z1 z2
    if kind == "shoebox":
        L = (5.0, 4.0, 3.0)
        verts = [[x, y, z] for x in (0, L[0]) for y in (0, L[1]) for z in (0, L[2])]
        centers = [[0, 2, 1.5], [5, 2, 1.5], [2.5, 0, 1.5], [2.5, 4, 1.5], [2.5, 2, 0], [2.5, 2, 3]]
        areas = [12.0, 12.0, 15.0, 15.0, 20.0, 20.0]
        volume = 60.0
        src, rec = (1.2, 1.4, 1.5), (3.6, 2.7, 1.2)
    else:
        # the tilted-ceiling room of configSingleParam_ARG_RIR.yml, walls named in its wallCenters order
        verts = [[0, 0, 0], [0, 0, 3.5], [0, 3, 2.5], [0, 3, 0], [4, 0, 0], [4, 0, 3.5], [4, 3, 2.5], [4, 3, 0]]
        centers = [[0, 1.5, 1.5], [2, 3, 1.25], [4, 1.5, 1.5], [2, 0, 1.75], [2, 1.5, 0], [2, 1.5, 3]]
        areas = [9.0, 10.0, 9.0, 14.0, 12.0, 4 * np.sqrt(10.0)]
        volume = 36.0
        src, rec = (1.1, 1.1, 1.3), (2.9, 1.9, 1.3)
    absorb = {"wall1": "0.10, 0.12, 0.15, 0.20, 0.25", "wall3": "0.30, 0.28, 0.25, 0.22, 0.20",
              "wall2": "0.05, 0.06, 0.08, 0.10, 0.12", "wall4": "0.15, 0.15, 0.15, 0.15, 0.15",
              "floor": "0.02, 0.03, 0.04, 0.05, 0.06", "ceiling": "0.60, 0.69, 0.71, 0.70, 0.63"}
    return {"absorption_coefficients": absorb, "should_cancel": False,
            "geometry": [{"vertices": verts, "room_volume": volume,
                          "room_areas": dict(zip(names, [float(a) for a in areas])),
                          "wall_centers": {n: ", ".join(f"{v:.4f}" for v in c) for n, c in zip(names, centers)}}],
            "results": [{"sourceX": src[0], "sourceY": src[1], "sourceZ": src[2], "resultType": "Deism",
                         "frequencies": [125, 250, 500, 1000, 2000],
                         "responses": [{"x": rec[0], "y": rec[1], "z": rec[2], "receiverResults": []}]}]}


def choras_cases():
    """The reference's CHORAS flow (examples/DeismInterface.py:167-243, minus the Gmsh geometry sync
    that fills the geometry block) through the reference's own DEISM class."""
    import json

    names = ("wall1", "wall3", "wall2", "wall4", "floor", "ceiling")
    for kind, overrides in (("shoebox", {"sampleRate": 2000, "maxReflOrder": 8}),
                            ("convex", {"sampleRate": 1000, "maxReflOrder": 4})):
        job = choras_job(kind)
        geo, res = job["geometry"][0], job["results"][0]
        bands = np.array(res["frequencies"])
        num = lambda v: np.array([float(x) for x in v.split(",")]) if isinstance(v, str) else np.array([float(v)])
        absorption, centers, areas = np.zeros((6, len(bands))), np.zeros((6, 3)), np.zeros((6, 1))
        for i, w in enumerate(names):
            absorption[i, :] = num(job["absorption_coefficients"][w])
            centers[i, :] = num(geo["wall_centers"][w])
            areas[i, :] = num(geo["room_areas"][w])
        vertices = np.array(geo["vertices"])
        d = ref_import.make_deism("RIR", kind)
        d.params.update(overrides)
        dims = vertices.max(axis=0) - vertices.min(axis=0) if kind == "shoebox" else vertices
        d.update_room(dims, centers, geo["room_volume"], areas)
        d.update_wall_materials(absorption, bands, "absorption")
        d.update_freqs()
        d.params["posSource"] = np.array([res["sourceX"], res["sourceY"], res["sourceZ"]])
        d.params["posReceiver"] = np.array([res["responses"][0][k] for k in "xyz"])
        d.update_source_receiver()
        d.update_directivities()
        d.run_DEISM(if_clean_up=True)
        rtf = np.array(d.params["RTF"], copy=True)
        rir = np.asarray(d.get_results())
        save(f"j_choras_{kind}", {"job_json": np.array(json.dumps(job)), "RTF": rtf, "RIR": rir,
                                  "overrides_json": np.array(json.dumps(overrides)),
                                  "impedance": np.asarray(d.params["impedance"]),
                                  "reverberationTime": np.asarray(d.params["reverberationTime"])})
        print(f"[j_choras_{kind}] K={len(rtf)} T60={d.params['reverberationTime']}")


def refit_case(name="r1_arg_refit_N3", n_img=40, N=3, n_dir=50, seed=3):
    """the reference's cal_C_nm_s_new (cd:1545-1605) + vectorize_C_nm_s_ARG
    (cd:1193-1217) on the reflection matrices of the x2 convex fixture and a synthetic sampled
    source field (the directivity .mat data sets are absent from the reference tree)."""
    ref_import.import_reference()
    from deism.core_deism import cal_C_nm_s_new, vectorize_C_nm_s_ARG

    x2 = np.load(os.path.join(GOLDEN, "x2_convex_rir_lc_o5_N2.npz"), allow_pickle=True)
    rng = np.random.default_rng(seed)
    R = np.ascontiguousarray(x2["params/reflection_matrix"][:, :, :n_img])
    k, freqs = x2["params/waveNumbers"], x2["params/freqs"]
    # quasi-uniform sampling sphere (Fibonacci lattice), rotated by a fixed ZXZ rotation
    i = np.arange(n_dir) + 0.5
    pol, az = np.arccos(1 - 2 * i / n_dir), np.pi * (1 + 5 ** 0.5) * i
    X = np.vstack((np.sin(pol) * np.cos(az), np.sin(pol) * np.sin(az), np.cos(pol)))
    from deism.core_deism import rotation_matrix_ZXZ
    X = rotation_matrix_ZXZ(0.3, 0.9, -1.1) @ X
    Psh = rng.standard_normal((len(k), n_dir)) + 1j * rng.standard_normal((len(k), n_dir))
    params = {"waveNumbers": k, "sourceOrder": N, "radiusSource": 0.4, "freqs": freqs,
              "track_updated_where": False}
    C4 = cal_C_nm_s_new(R, Psh, X, params)
    params["C_nm_s_ARG"] = C4.astype(np.complex64)                      # cd:1708
    params["images"] = {"R_sI_r_all": np.zeros((3, n_img), np.float32)}
    params = vectorize_C_nm_s_ARG(params)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, reflection_matrix=R, waveNumbers=k, freqs=freqs, coords=X, Psh=Psh,
                        radiusSource=0.4, sourceOrder=N, C_nm_s_ARG=params["C_nm_s_ARG"],
                        C_nm_s_ARG_vec=params["C_nm_s_ARG_vec"])
    print("  wrote", path, os.path.getsize(path) // 1024, "KiB")


def sampled_directivity_case(name="d1_sampled_directivities", N=3, V=2, n_dir=40, seed=9):
    """The reference's init_source_directivities / init_receiver_directivities (cd:1245-1397) on a
    synthetic sampled field (its loader is patched: the .mat data sets are absent from the tree)."""
    ref_import.import_reference()
    import deism.core_deism as cd

    rng = np.random.default_rng(seed)
    freqs = np.arange(50.0, 550.0, 50.0)
    k = 2 * np.pi * freqs / 343.0
    i = np.arange(n_dir) + 0.5
    Dir = np.stack(((np.pi * (1 + 5 ** 0.5) * i) % (2 * np.pi), np.arccos(1 - 2 * i / n_dir)), axis=1)
    Ps = rng.standard_normal((len(k), n_dir)) + 1j * rng.standard_normal((len(k), n_dir))
    Pr = rng.standard_normal((len(k), n_dir)) + 1j * rng.standard_normal((len(k), n_dir))
    data = {"source": (freqs, Ps, Dir, 0.4), "receiver": (freqs, Pr, Dir, 0.5)}
    keep = cd.load_directive_pressure
    cd.load_directive_pressure = lambda silent, which, name_: data[which]
    try:
        p = {"silentMode": 1, "sourceType": "synthetic", "receiverType": "synthetic", "freqs": freqs,
             "waveNumbers": k, "sourceOrder": N, "receiverOrder": V, "radiusSource": 0.4, "radiusReceiver": 0.5,
             "orientSource": np.array([30.0, 20.0, -10.0]), "orientReceiver": np.array([-75.0, 5.0, 40.0]),
             "ifReceiverNormalize": 1, "pointSrcStrength": 1j * k * 343.0 * 1.2 * 0.001,
             "track_updated_where": False}
        p = cd.init_receiver_directivities(cd.init_source_directivities(p))
    finally:
        cd.load_directive_pressure = keep
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, freqs=freqs, waveNumbers=k, Dir_all=Dir, Psh_source=Ps, Psh_receiver=Pr,
                        orientSource=p["orientSource"], orientReceiver=p["orientReceiver"],
                        pointSrcStrength=p["pointSrcStrength"], sourceOrder=N, receiverOrder=V,
                        C_nm_s=p["C_nm_s"], C_vu_r=p["C_vu_r"])
    print("  wrote", path, os.path.getsize(path) // 1024, "KiB")


def materials_case(name="m1_materials"):
    """the reference's material conversions (cd:807-1123) and interpolate_functions
    (cd:1126-1163) on a 4 x 3 x 2.5 m room."""
    ref_import.import_reference()
    from deism.core_deism import convert_imp_abs_t60_shoebox, interpolate_functions

    rng = np.random.default_rng(5)
    V, A, c = 30.0, np.array([7.5, 7.5, 10.0, 10.0, 12.0, 12.0]), 343.0
    out = {"V": V, "A": A, "c": c}
    Z = (rng.uniform(5, 40, (6, 3)) + 1j * rng.uniform(-8, 8, (6, 3)))
    out["imp_in"] = Z
    i, a, t = convert_imp_abs_t60_shoebox(V, A, c, Z, "impedance")
    out["imp_abs"], out["imp_t60"] = a, t
    absn = np.tile(np.array([0.6, 0.69, 0.71, 0.7, 0.63]), (6, 1)) * rng.uniform(0.5, 1.0, (6, 1))
    out["abs_in"] = absn
    i, a, t = convert_imp_abs_t60_shoebox(V, A, c, absn, "absorption")
    out["abs_imp"], out["abs_t60"] = i, t
    for key, T in (("t1", 1.0 / 3.0), ("t2", 2.0 / 3.0), ("t3", 0.184515)):
        i, a, t = convert_imp_abs_t60_shoebox(V, A, c, T, "reverberationTime")
        out[key + "_in"], out[key + "_imp"], out[key + "_abs"] = T, i, a
    dense = np.concatenate(([2.0, 100.0], np.linspace(125.0, 4000.0, 57), [4000.0, 6000.0, 24000.0]))
    out["dense"] = dense
    for key, bands, data in (
            ("i3", np.array([250.0, 1000.0, 4000.0]), Z),
            ("i5", np.array([125.0, 250.0, 500.0, 1000.0, 2000.0]), absn.astype(complex)),
            ("i2", np.array([500.0, 2000.0]), Z[:, :2]),
            ("i1", np.array([1000.0]), Z[:, :1]),
            ("i7", np.array([125.0, 250.0, 500.0, 1000.0, 2000.0, 3000.0, 4000.0]),
             np.array([[1, 1, 2, 2, 1.5, 3, 3], [5, 4, 3, 2, 1, 0, -1], [0, 0, 0, 1, 0, 0, 0],
                       [1, -1, 1, -1, 1, -1, 1]], dtype=float)
             + 1j * np.array([[0, 1, 0, 2, 0, 3, 0], [1, 1, 1, 1, 1, 1, 1], [3, 2, 2, 2, 5, 5, 9],
                              [0, 0.1, 0.4, 0.9, 1.6, 2.5, 3.6]]))):
        out[key + "_bands"], out[key + "_data"] = bands, data
        out[key + "_out"] = interpolate_functions(data, bands, dense)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print("  wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", nargs="*", default=None,
                    help="full-size configs to mint (c1 c2 c3 c4 c5); no value = all")
    ap.add_argument("--skip-small", action="store_true")
    ap.add_argument("--augment", action="store_true")
    ap.add_argument("--refit", action="store_true", help="mint only the ARG refit fixture (f2)")
    ap.add_argument("--materials", action="store_true", help="mint only the materials fixture (f4)")
    ap.add_argument("--directivities", action="store_true", help="mint only the sampled-directivity fixture")
    ap.add_argument("--class-cases", action="store_true", help="mint only the DEISM-class fixtures")
    ap.add_argument("--choras", action="store_true", help="mint only the CHORAS job fixtures")
    ap.add_argument("--highpass", action="store_true", help="mint only the high-pass RIR fixture")
    ap.add_argument("--scripted", action="store_true", help="mint only the as-scripted example fixture")
    args = ap.parse_args()
    if args.refit:
        refit_case()
        sys.exit(0)
    if args.materials:
        materials_case()
        sys.exit(0)
    if args.directivities:
        sampled_directivity_case()
        sys.exit(0)
    if args.class_cases:
        class_cases()
        sys.exit(0)
    if args.choras:
        choras_cases()
        sys.exit(0)
    if args.highpass:
        highpass_case()
        sys.exit(0)
    if args.scripted:
        scripted_case()
        sys.exit(0)
    if args.augment:
        augment_full()
        sys.exit(0)
    if not args.skip_small:
        small_cases()
    if args.full is not None:
        full_cases(args.full or ["c1", "c2", "c3", "c4", "c5"])
