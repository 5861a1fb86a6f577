"""GPU parity tests (run on the B200 box): CUDA path through the C ABI vs the reference-minted
golden fixtures and vs the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): image index lists / float32 geometry bit-exact; complex RTF
within relative L2 1e-8 of the reference's complex128 result.  The tolerance asserted here is
1e-9 (the observed error is ~1e-13; the budget left is for the 2^-29-rare float32 tie flips
discussed in DESIGN.md).
"""
import hashlib

import numpy as np
import pytest

from conftest import Golden, golden_names, rel_l2, rel_max

pytestmark = pytest.mark.gpu
RTF_TOL = 1e-9
SHOEBOX = golden_names("s")


def _images_from_params(p, method):
    from deism_b200 import images as im

    out = im.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
    return im.merge_images(out) if method in ("ORG", "LC") else out


@pytest.mark.parametrize("name", SHOEBOX)
def test_image_lists_bit_equal(name):
    g = Golden(name)
    p = g.params()
    ref = p.pop("images")
    if ref.get("storage") == "compact":
        p["shoeboxCompactImages"] = True
    got = _images_from_params(p, p["DEISM_method"])
    for key, val in ref.items():
        if key == "storage" or key.startswith("atten"):
            continue
        assert got[key].dtype == val.dtype and got[key].shape == val.shape, key
        assert np.array_equal(got[key].view(np.uint8), val.view(np.uint8)), key


@pytest.mark.parametrize("name", [n for n in SHOEBOX if "compact" not in n])
def test_materialised_attenuation_bit_equal(name):
    """atten[I,K] complex64 == the reference's (cd:2900-2928) up to float32 tie flips."""
    g = Golden(name)
    p = g.params()
    ref = p.pop("images")
    p["shoeboxMaterializeAtten"] = True
    got = _images_from_params(p, p["DEISM_method"])
    for key, val in ref.items():
        if not key.startswith("atten"):
            continue
        assert got[key].dtype == np.complex64 and got[key].shape == val.shape
        a, b = got[key].view(np.float32), val.view(np.float32)
        n_diff = int(np.count_nonzero(a != b))
        assert n_diff <= max(1, a.size // 1_000_000), f"{key}: {n_diff} of {a.size} differ"
        assert np.allclose(a, b, rtol=2e-7, atol=1e-38)


@pytest.mark.parametrize("name", [n for n in SHOEBOX if "_lc_" in n])
def test_lc_reference_layout(name):
    """Reference images dict as-is (materialised complex64 attenuation streamed to the GPU)."""
    from deism_b200 import backend

    g = Golden(name)
    got = backend.run_DEISM(g.params())
    assert got.dtype == np.complex128 and got.shape == g.rtf.shape
    assert rel_l2(got, g.rtf) < RTF_TOL
    assert rel_max(got, g.rtf) < 1e-8


@pytest.mark.parametrize("name", [n for n in SHOEBOX if "_lc_" in n])
def test_lc_fused(name):
    """Own image generator + attenuation fused in-kernel == reference run_DEISM."""
    from deism_b200 import backend

    g = Golden(name)
    p = g.params()
    p.pop("images")
    p["images"] = _images_from_params(p, "LC")
    assert p["images"]["storage"] == "fused" and "atten_all" not in p["images"]
    got = backend.run_DEISM(p)
    assert rel_l2(got, g.rtf) < RTF_TOL


@pytest.mark.parametrize("kind", ["complex", "real_plus_1e-16j", "real_same_opposite_walls", "mixed"])
@pytest.mark.parametrize("name", ["s5_shoebox_rtf_lc_o7_mono_noang", "s1_shoebox_rtf_mix_o5_mono",
                                  "s4_shoebox_rir_mix_o9_N2", "s2_shoebox_rtf_lc_o6_N3"])
def test_frequency_dependent_impedance_fused_vs_materialised(name, kind):
    """Impedance that varies with frequency (complex, different per wall): the in-kernel per-term
    attenuation (monopole, tensor-core and ORG kernels) against the same run fed with the
    attenuation array materialised by the image generator and against the CPU oracle on it."""
    from deism_b200 import backend, images as im
    from oracle import oracle

    p = Golden(name).params()
    p.pop("images")
    K = len(p["waveNumbers"])
    ramp = 1.0 + 0.3 * np.linspace(0.0, 1.0, K)
    Z = np.asarray(p["impedance"], dtype=np.complex128) * ramp[None, :]
    Z = Z * (1.0 + 0.1 * np.arange(6))[:, None] + 1j * (2.0 - 3.0 * np.linspace(0, 1, K))[None, :]
    if kind == "real_plus_1e-16j":
        # what the reference's material conversions return (cd:807-880): the kernels' real chain
        Z = Z.real + 1e-16j
    elif kind == "real_same_opposite_walls":
        Z = Z.real[[0, 0, 2, 2, 4, 4]] + 0j
    elif kind == "mixed":
        # x walls complex, y walls real, z walls real except at a few frequencies
        Z = np.concatenate([Z[:2], Z.real[2:4] + 1e-16j, Z.real[4:] + 0j])
        Z[4:, ::7] += 0.5j
    p["impedance"] = np.ascontiguousarray(Z)
    method = p["DEISM_method"]
    fused = im.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
    mat = im.pre_calc_images_src_rec_optimized_nofs_v2_numba(p, materialize=True)
    if method in ("ORG", "LC"):
        fused, mat = im.merge_images(fused), im.merge_images(mat)
    assert fused["storage"] == "fused" and mat["storage"] == "materialized"
    p["images"] = fused
    got = backend.run_DEISM(p)
    p["images"] = mat
    via_array = backend.run_DEISM(p)
    want = oracle.run_DEISM(p)
    assert rel_l2(via_array, want) < RTF_TOL
    assert rel_l2(got, want) < RTF_TOL


@pytest.mark.xfail(strict=False, reason="written after the round's GPU minutes were spent: not yet run on a GPU")
@pytest.mark.parametrize("kind", ["real_same_opposite_walls", "real_plus_1e-16j", "complex"])
@pytest.mark.parametrize("name", ["s2_shoebox_rtf_lc_o6_N3", "s4_shoebox_rir_mix_o9_N2",
                                  "s5_shoebox_rtf_lc_o7_mono_noang"])
def test_frequency_dependent_impedance_compact_storage(name, kind):
    """The reference's compact image storage (no attenuation array; beta rebuilt per batch from the
    float32 angles, pb:140-186, NOT rounded to complex64) with an impedance that varies with
    frequency: the kernels' unrounded per-term chains (two Newton steps) against the CPU oracle."""
    from deism_b200 import backend, images as im
    from oracle import oracle

    p = Golden(name).params()
    p.pop("images")
    K = len(p["waveNumbers"])
    Z = np.asarray(p["impedance"], dtype=np.complex128) * (1.0 + 0.3 * np.linspace(0.0, 1.0, K))[None, :]
    Z = Z * (1.0 + 0.1 * np.arange(6))[:, None] + 1j * (2.0 - 3.0 * np.linspace(0, 1, K))[None, :]
    if kind == "real_plus_1e-16j":
        Z = Z.real + 1e-16j
    elif kind == "real_same_opposite_walls":
        Z = Z.real[[0, 0, 2, 2, 4, 4]] + 0j
    p["impedance"] = np.ascontiguousarray(Z)
    p["shoeboxCompactImages"] = True
    images = im.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
    if p["DEISM_method"] in ("ORG", "LC"):
        images = im.merge_images(images)
    assert images["storage"] == "compact"
    p["images"] = images
    got = backend.run_DEISM(p)
    want = oracle.run_DEISM(p)
    assert rel_l2(got, want) < RTF_TOL


def test_lc_vs_oracle_random_inputs():
    """Seeded random geometry / coefficients / complex frequency-dependent impedance, ragged
    sizes (K and I not multiples of the tile sizes), against the CPU oracle."""
    from deism_b200 import backend
    from oracle import oracle

    rng = np.random.default_rng(7)
    # covers the monopole kernel (0,0), the resident-coefficient shape with one / several chunks per
    # pass (5,5), (5,6), (7,2), the streaming shape (6,6), (10,7) and a monopole receiver (3,0)
    for (I, K, N, V) in [(1, 1, 0, 0), (70, 33, 2, 1), (333, 97, 4, 3), (129, 64, 5, 5), (65, 40, 10, 7),
                         (200, 130, 5, 6), (90, 70, 6, 6), (150, 65, 7, 2), (64, 64, 3, 0), (300, 200, 0, 0)]:
        Js, Jr = (N + 1) ** 2, (V + 1) ** 2
        cs = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
        cr = (rng.standard_normal((K, V + 1, 2 * V + 1)) + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(np.complex64)
        n_all, m_all, csv = oracle.vectorize(cs, N)
        v_all, u_all, crv = oracle.vectorize(cr, V)
        R = lambda: np.stack([rng.uniform(-np.pi, np.pi, I), rng.uniform(0.05, 3.0, I),
                              rng.uniform(0.5, 300.0, I)], axis=1).astype(np.float32)
        atten = (rng.standard_normal((I, K)) + 1j * rng.standard_normal((I, K))).astype(np.complex64)
        params = {"DEISM_method": "LC", "waveNumbers": np.sort(rng.uniform(0.05, 400.0, K)),
                  "n_all": n_all, "m_all": m_all, "v_all": v_all, "u_all": u_all,
                  "C_nm_s_vec": csv, "C_vu_r_vec": crv, "silentMode": 1, "angDepFlag": 1,
                  "numParaImages": 100,
                  "images": {"A": np.zeros((I, 6), np.int32), "R_sI_r_all": R(), "R_s_rI_all": R(),
                             "R_r_sI_all": R(), "atten_all": atten}}
        want = oracle.run_DEISM(params)
        got = backend.run_DEISM(params)
        assert rel_l2(got, want) < RTF_TOL, (I, K, N, V)


def test_lc_arbitrary_harmonic_lists():
    """The LC entry point takes ANY (n_all, m_all) / (v_all, u_all) lists (pb:264-325 just loops
    over them): shuffled order, missing +-m partners, only negative m, repeated entries.  The
    kernels regroup them into real harmonics; the result must not depend on that."""
    from deism_b200 import backend
    from oracle import oracle

    rng = np.random.default_rng(11)
    I, K = 150, 70
    cases = [
        ([(0, 0)], [(3, -2)]),                                          # single terms
        ([(2, -2), (2, -1), (1, -1)], [(4, 3), (4, -3), (0, 0)]),          # only negative m / a +-pair
        ([(n, m) for n in range(4) for m in range(-n, n + 1)][::-1],      # complete, reversed
         [(n, m) for n in range(3) for m in range(-n, n + 1) if m >= 0]), # m >= 0 only
        ([(1, 1), (1, 1), (1, -1), (2, 0), (1, 1)], [(2, 1), (5, -4), (2, 1)]),   # duplicates
        (list(map(tuple, rng.permutation([(n, m) for n in range(6) for m in range(-n, n + 1)])[:23])),
         list(map(tuple, rng.permutation([(n, m) for n in range(5) for m in range(-n, n + 1)])[:17]))),
    ]
    R = lambda: np.stack([rng.uniform(-np.pi, np.pi, I), rng.uniform(0.05, 3.0, I),
                          rng.uniform(0.5, 300.0, I)], axis=1).astype(np.float32)
    for src, rec in cases:
        cplx = lambda J: (rng.standard_normal((K, J)) + 1j * rng.standard_normal((K, J))).astype(np.complex64)
        atten = (rng.standard_normal((I, K)) + 1j * rng.standard_normal((I, K))).astype(np.complex64)
        params = {"DEISM_method": "LC", "waveNumbers": np.linspace(0.3, 350.0, K),
                  "n_all": np.array([a for a, _ in src]), "m_all": np.array([b for _, b in src]),
                  "v_all": np.array([a for a, _ in rec]), "u_all": np.array([b for _, b in rec]),
                  "C_nm_s_vec": cplx(len(src)), "C_vu_r_vec": cplx(len(rec)), "silentMode": 1,
                  "angDepFlag": 1, "numParaImages": 64,
                  "images": {"A": np.zeros((I, 6), np.int32), "R_sI_r_all": R(), "R_s_rI_all": R(),
                             "R_r_sI_all": R(), "atten_all": atten}}
        want = oracle.run_DEISM(params)
        got = backend.run_DEISM(params)
        assert rel_l2(got, want) < RTF_TOL, (src, rec)


def test_zero_images_gives_zeros():
    from deism_b200 import backend

    p = {"DEISM_method": "LC", "waveNumbers": np.linspace(1, 2, 5), "silentMode": 1,
         "n_all": [0], "m_all": [0], "v_all": [0], "u_all": [0], "angDepFlag": 0,
         "C_nm_s_vec": np.ones((5, 1), np.complex64), "C_vu_r_vec": np.ones((5, 1), np.complex64),
         "images": {"A": np.zeros((0, 6), np.int32), "R_sI_r_all": np.zeros((0, 3), np.float32),
                    "R_s_rI_all": np.zeros((0, 3), np.float32), "R_r_sI_all": np.zeros((0, 3), np.float32),
                    "atten_all": np.zeros((0, 5), np.complex64)}}
    assert np.array_equal(backend.run_DEISM(p), np.zeros(5, np.complex128))
    p["DEISM_method"] = "nope"
    with pytest.raises(ValueError):
        backend.run_DEISM(p)


def _full_params(g):
    p = g.params()
    p["ifRemoveDirectPath"] = int(p.get("ifRemoveDirectPath", 0))
    return p


@pytest.mark.parametrize("name", ["c2_full_mono"])
def test_full_size_lc_monopole(name):
    """BASELINE config C2 (LC, order 20, K=8000, monopole) end to end on the GPU: image lists
    hashed against the reference's, RTF against the reference's full-size result."""
    from deism_b200 import backend
    from oracle import oracle

    g = Golden(name)
    p = _full_params(g)
    k = p["waveNumbers"]
    p["C_nm_s"] = oracle.monopole_coefficients(k)
    p["C_vu_r"] = oracle.monopole_coefficients(k)
    p["n_all"], p["m_all"], p["C_nm_s_vec"] = oracle.vectorize(p["C_nm_s"], 0)
    p["v_all"], p["u_all"], p["C_vu_r_vec"] = oracle.vectorize(p["C_vu_r"], 0)
    im = _images_from_params(p, "LC")
    hashes = g.group("imagehash")
    for key in ("A", "R_sI_r_all", "R_s_rI_all", "R_r_sI_all"):
        assert hashlib.sha256(np.ascontiguousarray(im[key]).tobytes()).hexdigest() == str(hashes[key]), key
    p["images"] = im
    got = backend.run_DEISM(p)
    assert rel_l2(got, g.rtf) < RTF_TOL


def test_fp64_peak_microbenchmarks_run():
    import ctypes as C
    from deism_b200._lib import check, lib

    for kind in (0, 1):
        tf, ms = C.c_double(0), C.c_double(0)
        check(lib().deism_fp64_peak(kind, 2000, C.byref(tf), C.byref(ms)), "fp64_peak")
        print(f"fp64 peak kind={kind}: {tf.value:.2f} TFLOP/s in {ms.value:.3f} ms")
        assert 1.0 < tf.value < 200.0


# --------------------------------------------------------------------------------------
# ORG / MIX
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in SHOEBOX if "_org_" in n or "_mix_" in n])
@pytest.mark.parametrize("algo", [1, 2])
def test_org_mix_reference_layout(name, algo):
    """Reference images dict as-is; factorised (1) and direct quintuple-sum (2) ORG kernels."""
    from deism_b200 import backend

    g = Golden(name)
    p = g.params()
    p["b200OrgAlgo"] = algo
    got = backend.run_DEISM(p)
    assert rel_l2(got, g.rtf) < RTF_TOL
    assert rel_max(got, g.rtf) < 1e-8


@pytest.mark.parametrize("name", [n for n in SHOEBOX if ("_org_" in n or "_mix_" in n) and "compact" not in n])
def test_org_mix_fused(name):
    from deism_b200 import backend, wigner

    g = Golden(name)
    p = g.params()
    p.pop("images")
    p.pop("Wigner")
    p = wigner.pre_calc_Wigner(p)           # native tables instead of the fixture's sympy ones
    p["images"] = _images_from_params(p, p["DEISM_method"])
    got = backend.run_DEISM(p)
    assert rel_l2(got, g.rtf) < RTF_TOL


def test_org_vs_oracle_random_inputs():
    """Random parities / geometry / complex attenuation, N != V, ragged sizes."""
    from deism_b200 import backend, wigner
    from oracle import oracle

    rng = np.random.default_rng(11)
    for (I, K, N, V) in [(1, 1, 0, 0), (45, 37, 2, 3), (130, 33, 4, 4), (67, 20, 6, 1), (40, 12, 8, 9),
                         (33, 65, 10, 10), (97, 31, 0, 5)]:
        cs = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
        cr = (rng.standard_normal((K, V + 1, 2 * V + 1)) + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(np.complex64)
        A = rng.integers(-3, 4, size=(I, 6)).astype(np.int32)
        A[:, 3:] = rng.integers(0, 2, size=(I, 3))
        R = np.stack([rng.uniform(-np.pi, np.pi, I), rng.uniform(0.05, 3.0, I),
                      rng.uniform(0.5, 60.0, I)], axis=1).astype(np.float32)
        atten = (rng.standard_normal((I, K)) + 1j * rng.standard_normal((I, K))).astype(np.complex64)
        params = {"DEISM_method": "ORG", "waveNumbers": np.sort(rng.uniform(0.5, 100.0, K)),
                  "sourceOrder": N, "receiverOrder": V, "C_nm_s": cs, "C_vu_r": cr,
                  "Wigner": wigner.wigner_tables(N, V), "silentMode": 1, "angDepFlag": 1,
                  "numParaImages": 50,
                  "images": {"A": A, "R_sI_r_all": R, "R_s_rI_all": R, "R_r_sI_all": R, "atten_all": atten}}
        want = oracle.run_DEISM(params)
        for algo in (1, 2):
            params["b200OrgAlgo"] = algo
            got = backend.run_DEISM(params)
            assert rel_l2(got, want) < RTF_TOL, (I, K, N, V, algo)


def _synth(K, N, V, seed=0):
    rng = np.random.default_rng(seed)
    cs = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
    cr = (rng.standard_normal((K, V + 1, 2 * V + 1)) + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(np.complex64)
    return cs, cr


def _full_case(name, method, N, V):
    """Rebuild a full-size BASELINE config from the scalars in its fixture, run it end to end
    (GPU image enumeration + fused accumulation), check hashes and the RTF."""
    from deism_b200 import backend, wigner
    from oracle import oracle

    g = Golden(name)
    p = _full_params(g)
    k = p["waveNumbers"]
    K = len(k)
    p["DEISM_method"] = method
    if N == 0:
        p["C_nm_s"] = oracle.monopole_coefficients(k)
        p["C_vu_r"] = oracle.monopole_coefficients(k)
    else:
        p["C_nm_s"], p["C_vu_r"] = _synth(K, N, V)
    p["sourceOrder"], p["receiverOrder"] = N, V
    if method in ("LC", "MIX"):
        p["n_all"], p["m_all"], p["C_nm_s_vec"] = oracle.vectorize(p["C_nm_s"], N)
        p["v_all"], p["u_all"], p["C_vu_r_vec"] = oracle.vectorize(p["C_vu_r"], V)
    if method in ("ORG", "MIX"):
        p = wigner.pre_calc_Wigner(p)
    im = _images_from_params(p, method)
    hashes = g.group("imagehash")
    for key, h in hashes.items():
        if key.startswith("atten"):
            continue
        assert hashlib.sha256(np.ascontiguousarray(im[key]).tobytes()).hexdigest() == str(h), key
    p["images"] = im
    got = backend.run_DEISM(p)
    err, worst = rel_l2(got, g.rtf), rel_max(got, g.rtf)
    print(f"{name}: vs reference run_DEISM rel L2 = {err:.3e}, worst single frequency = {worst:.3e}")
    assert err < RTF_TOL
    assert worst < 1e-8      # every frequency on its own, not only the norm
    return err


def test_full_c1_shipped_rtf_yaml_mix():
    _full_case("c1_full", "MIX", 0, 0)


def test_full_c2_lc_order20_N5():
    _full_case("c2_full_N5", "LC", 5, 5)


def test_full_c3_org_order25_N10_subsampled():
    _full_case("c3_sub100", "ORG", 10, 10)


def test_full_c3_org_order25_N10():
    """BASELINE C3 on its whole grid: 20 586 images x 2 000 frequencies (2:2:4000 Hz), N = V = 10,
    against the reference's own run_DEISM (core_deism.py:608-628 -> parallel_backends.py:543-591; one
    hour of its Numba kernel, frozen by `oracle/make_golden.py --full c3full`): rel-L2 < 1e-9 and
    every single frequency within 1e-8."""
    _full_case("c3_full", "ORG", 10, 10)


def test_full_c3_all_2000_frequencies_properties():
    """C3 at its full size (20 586 images x 2 000 frequencies, N = V = 10) through size-independent
    properties (the absolute parity of the full grid is test_full_c3_org_order25_N10): (a)
    frequencies are independent, so the 100 frequencies of the sub-sampled reference run (fixture
    c3_sub100) embedded in a full-grid run reproduce that reference; (b) additivity over a split of
    the image list; (c) linearity in the source coefficients."""
    from deism_b200 import backend, wigner

    g = Golden("c3_sub100")
    p = _full_params(g)
    f_ref = np.asarray(p["freqs"], dtype=np.float64)
    freqs = np.arange(2.0, 4002.0, 2.0)
    idx = np.searchsorted(freqs, f_ref)
    assert np.array_equal(freqs[idx], f_ref)
    K, N, V = len(freqs), 10, 10
    cs_ref, cr_ref = _synth(len(f_ref), N, V)               # what the reference run used
    cs, cr = _synth(K, N, V, seed=1)
    cs[idx], cr[idx] = cs_ref, cr_ref
    Z = np.repeat(np.asarray(p["impedance"])[:, :1], K, axis=1)
    p.update({"DEISM_method": "ORG", "freqs": freqs, "waveNumbers": 2 * np.pi * freqs / p["soundSpeed"],
              "impedance": Z, "C_nm_s": cs, "C_vu_r": cr, "sourceOrder": N, "receiverOrder": V})
    p = wigner.pre_calc_Wigner(p)
    im = _images_from_params(p, "ORG")
    p["images"] = im
    full = backend.run_DEISM(p)
    assert full.shape == (K,) and np.all(np.isfinite(full.view(np.float64)))
    assert rel_l2(full[idx], g.rtf) < RTF_TOL and rel_max(full[idx], g.rtf) < 1e-8        # (a)
    n = len(im["A"])
    cut = n // 3 + 5                                          # ragged split, not a tile multiple
    parts = []
    for sl in (slice(0, cut), slice(cut, n)):
        q = dict(p)
        q["images"] = {k: (v[sl] if isinstance(v, np.ndarray) and len(v) == n else v) for k, v in im.items()}
        parts.append(backend.run_DEISM(q))
    assert rel_l2(parts[0] + parts[1], full) < 1e-12                                     # (b)
    q = dict(p)
    q["C_nm_s"] = (-2.0 * cs).astype(np.complex64)            # a power of two: exact in complex64
    assert rel_l2(backend.run_DEISM(q), -2.0 * full) < 1e-13                             # (c)


def test_c1_from_the_yaml_scalars_through_every_stage():
    """BASELINE C1 without borrowing any derived array from the fixture: the scalars of
    examples/configSingleParam_RTF.yml -> materials (impedance 18 -> T60) -> frequency grid ->
    GPU image lattice -> monopole directivities -> Wigner tables -> MIX accumulation.  Every
    intermediate the fixture also holds must be bit-equal, the RTF must be the reference's."""
    from deism_b200 import backend, directivities as D, images as im, materials as M, wigner

    g = Golden("c1_full")
    ref = g.params()
    p = {"soundSpeed": 343.0, "roomSize": np.array([4.0, 3.0, 2.5]), "roomVolumn": 30.0,
         "roomAreas": np.array([7.5, 7.5, 10.0, 10.0, 12.0, 12.0]), "posSource": np.array([1.1, 1.1, 1.3]),
         "posReceiver": np.array([2.9, 1.9, 1.3]), "impedance": [18, 18, 18, 18, 18, 18],
         "givenMaterials": ["impedance"], "mode": "RTF", "startFreq": 2, "endFreq": 24000, "freqStep": 2,
         "angDepFlag": 1, "maxReflOrder": 25, "mixEarlyOrder": 2, "ifRemoveDirectPath": 0, "numParaImages": 1000,
         "DEISM_method": "MIX", "sourceType": "monopole", "receiverType": "monopole", "ifReceiverNormalize": 0,
         "silentMode": 1, "track_updated_where": False}
    p = M.update_freqs(M.update_wall_materials(p))
    assert p["reverberationTime"] == float(ref["reverberationTime"])
    assert np.array_equal(p["waveNumbers"], ref["waveNumbers"]) and np.array_equal(p["freqs"], ref["freqs"])
    assert np.asarray(p["impedance"]).shape == (6, 12000) and np.all(np.asarray(p["impedance"]) == 18)
    p["images"] = im.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
    for key, h in g.group("imagehash").items():
        if not key.startswith("atten"):
            assert hashlib.sha256(np.ascontiguousarray(p["images"][key]).tobytes()).hexdigest() == str(h), key
    p = D.init_receiver_directivities(D.init_source_directivities(p))
    p = D.vectorize_C_vu_r(D.vectorize_C_nm_s(p))
    p = wigner.pre_calc_Wigner(p)
    got = backend.run_DEISM(p)
    assert rel_l2(got, g.rtf) < RTF_TOL and rel_max(got, g.rtf) < 1e-8


def test_full_c5_mix_order40_N5():
    _full_case("c5_full_N5", "MIX", 5, 5)


def test_c_abi_argument_errors():
    """DEISM_ERR_INVALID surfaces as ValueError with the library's message."""
    import ctypes as C
    import torch
    from deism_b200._lib import ShoeboxAtten, check, lib

    out = torch.zeros(4, dtype=torch.complex128, device="cuda")
    k = torch.ones(4, dtype=torch.float64, device="cuda")
    at = ShoeboxAtten()
    at.mode = 7
    n = np.zeros(1, np.int64)
    x = torch.zeros(64, dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError, match="unknown attenuation mode"):
        check(lib().deism_lc_shoebox(2, 4, 1, 1, n.ctypes.data, n.ctypes.data, n.ctypes.data, n.ctypes.data,
                                     x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), C.byref(at),
                                     k.data_ptr(), out.data_ptr(), 0, None), "deism_lc_shoebox")
    bad = np.array([3], np.int64)      # |m| > n
    at.mode = 0
    at.d_atten = x.data_ptr()
    with pytest.raises(ValueError, match="bad"):
        check(lib().deism_lc_shoebox(2, 4, 1, 1, n.ctypes.data, bad.ctypes.data, n.ctypes.data, n.ctypes.data,
                                     x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), C.byref(at),
                                     k.data_ptr(), out.data_ptr(), 0, None), "deism_lc_shoebox")
    with pytest.raises(ValueError, match="out of range"):
        check(lib().deism_shoebox_count_images(np.ones(3).ctypes.data, np.ones(3).ctypes.data,
                                               np.ones(3).ctypes.data, 1.0, 9999, 2, 0, x.data_ptr(),
                                               x.data_ptr(), None), "deism_shoebox_count_images")


def test_image_lattice_properties_at_order_60():
    """Size-independent checks beyond the fixtures: the uncapped lattice has
    1 + 2N(2N^2+3N+4)/3 images, every image satisfies |q-p|+|q| sums = order, the
    distance cut only removes images, and lists are deterministic."""
    from deism_b200 import images as im

    p = {"roomSize": np.array([4.0, 3.0, 2.5]), "posSource": np.array([1.1, 1.1, 1.3]),
         "posReceiver": np.array([2.9, 1.9, 1.3]), "soundSpeed": 343.0, "reverberationTime": 100.0,
         "maxReflOrder": 60, "mixEarlyOrder": 2, "ifRemoveDirectPath": 0, "angDepFlag": 1,
         "impedance": np.full((6, 2), 18.0 + 0j)}
    a = im.merge_images(im.shoebox_images(p))
    N = 60
    assert len(a["A"]) == 1 + 2 * N * (2 * N * N + 3 * N + 4) // 3
    A = a["A"].astype(np.int64)
    order = (np.abs(A[:, 0] - A[:, 3]) + np.abs(A[:, 0]) + np.abs(A[:, 1] - A[:, 4]) + np.abs(A[:, 1])
             + np.abs(A[:, 2] - A[:, 5]) + np.abs(A[:, 2]))
    assert order.max() == N and order.min() == 0
    assert len(np.unique(A, axis=0)) == len(A)
    b = im.merge_images(im.shoebox_images(p))
    assert all(np.array_equal(a[k], b[k]) for k in ("A", "R_sI_r_all", "R_s_rI_all", "R_r_sI_all"))
    p["reverberationTime"] = 0.1
    c = im.merge_images(im.shoebox_images(p))
    assert 0 < len(c["A"]) < len(a["A"])
    assert np.all(c["R_sI_r_all"][:, 2] <= np.float32(34.3) * (1 + 1e-6))
