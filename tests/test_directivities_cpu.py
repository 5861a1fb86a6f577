"""CPU: host-side precompute mirrors (vectorisers, monopole coefficients) reproduce the arrays the
reference stored in the fixtures, bit for bit."""
import numpy as np
import pytest

from conftest import Golden, golden_names


@pytest.mark.parametrize("name", [n for n in golden_names("s") if "_lc_" in n or "_mix_" in n])
def test_vectorisers_shoebox(name):
    from deism_b200 import directivities as D

    p = Golden(name).params()
    ref = {k: p[k] for k in ("n_all", "m_all", "v_all", "u_all", "C_nm_s_vec", "C_vu_r_vec")}
    q = {"C_nm_s": p["C_nm_s"], "C_vu_r": p["C_vu_r"], "sourceOrder": p["sourceOrder"],
         "receiverOrder": p["receiverOrder"], "track_updated_where": True, "updated_where": {}}
    q = D.vectorize_C_vu_r(D.vectorize_C_nm_s(q))
    for key, val in ref.items():
        assert np.array_equal(q[key], val) and q[key].dtype == val.dtype, key
    assert q["updated_where"]["C_nm_s_vec"] == ["vectorize_C_nm_s"]


def test_vectoriser_arg_and_monopoles():
    from deism_b200 import directivities as D

    p = Golden("x2_convex_rir_lc_o5_N2").params()
    q = {"C_nm_s_ARG": p["C_nm_s_ARG"], "sourceOrder": p["sourceOrder"]}
    q = D.vectorize_C_nm_s_ARG(q)
    assert np.array_equal(q["C_nm_s_ARG_vec"], p["C_nm_s_ARG_vec"])
    m = Golden("s1_shoebox_rtf_mix_o5_mono").params()
    q = {"sourceType": "monopole", "receiverType": "monopole", "waveNumbers": m["waveNumbers"]}
    q = D.init_receiver_directivities(D.init_source_directivities(q))
    assert np.array_equal(q["C_nm_s"], m["C_nm_s"]) and np.array_equal(q["C_vu_r"], m["C_vu_r"])
    assert q["sourceOrder"] == 0 and q["receiverOrder"] == 0 and q["ifReceiverNormalize"] == 0
    x = Golden("x1_convex_rir_mix_o4_mono").params()
    q = {"sourceType": "monopole", "waveNumbers": x["waveNumbers"], "reflection_matrix": x["reflection_matrix"]}
    q = D.init_source_directivities_ARG(q)
    assert np.array_equal(q["C_nm_s_ARG"], x["C_nm_s_ARG"])
    with pytest.raises(FileNotFoundError):      # what the reference's loader raises without its .mat file
        D.init_source_directivities({"sourceType": "speaker_cuboid_cyldriver_1", "waveNumbers": [1.0]})


def _refit_fixture():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "r1_arg_refit_N3.npz"))


def test_refit_oracle_matches_reference_fixture():
    """the NumPy restatement of cal_C_nm_s_new + vectorize_C_nm_s_ARG reproduces the
    reference's complex64 output (identical up to rounding flips of values that agree to ~1e-13)."""
    from oracle import oracle

    f = _refit_fixture()
    got = oracle.cal_C_nm_s_ARG_vec(f["reflection_matrix"], f["Psh"], f["coords"], f["waveNumbers"],
                                    int(f["sourceOrder"]), float(f["radiusSource"]))
    want = f["C_nm_s_ARG_vec"]
    assert got.shape == want.shape and got.dtype == want.dtype == np.complex64
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-8
    assert np.mean(got != want) < 1e-3
    # the 4-D layout the reference keeps next to it: signed m wraps NumPy-style
    N = int(f["sourceOrder"])
    for n in range(N + 1):
        for m in range(-n, n + 1):
            assert np.array_equal(f["C_nm_s_ARG"][:, n, m, :], want[:, n * n + n + m, :])


def test_rotation_matrix_zxz():
    from deism_b200 import directivities as D

    R = D.rotation_matrix_ZXZ(0.3, 0.9, -1.1)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-15) and np.isclose(np.linalg.det(R), 1.0)
    assert np.allclose(D.rotation_matrix_ZXZ(0.0, 0.0, 0.0), np.eye(3))
    # the fixture's sampling points were rotated with the reference's own implementation
    f = _refit_fixture()
    n_dir = f["coords"].shape[1]
    i = np.arange(n_dir) + 0.5
    pol, az = np.arccos(1 - 2 * i / n_dir), np.pi * (1 + 5 ** 0.5) * i
    X = np.vstack((np.sin(pol) * np.cos(az), np.sin(pol) * np.sin(az), np.cos(pol)))
    assert np.allclose(R @ X, f["coords"], atol=1e-15)


def test_sampled_directivities_match_reference_fixture():
    """Non-monopole init_source_directivities / init_receiver_directivities (rotation of the sampling
    directions, least-squares fit, division by h_n, receiver normalisation) against the arrays the
    reference produced for the same injected field (fixture d1_sampled_directivities)."""
    import os

    from deism_b200 import directivities as D

    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "d1_sampled_directivities.npz"))
    p = {"sourceType": "synthetic", "receiverType": "synthetic", "freqs": f["freqs"], "waveNumbers": f["waveNumbers"],
         "sourceOrder": int(f["sourceOrder"]), "receiverOrder": int(f["receiverOrder"]), "radiusSource": 0.4,
         "radiusReceiver": 0.5, "orientSource": f["orientSource"], "orientReceiver": f["orientReceiver"],
         "ifReceiverNormalize": 1, "pointSrcStrength": f["pointSrcStrength"],
         "sampledSource": {"freqs": f["freqs"], "Psh": f["Psh_source"], "Dir_all": f["Dir_all"], "r0": 0.4},
         "sampledReceiver": {"freqs": f["freqs"], "Psh": f["Psh_receiver"], "Dir_all": f["Dir_all"], "r0": 0.5}}
    p = D.init_receiver_directivities(D.init_source_directivities(p))
    for key in ("C_nm_s", "C_vu_r"):
        got, want = p[key], f[key]
        assert got.dtype == want.dtype == np.complex64 and got.shape == want.shape
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-7 and np.mean(got != want) < 5e-3, key
    bad = dict(p, freqs=f["freqs"] + 1.0)
    with pytest.raises(ValueError):
        D.init_source_directivities(bad)


def test_sampled_directivities_from_mat_profiles(tmp_path, monkeypatch):
    """The reference's own route: sourceType / receiverType name a .mat profile under
    data/sampled_directivity/<source|receiver>/ (data_loader.py:1099-1145).  The data sets are absent
    from the reference tree, so the profile is written here from the fixture's sampled field; the
    coefficients must equal the ones computed from the injected field (= the reference's, fixture d1)."""
    import os

    import scipy.io as sio

    from deism_b200 import data_loader, directivities as D

    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "d1_sampled_directivities.npz"))
    for kind, key, r0 in (("source", "Psh_source", 0.4), ("receiver", "Psh_receiver", 0.5)):
        d = tmp_path / "data" / "sampled_directivity" / kind
        d.mkdir(parents=True)
        sio.savemat(str(d / "probe.mat"), {"freqs_mesh": f["freqs"][None, :], "Psh": f[key], "Dir_all": f["Dir_all"],
                                           "r0": np.array([[r0]])})
    monkeypatch.chdir(tmp_path)
    freqs, Psh, Dir_all, r0 = data_loader.load_directive_pressure(1, "source", "probe")
    assert np.array_equal(freqs, f["freqs"]) and np.array_equal(Psh, f["Psh_source"]) and float(r0) == 0.4
    with pytest.raises(FileNotFoundError):
        data_loader.load_directive_pressure(1, "source", "no_such_profile")
    base = {"freqs": f["freqs"], "waveNumbers": f["waveNumbers"], "silentMode": 1,
            "sourceOrder": int(f["sourceOrder"]), "receiverOrder": int(f["receiverOrder"]), "radiusSource": 0.4,
            "radiusReceiver": 0.5, "orientSource": f["orientSource"], "orientReceiver": f["orientReceiver"],
            "ifReceiverNormalize": 1, "pointSrcStrength": f["pointSrcStrength"]}
    p_mat = dict(base, sourceType="probe", receiverType="probe")
    p_inj = dict(base, sourceType="probe", receiverType="probe",
                 sampledSource={"freqs": f["freqs"], "Psh": f["Psh_source"], "Dir_all": f["Dir_all"], "r0": 0.4},
                 sampledReceiver={"freqs": f["freqs"], "Psh": f["Psh_receiver"], "Dir_all": f["Dir_all"], "r0": 0.5})
    for fn, key in ((D.init_source_directivities, "C_nm_s"), (D.init_receiver_directivities, "C_vu_r")):
        a, b = fn(dict(p_mat))[key], fn(dict(p_inj))[key]
        assert a.dtype == np.complex64 and np.array_equal(a, b)
