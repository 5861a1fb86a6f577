"""CPU: the wall-material pipeline reproduces what the reference's own functions
returned for the fixture tests/golden/m1_materials.npz (minted by oracle/make_golden.py
--materials), and the CHORAS JSON wire format round-trips."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(HERE, "golden", "m1_materials.npz"))


def test_forward_conversions_bit_equal(fx):
    from deism_b200 import materials as M

    imp, a, t60 = M.convert_imp_abs_t60_shoebox(float(fx["V"]), fx["A"], float(fx["c"]), fx["imp_in"], "impedance")
    assert imp is not None and np.array_equal(a, fx["imp_abs"]) and t60 == float(fx["imp_t60"])


def test_inverse_conversions_match_reference_fit(fx):
    """Same solver, objective, start values and bounds as the reference: same impedance."""
    from deism_b200 import materials as M

    V, A, c = float(fx["V"]), fx["A"], float(fx["c"])
    for key in ("t1", "t2", "t3"):
        imp, a, t60 = M.convert_imp_abs_t60_shoebox(V, A, c, float(fx[key + "_in"]), "reverberationTime")
        assert imp.shape == (6, 1) and np.array_equal(imp, fx[key + "_imp"]), key
        assert np.array_equal(a, fx[key + "_abs"]) and t60 == float(fx[key + "_in"])
    imp, a, t60 = M.convert_imp_abs_t60_shoebox(V, A, c, fx["abs_in"], "absorption")
    assert np.array_equal(imp, fx["abs_imp"]) and t60 == float(fx["abs_t60"])
    # the reference's bounded fit ends on its integer grid search here (|error| is not smooth):
    # Paris' formula gives the absorption back to that grid's resolution
    assert np.allclose(M.convert_imp_to_abs(imp), fx["abs_in"], rtol=0.06)
    with pytest.raises(ValueError):
        M.convert_imp_abs_t60_shoebox(V, A, c, 1.0, "porosity")


@pytest.mark.parametrize("key", ["i3", "i5", "i2", "i1", "i7"])
def test_interpolation_matches_reference(fx, key):
    from deism_b200 import materials as M

    got = M.interpolate_functions(fx[key + "_data"], fx[key + "_bands"], fx["dense"])
    assert np.array_equal(got, fx[key + "_out"])
    with pytest.raises(ValueError):
        M.interpolate_functions(fx[key + "_data"], np.append(fx[key + "_bands"], 1e5), fx["dense"])


def test_params_drivers_shipped_c2_c5_values():
    """T60 = 1/3 s -> Z = 32.5409+1e-16j (C2), 2/3 s -> 65.0972+1e-16j (C5) in the
    shipped 4 x 3 x 2.5 m shoebox; frequency grids K = 8000 / 16000 at 48 kHz."""
    from deism_b200 import materials as M

    for T, z_want, K in ((1.0 / 3.0, 32.5409, 8000), (2.0 / 3.0, 65.0972, 16000)):
        p = {"roomVolumn": 30.0, "roomAreas": np.array([7.5, 7.5, 10.0, 10.0, 12.0, 12.0]),
             "soundSpeed": 343.0, "roomSize": np.array([4.0, 3.0, 2.5]), "mode": "RIR",
             "sampleRate": 48000, "RIRLength": 1.0, "ifReceiverNormalize": 0}
        p = M.update_freqs(M.update_wall_materials(p, T, None, "reverberationTime"))
        assert p["givenMaterials"] == ["reverberationTime"] and p["freqs_bands"].tolist() == [1000]
        assert p["impedance"].shape == (6, K) and abs(p["impedance"][0, 0].real - z_want) < 5e-5
        # ... and to the last bit what the reference's own run stored (fixtures c2_full_N5 / c5_full_N5)
        fx = np.load(os.path.join(HERE, "golden", "c2_full_N5.npz" if K == 8000 else "c5_full_N5.npz"))
        assert np.array_equal(p["impedance"][:, 0], fx["params/impedance_col0"])
        assert np.array_equal(p["waveNumbers"], fx["params/waveNumbers"])
        assert p["reverberationTime"] == float(fx["params/reverberationTime"])
        assert p["impedance"][3, 17].imag == 1e-16 and np.all(p["impedance"] == p["impedance"][0, 0])
        assert (p["n1"], p["n2"], p["n3"]) == tuple(int(np.floor(343.0 * T / L / 2)) for L in (4.0, 3.0, 2.5))
        assert np.allclose(p["waveNumbers"], 2 * np.pi * p["freqs"] / 343.0)
    with pytest.raises(ValueError):
        M.update_wall_materials({"roomVolumn": 1, "roomAreas": np.ones(6), "soundSpeed": 343.0}, 0.5, None,
                                "reverberationTime", roomtype="convex")
    assert M.load_format_materials_checks(18, "impedance").shape == (6, 1)
    assert M.load_format_materials_checks([1, 2, 3, 4, 5, 6], "absorption").shape == (6, 1)
    with pytest.raises(ValueError):
        M.load_format_materials_checks(np.ones(6), "impedance")


def test_choras_json_wire_format(tmp_path):
    from deism_b200 import materials as M

    job = {"absorption_coefficients": {w: "0.6, 0.69, 0.71, 0.7, 0.63" for w in
                                       ("floor", "wall1", "ceiling", "wall2", "wall3", "wall4")},
           "geometry": [{"vertices": [[0, 0, 0], [4, 0, 0], [4, 3, 0], [0, 3, 0], [0, 0, 2.5], [4, 0, 2.5],
                                      [4, 3, 2.5], [0, 3, 2.5]],
                         "room_areas": {"floor": 12.0, "wall1": 7.5, "ceiling": 12.0, "wall2": 10.0,
                                        "wall3": 7.5, "wall4": 10.0},
                         "wall_centers": {"floor": "2.0, 1.5, 0.0", "wall1": "0.0, 1.5, 1.25",
                                          "ceiling": "2.0, 1.5, 2.5", "wall2": "2.0, 0.0, 1.25",
                                          "wall3": "4.0, 1.5, 1.25", "wall4": "2.0, 3.0, 1.25"},
                         "room_volume": 30.0}],
           "results": [{"sourceX": 1.1, "sourceY": 1.1, "sourceZ": 1.3, "frequencies": [125, 250, 500, 1000, 2000],
                        "responses": [{"x": 2.9, "y": 1.9, "z": 1.3, "receiverResults": []}]}]}
    path = tmp_path / "job.json"
    path.write_text(json.dumps(job))
    got = M.read_choras_json(str(path))
    assert got["absorption"].shape == (6, 5) and got["absorption"][2, 1] == 0.69
    assert got["roomAreas"][:, 0].tolist() == [7.5, 7.5, 10.0, 10.0, 12.0, 12.0]      # x1 x2 y1 y2 z1 z2
    assert got["wallCenters"][1].tolist() == [4.0, 1.5, 1.25] and got["roomVolumn"] == 30.0
    assert got["posSource"].tolist() == [1.1, 1.1, 1.3] and got["posReceiver"].tolist() == [2.9, 1.9, 1.3]
    assert got["freqs_bands"].tolist() == [125, 250, 500, 1000, 2000] and got["should_cancel"] is False
    M.write_choras_result(str(path), np.array([0.0, 0.5, -0.25]))
    assert json.loads(path.read_text())["results"][0]["responses"][0]["receiverResults"] == [0.0, 0.5, -0.25]
    del job["absorption_coefficients"]["wall3"]
    path.write_text(json.dumps(job))
    with pytest.raises(KeyError):
        M.read_choras_json(str(path))


def test_vectorised_paris_grid_is_bit_equal_to_the_scalar_objective():
    """convert_abs_to_imp ends in the reference's 1000-point grid search (cd:897-905); the vectorised
    evaluation of that grid must select the same impedance: equal error values, bit for bit."""
    from deism_b200 import materials as M

    grid = np.linspace(1, 1000, 1000)
    for target in (0.01, 0.1, 0.3, 0.63, 0.71, 0.95, 0.999, 1e-12):
        scalar = np.array([M._paris_error(z, target) for z in grid])
        assert np.array_equal(scalar, M._paris_error_grid(grid, target))
    rng = np.random.default_rng(3)
    table = rng.uniform(0.02, 0.9, (6, 3))
    got = M.convert_abs_to_imp(table)
    for i in range(6):
        for j in range(3):
            want = grid[int(np.argmin([M._paris_error(z, table[i, j]) for z in grid]))] + 1e-16j
            assert got[i, j] == want
