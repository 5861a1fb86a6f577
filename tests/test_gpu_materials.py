"""GPU: deism_pchip_interp against the reference's interpolate_functions outputs (fixture
m1_materials.npz) and scipy on random band tables; the device-resident Z_S drives run_DEISM."""
import os

import numpy as np
import pytest

from conftest import Golden, rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-13          # relative, on values of O(1..40): a few ulp of the cubic's evaluation


@pytest.mark.parametrize("key", ["i3", "i5", "i2", "i1", "i7"])
def test_pchip_reference_fixture(key):
    from deism_b200 import materials as M

    fx = np.load(os.path.join(HERE, "golden", "m1_materials.npz"))
    got = M.interpolate_functions_device(fx[key + "_data"], fx[key + "_bands"], fx["dense"]).cpu().numpy()
    want = fx[key + "_out"]
    assert got.shape == want.shape and got.dtype == np.complex128
    assert np.max(np.abs(got - want)) <= TOL * np.max(np.abs(want))
    # knots are reproduced (exactly where a segment starts; the last one is the end of its cubic),
    # end values are held outside the band
    bands = fx[key + "_bands"]
    data = fx[key + "_data"].astype(np.complex128)
    at = M.interpolate_functions_device(data, bands, np.concatenate(([1.0], bands, [1e6]))).cpu().numpy()
    if len(bands) > 1:
        assert np.array_equal(at[:, 1:-2], data[:, :-1])
    assert np.max(np.abs(at[:, 1:-1] - data)) <= TOL * np.max(np.abs(data))
    assert np.array_equal(at[:, 0], at[:, 1]) and np.array_equal(at[:, -1], at[:, -2])


def test_pchip_random_tables_vs_scipy():
    from deism_b200 import materials as M

    rng = np.random.default_rng(2)
    for rows, B, K in ((6, 31, 16000), (1, 3, 7), (12, 64, 1000), (6, 4, 257)):
        bands = np.sort(rng.uniform(20.0, 20000.0, B))
        data = rng.standard_normal((rows, B)) + 1j * rng.standard_normal((rows, B))
        data[:, B // 2] = data[:, B // 2 - 1]                       # flat segments -> zero slopes
        dense = np.sort(rng.uniform(1.0, 24000.0, K))
        got = M.interpolate_functions_device(data, bands, dense).cpu().numpy()
        want = M.interpolate_functions(data, bands, dense)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), (rows, B, K)
    with pytest.raises(ValueError):
        M.interpolate_functions_device(np.ones((6, 65), complex), np.arange(65.0) + 1, np.arange(10.0))
    with pytest.raises(ValueError):
        M.interpolate_functions_device(np.ones((6, 3), complex), np.arange(4.0) + 1, np.arange(10.0))


def test_device_impedance_feeds_run_deism():
    """Three-band complex impedance (fixture s2): Z_S interpolated on the GPU and handed to
    run_DEISM as a device tensor gives the reference's RTF."""
    from deism_b200 import backend, materials as M

    g = Golden("s2_shoebox_rtf_lc_o6_N3")
    p = g.params()
    Z = np.asarray(p["impedance"])
    # rebuild the dense table from three of its own samples taken at the fixture's band centres
    freqs = np.asarray(p["freqs"])
    idx = [0, len(freqs) // 2, len(freqs) - 1]
    dense_from_three = M.interpolate_functions(Z[:, idx], freqs[idx], freqs)
    p["impedance"] = M.interpolate_functions_device(Z[:, idx], freqs[idx], freqs)
    got = backend.run_DEISM(p)
    p["impedance"] = dense_from_three
    want = backend.run_DEISM(p)
    assert rel_l2(got, want) < 1e-12
