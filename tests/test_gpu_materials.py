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


def test_update_freqs_keeps_impedance_on_device():
    """update_wall_materials -> update_freqs(device_impedance=True): Z_S[6, K] is a CUDA tensor equal
    to the host pipeline's array; a five-band absorption table goes through the same route."""
    import torch

    from deism_b200 import materials as M

    base = {"roomVolumn": 30.0, "roomAreas": np.array([7.5, 7.5, 10.0, 10.0, 12.0, 12.0]), "soundSpeed": 343.0,
            "roomSize": np.array([4.0, 3.0, 2.5]), "mode": "RTF", "startFreq": 20, "endFreq": 4000, "freqStep": 4,
            "ifReceiverNormalize": 0}
    Z = (np.random.default_rng(3).uniform(5, 30, (6, 4)) + 1j * np.random.default_rng(4).uniform(-5, 5, (6, 4)))
    bands = np.array([125.0, 500.0, 1000.0, 2000.0])
    host = M.update_freqs(M.update_wall_materials(dict(base), Z, bands, "impedance"))
    devp = M.update_freqs(M.update_wall_materials(dict(base), Z, bands, "impedance"), device_impedance=True)
    assert isinstance(devp["impedance"], torch.Tensor) and devp["impedance"].is_cuda
    assert tuple(devp["impedance"].shape) == host["impedance"].shape == (6, len(host["freqs"]))
    assert np.max(np.abs(devp["impedance"].cpu().numpy() - host["impedance"])) <= 1e-13 * np.max(np.abs(host["impedance"]))


def test_empty_inputs_are_accepted():
    import ctypes as C

    import torch

    from deism_b200 import _lib

    lib = _lib.lib()
    x = torch.zeros(16, dtype=torch.float64, device="cuda")
    assert lib.deism_pchip_interp(0, 3, 10, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), None) == 0
    assert lib.deism_pchip_interp(6, 3, 0, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), None) == 0
    assert lib.deism_arg_refit(0, 5, 2, 20, None, None, None, None, C.c_double(0.4), None, None) == 0
    assert lib.deism_arg_refit(5, 0, 2, 20, None, None, None, None, C.c_double(0.4), None, None) == 0
    assert lib.deism_arg_refit(5, 5, 2, 20, None, None, None, None, C.c_double(0.4), None, None) != 0   # null pointers
