"""GPU parity tests for the convex (DEISM-ARG) accumulation kernels against reference-minted
fixtures (reference image lists as input) and the CPU oracle."""
import numpy as np
import pytest

from conftest import Golden, golden_names, rel_l2

pytestmark = pytest.mark.gpu
RTF_TOL = 1e-9


@pytest.mark.parametrize("name", golden_names("x"))
def test_arg_accumulation_reference_layout(name):
    from deism_b200 import backend

    g = Golden(name)
    got = backend.run_DEISM_ARG(g.params())
    assert got.dtype == np.complex128 and got.shape == g.rtf.shape
    assert rel_l2(got, g.rtf) < RTF_TOL


def test_arg_unknown_method():
    from deism_b200 import backend

    g = Golden("x1_convex_rir_mix_o4_mono")
    p = g.params()
    p["DEISM_method"] = "XYZ"
    with pytest.raises(ValueError):
        backend.run_DEISM_ARG(p)
