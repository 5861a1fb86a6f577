"""GPU: RTF -> RIR synthesis against DEISM.get_results() outputs frozen from the
reference (band-pass window + numpy irfft)."""
import numpy as np
import pytest

from conftest import Golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["s4_shoebox_rir_mix_o9_N2", "x1_convex_rir_mix_o4_mono", "c2_full_mono"])
def test_rir_matches_reference_get_results(name):
    from deism_b200 import rir

    g = Golden(name)
    p = g.params()
    got = rir.get_results(p, g.rtf, mode="RIR")
    want = g.z["RIR"]
    assert got.shape == want.shape and got.dtype == np.float64
    assert rel_l2(got, want) < 1e-12


def test_rtf_mode_passthrough_and_errors():
    from deism_b200 import rir

    g = Golden("s4_shoebox_rir_mix_o9_N2")
    assert np.array_equal(rir.get_results(g.params(), g.rtf, mode="RTF"), g.rtf)
    with pytest.raises(ValueError):
        rir.get_results(g.params(), g.rtf, highpass_filter=True, bandpass_window=True)


def test_highpass_variant_matches_reference():
    from deism_b200 import rir

    g, h = Golden("s4_shoebox_rir_mix_o9_N2"), Golden("h1_rir_highpass")
    for tag, zero_phase, cut in (("zero_phase_30", True, 30.0), ("causal_100", False, 100.0)):
        got = rir.get_results(g.params(), g.rtf, mode="RIR", highpass_filter=True, bandpass_window=False,
                              cut_freq=cut, zero_phase=zero_phase)
        assert rel_l2(got, h.z["RIR_" + tag]) < 1e-12
