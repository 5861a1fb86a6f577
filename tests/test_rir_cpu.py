"""CPU: the host-side pieces of the RIR synthesis (window, IFFT length) against the reference's
frozen get_results() output, with numpy's irfft standing in for cuFFT."""
import numpy as np
import pytest

from conftest import Golden, rel_l2


@pytest.mark.parametrize("name", ["s4_shoebox_rir_mix_o9_N2", "x1_convex_rir_mix_o4_mono", "c2_full_mono"])
def test_window_and_length_reproduce_reference_rir(name):
    from deism_b200 import rir

    g = Golden(name)
    p = g.params()
    fs = p["sampleRate"]
    N = rir.n_samples_for(p)
    f_low, f_high = 150, int(fs / 2 * 0.7)
    w = rir.create_bandpass_window(p["freqs"], f_low, f_high, int(f_low * 0.3), int(f_high * 0.15))
    x = np.fft.irfft(np.concatenate([[0], g.rtf * w]), n=N)
    n_out = int(p["RIRLength"] * fs)
    x = np.concatenate([x, np.zeros(max(0, n_out - len(x)))])[:n_out]
    assert rel_l2(x, g.z["RIR"]) < 1e-13


def test_highpass_filter_reproduces_reference():
    """rir.highpass_RIR (utilities.py:22-92) on numpy's irfft of the frozen s4 RTF vs the reference's
    get_results(highpass_filter=True, bandpass_window=False), zero-phase and causal."""
    from deism_b200 import rir

    g, h = Golden("s4_shoebox_rir_mix_o9_N2"), Golden("h1_rir_highpass")
    p = g.params()
    fs = p["sampleRate"]
    x = np.fft.irfft(np.concatenate([[0], g.rtf]), n=rir.n_samples_for(p))
    n_out = int(p["RIRLength"] * fs)
    for tag, zero_phase, cut in (("zero_phase_30", True, 30.0), ("causal_100", False, 100.0)):
        y = rir.highpass_RIR(x, fs, cut, zero_phase=zero_phase)
        y = np.concatenate([y, np.zeros(max(0, n_out - len(y)))])[:n_out]
        assert rel_l2(y, h.z["RIR_" + tag]) < 1e-13
    with pytest.raises(ValueError):
        rir.highpass_RIR(x, fs, fs)            # cut-off beyond Nyquist
    with pytest.raises(ValueError):
        rir.highpass_RIR(x, 0.0, 10.0)
