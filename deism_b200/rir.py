"""RTF -> RIR synthesis on the GPU: the step right after the hot path in RIR mode.

Mirrors ``DEISM.get_results`` (core_deism.py:675-769) and ``create_bandpass_window``
(core_deism.py:538-606): cosine band-pass window (150 Hz ... 0.7 * Nyquist), DC = 0,
``irfft`` to a whole number of 1/f_step periods, pad / truncate to ``RIRLength * sampleRate``.
The inverse FFT is cuFFT through ``torch.fft.irfft`` (a plain library transform, like the
reference's ``numpy.fft.irfft``); window, DC insertion and truncation are fused around it on the
device so the RTF never leaves HBM between ``run_DEISM`` and the impulse response.
"""
from __future__ import annotations

import numpy as np
import torch


def create_bandpass_window(freqs, f_low, f_high, transition_width_low=None,
                           transition_width_high=None):
    """core_deism.py:538-606 (host, O(K))."""
    freqs = np.asarray(freqs, dtype=np.float64)
    f_nyquist = freqs[-1] + (freqs[1] - freqs[0])
    if transition_width_low is None:
        transition_width_low = max(10.0, 0.1 * f_low)
    if transition_width_high is None:
        transition_width_high = max(100.0, 0.05 * (f_nyquist - f_high))
    window = np.ones_like(freqs, dtype=np.float64)
    low_start = max(0, f_low - transition_width_low)
    m = (freqs >= low_start) & (freqs < f_low)
    if np.any(m) and f_low - low_start > 0:
        window[m] = 0.5 * (1 - np.cos(np.pi * (freqs[m] - low_start) / (f_low - low_start)))
    window[freqs < low_start] = 0.0
    high_end = min(f_nyquist, f_high + transition_width_high)
    m = (freqs > f_high) & (freqs <= high_end)
    if np.any(m) and high_end - f_high > 0:
        window[m] = 0.5 * (1 + np.cos(np.pi * (freqs[m] - f_high) / (high_end - f_high)))
    window[freqs > high_end] = 0.0
    return window


def n_samples_for(params):
    """core_deism.py:691-729: IFFT length aligned with the frequency spacing."""
    fs, T60 = params["sampleRate"], params["reverberationTime"]
    freqs = np.asarray(params["freqs"])
    if len(freqs) > 1:
        f_step = freqs[1] - freqs[0]
        N_period = int(np.round((1 / f_step) * fs))
        N_rt60 = int(np.round(T60 * fs))
        if N_rt60 <= N_period:
            return N_period
        return int(np.ceil(N_rt60 / N_period)) * N_period
    return int(np.round(T60 * fs)) + 1


def _full_length_response(params, RTF, bandpass_window):
    """Band-pass window + irfft on the device, before truncation (cd:731-752)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    rtf = RTF if isinstance(RTF, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(RTF))
    rtf = rtf.to(dev, non_blocking=True)
    fs = params["sampleRate"]
    if bandpass_window:
        f_low, f_high = 150, int(fs / 2 * 0.7)
        w = create_bandpass_window(params["freqs"], f_low, f_high, int(f_low * 0.3), int(f_high * 0.15))
        rtf = rtf * torch.from_numpy(w).to(dev, non_blocking=True)
    full = torch.cat([torch.zeros(1, dtype=rtf.dtype, device=dev), rtf])
    return torch.fft.irfft(full, n=n_samples_for(params))


def _fit_length(p, n_out):
    """cd:759-764: zero-pad or cut to RIRLength * sampleRate samples."""
    if p.shape[0] < n_out:
        pad = torch.zeros(n_out - p.shape[0], dtype=p.dtype, device=p.device) if isinstance(p, torch.Tensor) \
            else np.zeros(n_out - p.shape[0])
        return torch.cat([p, pad]) if isinstance(p, torch.Tensor) else np.concatenate([p, pad])
    return p[:n_out]


def get_results_device(params, RTF, mode="RIR", highpass_filter=False, bandpass_window=True):
    """Device-side ``get_results``: RTF may be a CUDA tensor (e.g. from run_DEISM_device) or a
    NumPy array; returns a float64 CUDA tensor of length RIRLength*sampleRate (RIR mode) or the
    RTF itself (RTF mode).  Unlike the reference it does not overwrite params['RTF']."""
    if highpass_filter and bandpass_window:
        raise ValueError("Only one of highpass_filter or bandpass_window can be True")
    if mode == "RTF":
        rtf = RTF if isinstance(RTF, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(RTF))
        return rtf.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    if highpass_filter:
        raise NotImplementedError("the recursive high-pass filter (utilities.py:9-92) runs on the host: "
                                  "use rir.get_results(..., highpass_filter=True)")
    p = _full_length_response(params, RTF, bandpass_window)
    return _fit_length(p, int(params["RIRLength"] * params["sampleRate"]))


def highpass_RIR(h, fs, fcut=100.0, *, axis=0, zero_phase=False, dtype=None):
    """utilities.py:22-92: the audiolabs rir-generator high-pass (one biquad), causal (lfilter) or
    zero-phase (sosfiltfilt).  A serial recursion over the samples of one response: host code."""
    from scipy.signal import lfilter, sosfiltfilt, tf2sos

    if fs <= 0:
        raise ValueError("fs must be positive.")
    if not (0.0 < fcut < fs / 2.0):
        raise ValueError(f"fcut must be in (0, fs/2). Got fcut={fcut}, fs={fs}.")
    target = dtype or np.result_type(h, np.float32)
    x = np.asarray(h, dtype=target)
    if x.ndim == 0:
        raise ValueError("h must be at least 1-dimensional.")
    if axis < 0:
        axis += x.ndim
    if axis < 0 or axis >= x.ndim:
        raise ValueError(f"axis={axis} is out of bounds for input with ndim={x.ndim}.")
    W = 2.0 * np.pi * float(fcut) / float(fs)
    R1 = np.exp(-W)
    b = np.array([1.0, -(1.0 + R1), R1], dtype=np.float64)
    a = np.array([1.0, -2.0 * R1 * np.cos(W), R1 * R1], dtype=np.float64)
    y = sosfiltfilt(tf2sos(b, a), x, axis=axis) if zero_phase else lfilter(b, a, x, axis=axis)
    return np.asarray(y, dtype=target)


def get_results(params, RTF=None, mode="RIR", highpass_filter=False, bandpass_window=True, cut_freq=30.0,
                zero_phase=True):
    """NumPy in / NumPy out drop-in for DEISM.get_results() (cd:675-769)."""
    if highpass_filter and bandpass_window:
        raise ValueError("Only one of highpass_filter or bandpass_window can be True")
    rtf = params["RTF"] if RTF is None else RTF
    if mode == "RTF":
        return rtf.cpu().numpy() if isinstance(rtf, torch.Tensor) else rtf
    p = _full_length_response(params, rtf, bandpass_window).cpu().numpy()
    if highpass_filter:
        p = highpass_RIR(p, params["sampleRate"], cut_freq, zero_phase=zero_phase)
    return _fit_length(p, int(params["RIRLength"] * params["sampleRate"]))
