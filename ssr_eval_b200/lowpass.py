"""Low-resolution simulation with the reference's interface (ssr_eval/lowpass.py).

``stft_hard`` (K4), ``subsampling`` (K3 twice) and the IIR zero-phase filters (butter / cheby1 / ellip /
bessel: scipy designs the second-order sections on the host exactly as the reference does, the
``sosfiltfilt`` recursion runs in kernel K7) all execute on the GPU; there is no CPU fallback."""
import numpy as np
from scipy.signal import butter, cheby1, cheby2, ellip, bessel

import os

from .engine import HardLowpass, HardLowpassDense, PolyphaseResampler, sosfiltfilt_batch

_hard = {}
_resamplers = {}

# Arithmetic of ``stft_hard``: "fft" (default) = K4, a float32 FFT -- the same algorithm to 2e-5 per sample, ~100x
# faster, but a LOWER rounding-noise floor above the cutoff than the reference's dense float32 DFT products, which
# moves LSD / log-sispec of an unprocessed proc_fft_* input by ~+0.25 / -0.05; "dense" = K4d,
# the reference's own arithmetic (bit-identical to torchlibrosa on an AVX-512 host for almost every sample).
# Set here, per call (``mode=``), or through the environment variable SSR_STFT_HARD_MODE.
STFT_HARD_MODE = os.environ.get("SSR_STFT_HARD_MODE", "fft")


def _hard_lowpass(n_fft=2048, hop=441, mode=None):
    mode = STFT_HARD_MODE if mode is None else mode
    if mode not in ("fft", "dense"):
        raise ValueError("stft_hard mode must be 'fft' or 'dense', got %r" % (mode,))
    key = (n_fft, hop, mode)
    if key not in _hard:
        # FDomainHelper defaults, ssr_eval/dsp.py:9-10
        _hard[key] = HardLowpass(n_fft, hop) if mode == "fft" else HardLowpassDense(n_fft, hop)
    return _hard[key]


def _resampler(up, down):
    key = (up, down)
    if key not in _resamplers:
        _resamplers[key] = PolyphaseResampler(up, down)
    return _resamplers[key]


def stft_hard_lowpass_v0(data, lowpass_ratio, mode=None):
    """ssr_eval/lowpass.py:17-28 -> float32 numpy of the input length."""
    x = np.asarray(data, dtype=np.float32)
    return _hard_lowpass(mode=mode).apply([x], [lowpass_ratio])[0]


def stft_hard_lowpass_batch(waves, lowpass_ratios, mode=None):
    """Batched form: one kernel launch (sequence) for all (utterance, ratio) pairs."""
    return _hard_lowpass(mode=mode).apply([np.asarray(w, dtype=np.float32) for w in waves], lowpass_ratios)


def align_length(x, y):
    """Zero-pad or cut ``y`` to len(x) (ssr_eval/lowpass.py:31-51)."""
    Lx, Ly = len(x), len(y)
    if Lx == Ly:
        return y
    if Lx > Ly:
        return np.pad(y, (0, Lx - Ly), mode="constant")
    return y[:Lx]


def subsampling_batch(waves, lowpass_ratio, fs_ori=44100):
    """``subsampling`` for a list of utterances: two K3 launches for the whole list."""
    fs_down = int(lowpass_ratio * fs_ori)
    xs = [np.asarray(w, dtype=np.float32) for w in waves]
    ys = _resampler(fs_down, fs_ori).resample(xs)
    ys = _resampler(fs_ori, fs_down).resample(ys)
    return [y if len(y) == len(x) else align_length(x, y) for x, y in zip(xs, ys)]


def subsampling(data, lowpass_ratio, fs_ori=44100):
    """Down- then up-sample by polyphase filtering (ssr_eval/lowpass.py:134-144; fs_ori stays
    44100 whatever the true rate is, as in the reference)."""
    return subsampling_batch([data], lowpass_ratio, fs_ori)[0]


def _design(order, band, btype, ftype):
    if ftype == "butter":
        return butter(order, band, btype=btype, output="sos")
    if ftype == "cheby1":
        return cheby1(order, 0.1, band, btype=btype, output="sos")
    if ftype == "cheby2":
        return cheby2(order, 60, band, btype=btype, output="sos")
    if ftype == "ellip":
        return ellip(order, 0.1, 60, band, btype=btype, output="sos")
    if ftype == "bessel":
        return bessel(order, band, btype=btype, output="sos")
    raise Exception(f"The {btype}pass filter {ftype} is not supported!")


def sosfiltfilt(sos, x):
    """scipy.signal.sosfiltfilt(sos, x) for one 1-D signal, on the GPU (K7); float64 result."""
    return sosfiltfilt_batch(sos, [np.asarray(x, dtype=np.float32)])[0]


def lowpass_filter_batch(waves, highcut, fs, order, ftype):
    """``lowpass_filter`` for a list of utterances: one K7 launch for the whole list."""
    sos = _design(order, highcut / (0.5 * fs), "low", ftype)
    ys = sosfiltfilt_batch(sos, [np.asarray(w, dtype=np.float32) for w in waves])
    return [align_length(x, y) if len(y) != len(x) else y for x, y in zip(waves, ys)]


def lowpass_filter(x, highcut, fs, order, ftype):
    """Zero-phase IIR low-pass (ssr_eval/lowpass.py:94-131): scipy design + GPU sosfiltfilt."""
    return lowpass_filter_batch([x], highcut, fs, order, ftype)[0]


def lowpass_batch(waves, highcut, fs, order=5, _type="butter"):
    """``lowpass`` (same dispatch, same checks) for a list of utterances that share the filter settings."""
    order = limit(order, high=10, low=2)
    for data in waves:
        _check_1d(data)
    for name in ("butter", "cheby1", "ellip", "bessel"):
        if _type in name:
            return lowpass_filter_batch(waves, int(highcut), fs, order, name)
    if _type in "subsampling":
        return subsampling_batch(waves, lowpass_ratio=highcut / int(fs / 2))
    if _type in "stft_hard":
        return stft_hard_lowpass_batch(waves, [highcut / int(fs / 2)] * len(waves))
    raise ValueError("Error: Unexpected filter type " + _type)


def bandpass_filter(x, lowcut, highcut, fs, order, ftype):
    """Zero-phase IIR band-pass (ssr_eval/lowpass.py:54-91): scipy design + GPU sosfiltfilt."""
    nyq = 0.5 * fs
    sos = _design(order, [lowcut / nyq, highcut / nyq], "band", ftype)
    y = sosfiltfilt(sos, x)
    return align_length(x, y) if len(y) != len(x) else y


def limit(integer, high, low):
    """ssr_eval/lowpass.py:147-153."""
    if integer > high:
        return high
    if integer < low:
        return low
    return int(integer)


def _check_1d(data):
    if len(list(data.shape)) != 1:
        raise ValueError("Error (chebyshev_lowpass_filter): Data " + str(data.shape)
                         + " should be type 1d time array, (samples,) , can not be (samples, 1)")


def lowpass(data, highcut, fs, order=5, _type="butter"):
    """Dispatcher with the reference's substring matching on ``_type`` (ssr_eval/lowpass.py:156-196)."""
    order = limit(order, high=10, low=2)
    _check_1d(data)
    for name in ("butter", "cheby1", "ellip", "bessel"):
        if _type in name:
            return lowpass_filter(x=data, highcut=int(highcut), fs=fs, order=order, ftype=name)
    if _type in "subsampling":
        return subsampling(data, lowpass_ratio=highcut / int(fs / 2))
    if _type in "stft_hard":
        return stft_hard_lowpass_v0(data, lowpass_ratio=highcut / int(fs / 2))
    raise ValueError("Error: Unexpected filter type " + _type)


def bandpass(data, lowcut, highcut, fs, order=5, _type="butter"):
    """ssr_eval/lowpass.py:199-256."""
    _check_1d(data)
    for name in ("butter", "cheby1", "ellip", "bessel"):
        if _type in name:
            return bandpass_filter(x=data, lowcut=int(lowcut), highcut=int(highcut), fs=fs,
                                   order=limit(order, high=10, low=2), ftype=name)
    raise ValueError("Error: Unexpected filter type " + _type)
