"""librosa.stft (0.9.x) restated -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call site: ssr_eval/metrics.py:26-30 (``np.abs(librosa.stft(wav, hop_length, n_fft))``),
parameters from ssr_eval/metrics.py:16-19.  librosa is a third-party dependency that is NOT
vendored in /root/reference and NOT installed here (setup.py:37-45, unpinned; era 0.9.x), so its
published algorithm is restated: PARITY UNPINNED for the framing rules, cross-checked against
``torch.stft`` (float64) in tests/test_oracle.py.
"""
import numpy as np
import scipy.signal


def hann_periodic(n_fft):
    """``scipy.signal.get_window("hann", n_fft, fftbins=True)`` -- the float64 window
    librosa.stft builds (librosa/core/spectrum.py ``get_window(window, win_length, fftbins=True)``)."""
    return scipy.signal.get_window("hann", int(n_fft), fftbins=True)


def n_frames(length, n_fft, hop):
    """Frame count of a centred STFT: ``1 + (L + 2*(n_fft//2) - n_fft) // hop``."""
    return 1 + (int(length) + 2 * (n_fft // 2) - n_fft) // hop


def stft_complex(y, n_fft, hop):
    """librosa.stft(y, n_fft=n_fft, hop_length=hop) with 0.9.x defaults: win_length=n_fft,
    window='hann' (periodic), center=True, pad_mode='reflect', dtype=complex64 for f32 input.

    The frame matrix is (float64 window) * (input-dtype frames) -> float64, transformed by
    numpy's pocketfft ``rfft`` in float64 and STORED as complex64 (f32 input) / complex128.
    Returns (1 + n_fft//2, n_frames).
    """
    y = np.asarray(y)
    assert y.ndim == 1
    pad = n_fft // 2
    ypad = np.pad(y, pad, mode="reflect")
    T = 1 + (ypad.shape[0] - n_fft) // hop
    frames = np.lib.stride_tricks.as_strided(
        ypad, shape=(n_fft, T), strides=(ypad.strides[0], ypad.strides[0] * hop), writeable=False)
    win = hann_periodic(n_fft).reshape(-1, 1)
    out_dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    out = np.empty((1 + n_fft // 2, T), dtype=out_dtype, order="F")
    # librosa processes column blocks to bound memory; the arithmetic per column is identical.
    blk = max(1, (2 ** 8 * 2 ** 10) // (out.shape[0] * out.itemsize))
    for s in range(0, T, blk):
        e = min(s + blk, T)
        out[:, s:e] = np.fft.rfft(win * frames[:, s:e], axis=0)
    return out


def stft_mag(wav, n_fft, hop):
    """ssr_eval/metrics.py:26-30 ``wav_to_spectrogram`` minus the torch wrapping:
    ``np.abs(stft)`` transposed to (T, F).  float32 for float32 input."""
    f = np.abs(stft_complex(wav, n_fft, hop))
    return np.ascontiguousarray(np.transpose(f, (1, 0)))
