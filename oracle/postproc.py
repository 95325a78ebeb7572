"""BasicTestee.postprocessing restated on CPU -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ssr_eval/eval.py:21-41 (in-repo: _find_cutoff, _get_cutoff_index, postprocessing) and
librosa 0.9.x ``istft`` (third-party, absent here, PARITY UNPINNED): irfft (float64) of the complex64
matrix x float64 periodic Hann, overlap-added into a float32 buffer, divided by the float32
window-sum-square where it exceeds tiny, centre trim, fix_length to ``length``.
"""
import numpy as np

from .stft import hann_periodic, stft_complex


def istft(stft_matrix, length=None):
    """librosa.istft(stft_matrix, length=length) with defaults (hop = n_fft // 4, hann, center)."""
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    hop = n_fft // 4
    win = hann_periodic(n_fft)[:, None]
    n_frames = stft_matrix.shape[1]
    if length:
        n_frames = min(n_frames, int(np.ceil((length + n_fft) / hop)))
    dtype = np.float32 if stft_matrix.dtype == np.complex64 else np.float64
    y = np.zeros(n_fft + hop * (n_frames - 1), dtype=dtype)
    ytmp = win * np.fft.irfft(stft_matrix[:, :n_frames], axis=0)
    for f in range(n_frames):
        y[f * hop:f * hop + n_fft] += ytmp[:, f]
    # librosa.filters.window_sumsquare (float32 accumulation of hann^2)
    wsq = (hann_periodic(n_fft) ** 2)
    wss = np.zeros(n_fft + hop * (n_frames - 1), dtype=dtype)
    for f in range(n_frames):
        wss[f * hop:f * hop + n_fft] += wsq
    nz = wss > np.finfo(dtype).tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2:]
    if length is None:
        return y[:-(n_fft // 2)]
    if len(y) >= length:
        return y[:length]
    return np.pad(y, (0, length - len(y)))


def find_cutoff(x, threshold=0.95):
    """ssr_eval/eval.py:21-26."""
    level = x[-1] * threshold
    for i in range(1, x.shape[0]):
        if x[-i] < level:
            return x.shape[0] - i
    return 0


def get_cutoff_index(x):
    """ssr_eval/eval.py:28-31 (librosa.stft defaults: n_fft 2048, hop 512)."""
    mag = np.abs(stft_complex(x, 2048, 512))
    energy = np.cumsum(np.sum(mag, axis=-1))
    return find_cutoff(energy, 0.97)


def postprocessing(x, out):
    """ssr_eval/eval.py:33-41: replace the bins below the input's cutoff by the input's own."""
    length = out.shape[0]
    cutoff = get_cutoff_index(x)
    s_gt = stft_complex(x, 2048, 512)
    s_out = stft_complex(out, 2048, 512)
    s_out[:cutoff, ...] = s_gt[:cutoff, ...]
    return istft(s_out, length=length)
