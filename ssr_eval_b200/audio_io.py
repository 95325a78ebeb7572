"""Minimal audio file I/O for the drop-in API (the reference uses librosa.load / soundfile, which
are not in this image).  WAV only (PCM 8/16/24/32-bit and float); rate conversion uses the
polyphase K3 kernel -- a documented deviation from librosa's kaiser_best / sox resamplers
(parity unpinned, "next")."""
import numpy as np


def read_wav(path):
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1).astype(np.float32)
    return x, int(sr)


def write_wav(path, x, sr):
    from scipy.io import wavfile
    wavfile.write(path, int(sr), np.asarray(x, dtype=np.float32))


def load_audio(path, sr=None, res_type="polyphase"):
    """Mono float32 load with optional rate conversion (librosa.load(path, sr=sr) stand-in).  ``res_type``:
    "polyphase" (scipy resample_poly's Kaiser-5.0 FIR, the default here) or "kaiser_best" (the band-limited
    Kaiser-windowed sinc librosa 0.9 defaults to; same K3 kernel, different taps -- see engine.kaiser_best_taps)."""
    if not str(path).lower().endswith(".wav"):
        raise ValueError("only .wav files are supported in this image (no soundfile/flac codec): %s" % path)
    x, native = read_wav(path)
    if sr is None or int(sr) == native:
        return x, native
    from math import gcd
    from .engine import PolyphaseResampler, kaiser_best_taps
    if res_type == "polyphase":
        rs = PolyphaseResampler(int(sr), native)
    elif res_type == "kaiser_best":
        g = gcd(int(sr), native)
        rs = PolyphaseResampler(int(sr), native, taps=kaiser_best_taps(int(sr) // g, native // g))
    else:
        raise ValueError("res_type must be 'polyphase' or 'kaiser_best', got %r" % (res_type,))
    y = rs.resample([x])[0]
    n = int(np.ceil(len(x) * float(sr) / native))
    if len(y) > n:
        y = y[:n]
    elif len(y) < n:
        y = np.pad(y, (0, n - len(y)))
    return y.astype(np.float32), int(sr)
