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


_resamplers = {}


def _resampler(sr, native, res_type):
    """One plan per (target rate, native rate, filter): designing the FIR and uploading it is not per-file work."""
    from math import gcd
    from .engine import PolyphaseResampler, kaiser_best_taps
    key = (int(sr), int(native), res_type)
    if key not in _resamplers:
        if res_type == "polyphase":
            rs = PolyphaseResampler(int(sr), native)
        elif res_type == "kaiser_best":
            rs = PolyphaseResampler(int(sr), native, bank="resampy_kaiser_best")
        elif res_type == "kaiser_best_exact":
            g = gcd(int(sr), native)
            rs = PolyphaseResampler(int(sr), native, taps=kaiser_best_taps(int(sr) // g, native // g))
        else:
            raise ValueError("res_type must be 'polyphase', 'kaiser_best' or 'kaiser_best_exact', got %r" % (res_type,))
        _resamplers[key] = rs
    return _resamplers[key]


def load_audio_batch(paths, sr=None, res_type="polyphase"):
    """``load_audio`` for a list of files: files that share a native rate are converted in ONE K3 launch."""
    raw = [read_wav(_check_wav(p)) for p in paths]
    out = [None] * len(paths)
    groups = {}
    for i, (x, native) in enumerate(raw):
        if sr is None or int(sr) == native:
            out[i] = (x, native)
        else:
            groups.setdefault(native, []).append(i)
    for native, idx in groups.items():
        ys = _resampler(sr, native, res_type).resample([raw[i][0] for i in idx])
        for i, y in zip(idx, ys):
            n = int(np.ceil(len(raw[i][0]) * float(sr) / native))  # librosa.resample: fix_length to ceil(L * ratio)
            y = y[:n] if len(y) >= n else np.pad(y, (0, n - len(y)))
            out[i] = (y.astype(np.float32), int(sr))
    return out


def _check_wav(path):
    if not str(path).lower().endswith(".wav"):
        raise ValueError("only .wav files are supported in this image (no soundfile / flac codec): %s" % path)
    return path


def load_audio(path, sr=None, res_type="polyphase"):
    """Mono float32 load with optional rate conversion (librosa.load(path, sr=sr) stand-in).  ``res_type``:
    "polyphase" -- scipy resample_poly's Kaiser-5.0 FIR (bit-exact with scipy; the stand-in for the sox binary);
    "kaiser_best" -- resampy's table-interpolating resampler, what librosa 0.9's load runs (engine.resampy_kaiser_best_bank);
    "kaiser_best_exact" -- the same band-limited kernel without resampy's table interpolation (engine.kaiser_best_taps)."""
    return load_audio_batch([path], sr=sr, res_type=res_type)[0]
