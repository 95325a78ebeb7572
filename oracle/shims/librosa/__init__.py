"""Minimal stand-in for librosa 0.9.x (see ../README.md) -- forwards to the oracle restatement."""
import numpy as np
from oracle.stft import stft_complex
from oracle.lowpass import librosa_resample_polyphase


def stft(y, n_fft=2048, hop_length=None, **kw):
    assert not kw, kw
    return stft_complex(y, n_fft, hop_length if hop_length is not None else n_fft // 4)


def istft(stft_matrix, length=None, **kw):
    assert not kw, kw
    from oracle.postproc import istft as _istft
    return _istft(stft_matrix, length=length)


def resample(y, orig_sr, target_sr, res_type="kaiser_best", **kw):
    assert res_type == "polyphase", "only the polyphase path (ssr_eval/eval.py:144-150) is restated"
    return librosa_resample_polyphase(y, orig_sr, target_sr)


def load(*a, **k):
    raise NotImplementedError("file loading is outside the restated path")
