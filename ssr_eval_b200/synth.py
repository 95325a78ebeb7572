"""Deterministic synthetic speech-like utterances.

There is no network for VCTK, so tests and ``bench.py`` use seeded synthetic audio: ~60 harmonics
of a wandering f0 (120 +- 40 Hz) with 1/k amplitudes and random phases, a 2-3 Hz amplitude
envelope, plus pink-ish (1/sqrt(f)) Gaussian noise at 0.3 relative -- full band up to Nyquist so
the high-frequency bins are non-trivial.  float32 in [-0.5, 0.5].
"""
import numpy as np


def speech_like(n_samples, sr=48000, seed=0, n_harm=60):
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sr
    # wandering f0: 120 Hz +- 40 Hz, slow random walk through a few sinusoids
    f0 = 120.0 + 40.0 * np.sin(2 * np.pi * (0.7 + 0.6 * rng.random()) * t + 2 * np.pi * rng.random()) \
        * np.sin(2 * np.pi * 0.23 * t + 2 * np.pi * rng.random())
    phase0 = 2 * np.pi * np.cumsum(f0) / sr
    sig = np.zeros(n_samples, dtype=np.float64)
    ph = rng.random(n_harm) * 2 * np.pi
    for k in range(1, n_harm + 1):
        # harmonics above Nyquist are dropped (the instantaneous f0 never exceeds 160 Hz)
        if k * 160.0 >= sr / 2:
            break
        sig += (1.0 / k) * np.sin(k * phase0 + ph[k - 1])
    env = 0.55 + 0.45 * np.sin(2 * np.pi * (2.0 + rng.random()) * t + 2 * np.pi * rng.random())
    sig *= env
    # pink-ish noise: white Gaussian shaped by 1/sqrt(f) in the frequency domain
    w = rng.standard_normal(n_samples)
    W = np.fft.rfft(w)
    f = np.fft.rfftfreq(n_samples, 1.0 / sr)
    shape = 1.0 / np.sqrt(np.maximum(f, 20.0))
    noise = np.fft.irfft(W * shape, n=n_samples)
    noise *= 0.3 * (np.std(sig) + 1e-9) / (np.std(noise) + 1e-12)
    out = sig + noise
    out *= 0.5 / (np.max(np.abs(out)) + 1e-12)
    return out.astype(np.float32)


def utterance_seed(speaker_idx, utt_idx):
    """seed = 1000 * speaker_idx + utt_idx."""
    return 1000 * int(speaker_idx) + int(utt_idx)


def ragged_lengths(n, sr=48000, lo_s=2.0, hi_s=8.0, seed=0):
    """VCTK-shaped ragged set: uniform 2-8 s."""
    rng = np.random.default_rng(seed)
    return (rng.uniform(lo_s, hi_s, size=n) * sr).astype(np.int64)
