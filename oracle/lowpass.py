"""Degradation path of ssr_eval restated on CPU -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ssr_eval/lowpass.py:17-28,31-51,94-131,134-144,147-196 and ssr_eval/dsp.py:7-59,76-119.
``torchlibrosa`` (0.0.7-0.0.9; third-party, absent here) is restated: conv1d with
(DFT matrix x periodic Hann) float32 kernels, reflect pad n_fft//2, stride hop; ISTFT = Hermitian
extension -> 1x1 conv with (IDFT matrix / n_fft x window) kernels -> overlap-add -> divide by the
clamped overlap-added squared window -> drop n_fft//2 -> cut / zero-pad to ``length``.
PARITY UNPINNED for torchlibrosa; cross-checked against numpy rfft/irfft overlap-add in
tests/test_oracle.py.  ``scipy.signal.resample_poly`` / ``sosfiltfilt`` / IIR design are the
installed scipy, called directly.
"""
from functools import lru_cache

import numpy as np
import torch
import torch.nn.functional as F
from scipy.signal import butter, cheby1, ellip, bessel, sosfiltfilt, resample_poly  # noqa: F401

from .stft import hann_periodic


@lru_cache(maxsize=4)
def _dft_matrix(n, sign):
    # torchlibrosa DFTBase.dft_matrix / idft_matrix: np.power(exp(-+2*pi*1j/n), x*y)
    x, y = np.meshgrid(np.arange(n), np.arange(n))
    omega = np.exp(sign * 2 * np.pi * 1j / n)
    return np.power(omega, x * y)


class TorchlibrosaSTFT:
    """torchlibrosa.stft.STFT(n_fft, hop, win_length=n_fft, 'hann', center=True, 'reflect')."""

    def __init__(self, n_fft=2048, hop_length=441):
        self.n_fft, self.hop = n_fft, hop_length
        win = hann_periodic(n_fft)
        W = _dft_matrix(n_fft, -1)
        F_ = n_fft // 2 + 1
        self.w_real = torch.tensor(np.real(W[:, :F_] * win[:, None]).T, dtype=torch.float32)[:, None, :]
        self.w_imag = torch.tensor(np.imag(W[:, :F_] * win[:, None]).T, dtype=torch.float32)[:, None, :]

    def __call__(self, x):  # x: (B, L) float32
        x = x[:, None, :]
        x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode="reflect")
        real = F.conv1d(x, self.w_real, stride=self.hop)
        imag = F.conv1d(x, self.w_imag, stride=self.hop)
        return real[:, None].transpose(2, 3), imag[:, None].transpose(2, 3)  # (B,1,T,F)


class TorchlibrosaISTFT:
    """torchlibrosa.stft.ISTFT with the same parameters."""

    def __init__(self, n_fft=2048, hop_length=441):
        self.n_fft, self.hop = n_fft, hop_length
        win = hann_periodic(n_fft)
        W = _dft_matrix(n_fft, +1) / n_fft
        self.w_real = torch.tensor(np.real(W * win[None, :]).T, dtype=torch.float32)[:, :, None]
        self.w_imag = torch.tensor(np.imag(W * win[None, :]).T, dtype=torch.float32)[:, :, None]
        self.ola_window = torch.tensor(win ** 2, dtype=torch.float32)

    def __call__(self, real_stft, imag_stft, length):
        n_fft, hop = self.n_fft, self.hop
        T = real_stft.shape[2]
        re = real_stft[:, 0].transpose(1, 2)  # (B, F, T)
        im = imag_stft[:, 0].transpose(1, 2)
        full_re = torch.cat((re, torch.flip(re[:, 1:-1, :], dims=[1])), dim=1)
        full_im = torch.cat((im, -torch.flip(im[:, 1:-1, :], dims=[1])), dim=1)
        s = F.conv1d(full_re, self.w_real) - F.conv1d(full_im, self.w_imag)  # (B, n_fft, T)
        out_len = (T - 1) * hop + n_fft
        y = F.fold(s, output_size=(1, out_len), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, :]
        wsum = F.fold(self.ola_window[None, :, None].repeat(1, 1, T), output_size=(1, out_len),
                      kernel_size=(1, n_fft), stride=(1, hop)).squeeze()
        wsum = torch.clamp(wsum, 1e-11, np.inf)
        y = y / wsum[None, :]
        start = n_fft // 2
        y = y[:, start:start + length]
        if y.shape[-1] < length:
            y = torch.cat((y, torch.zeros(y.shape[0], length - y.shape[-1])), dim=-1)
        return y


class FDomainHelperOracle:
    """ssr_eval/dsp.py:7-59 (defaults window 2048, hop 441) -- only the three methods
    stft_hard_lowpass_v0 uses (dsp.py:76-81, 83-105, 107-119)."""

    def __init__(self, window_size=2048, hop_size=441):
        self.stft = TorchlibrosaSTFT(window_size, hop_size)
        self.istft = TorchlibrosaISTFT(window_size, hop_size)

    def spectrogram_phase(self, x, eps=0.0):
        real, imag = self.stft(x.float())
        mag = torch.clamp(real ** 2 + imag ** 2, eps, np.inf) ** 0.5
        return mag, real / mag, imag / mag

    def wav_to_spectrogram_phase(self, x, eps=1e-8):  # x: (B, C, L)
        outs = [self.spectrogram_phase(x[:, c, :], eps=eps) for c in range(x.shape[1])]
        return tuple(torch.cat([o[i] for o in outs], dim=1) for i in range(3))

    def spectrogram_phase_to_wav(self, sps, coss, sins, length):
        res = []
        for i in range(sps.shape[1]):
            res.append(self.istft(sps[:, i:i + 1] * coss[:, i:i + 1],
                                  sps[:, i:i + 1] * sins[:, i:i + 1], length).unsqueeze(1))
        return torch.cat(res, dim=1)


_f_helper = None


def _helper():
    global _f_helper
    if _f_helper is None:
        _f_helper = FDomainHelperOracle()
    return _f_helper


def stft_hard_lowpass_v0(data, lowpass_ratio):
    """ssr_eval/lowpass.py:17-28."""
    length = data.shape[0]
    x = torch.tensor(np.asarray(data), dtype=torch.float32)
    sps, coss, sins = _helper().wav_to_spectrogram_phase(x[None, None, ...])
    cut = int(sps.size()[-1] * lowpass_ratio)
    sps[..., cut:] = 0.0
    y = _helper().spectrogram_phase_to_wav(sps, coss, sins, length)
    return y[0, 0, :].numpy()


def align_length(x, y):
    """ssr_eval/lowpass.py:31-51."""
    Lx, Ly = len(x), len(y)
    if Lx == Ly:
        return y
    if Lx > Ly:
        return np.pad(y, (0, Lx - Ly), mode="constant")
    return y[:Lx]


def subsampling(data, lowpass_ratio, fs_ori=44100):
    """ssr_eval/lowpass.py:134-144 (fs_ori is 44100 whatever the real rate is)."""
    fs_down = int(lowpass_ratio * fs_ori)
    y = resample_poly(data, fs_down, fs_ori)
    y = resample_poly(y, fs_ori, fs_down)
    if len(y) != len(data):
        y = align_length(data, y)
    return y


def _limit(v, high, low):
    """ssr_eval/lowpass.py:147-153."""
    return high if v > high else low if v < low else int(v)


def lowpass_filter(x, highcut, fs, order, ftype):
    """ssr_eval/lowpass.py:94-131 (the trailing discarded subsampling call is omitted)."""
    hi = highcut / (0.5 * fs)
    if ftype == "butter":
        sos = butter(order, hi, btype="low", output="sos")
    elif ftype == "cheby1":
        sos = cheby1(order, 0.1, hi, btype="low", output="sos")
    elif ftype == "ellip":
        sos = ellip(order, 0.1, 60, hi, btype="low", output="sos")
    elif ftype == "bessel":
        sos = bessel(order, hi, btype="low", output="sos")
    else:
        raise Exception(f"The lowpass filter {ftype} is not supported!")
    y = sosfiltfilt(sos, x)
    if len(y) != len(x):
        y = align_length(x, y)
    return y


def lowpass(data, highcut, fs, order=5, _type="butter"):
    """ssr_eval/lowpass.py:156-196 (substring dispatch, order clamp, 1-D check)."""
    order = _limit(order, high=10, low=2)
    if len(list(data.shape)) != 1:
        raise ValueError("Error (chebyshev_lowpass_filter): Data " + str(data.shape)
                         + " should be type 1d time array, (samples,) , can not be (samples, 1)")
    for name in ("butter", "cheby1", "ellip", "bessel"):
        if _type in name:
            return lowpass_filter(data, int(highcut), fs, order, name)
    if _type in "subsampling":
        return subsampling(data, lowpass_ratio=highcut / int(fs / 2))
    if _type in "stft_hard":
        return stft_hard_lowpass_v0(data, lowpass_ratio=highcut / int(fs / 2))
    raise ValueError("Error: Unexpected filter type " + _type)


def librosa_resample_polyphase(y, orig_sr, target_sr):
    """librosa.resample(y, orig_sr, target_sr, res_type="polyphase") of librosa 0.9.x
    (call site ssr_eval/eval.py:144-150): gcd-reduced scipy resample_poly, length fixed to
    ceil(L * target/orig) (cut or zero-pad), no rescale, input dtype kept."""
    if orig_sr == target_sr:
        return y
    n_samples = int(np.ceil(y.shape[-1] * float(target_sr) / orig_sr))
    g = np.gcd(int(orig_sr), int(target_sr))
    y_hat = resample_poly(y, int(target_sr) // g, int(orig_sr) // g, axis=-1)
    n = y_hat.shape[-1]
    if n > n_samples:
        y_hat = y_hat[:n_samples]
    elif n < n_samples:
        y_hat = np.pad(y_hat, (0, n_samples - n), mode="constant")
    return np.asarray(y_hat, dtype=y.dtype)
