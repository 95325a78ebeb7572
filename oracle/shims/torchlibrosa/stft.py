"""Stand-in for torchlibrosa.stft (0.0.7-0.0.9) -- forwards to the oracle restatement."""
import torch
import torch.nn as nn
from oracle.lowpass import TorchlibrosaSTFT, TorchlibrosaISTFT


class STFT(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", freeze_parameters=True):
        super().__init__()
        assert window == "hann" and center and pad_mode == "reflect" and win_length in (None, n_fft)
        self._impl = TorchlibrosaSTFT(n_fft, hop_length if hop_length is not None else n_fft // 4)

    def forward(self, x):
        return self._impl(x)


class ISTFT(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", freeze_parameters=True):
        super().__init__()
        assert window == "hann" and center and pad_mode == "reflect" and win_length in (None, n_fft)
        self._impl = TorchlibrosaISTFT(n_fft, hop_length if hop_length is not None else n_fft // 4)

    def forward(self, real_stft, imag_stft, length):
        return self._impl(real_stft, imag_stft, length)


def magphase(real, imag):
    mag = (real ** 2 + imag ** 2) ** 0.5
    return mag, real / torch.clamp(mag, 1e-10, float("inf")), imag / torch.clamp(mag, 1e-10, float("inf"))
