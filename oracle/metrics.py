"""Metric formulas of ssr_eval restated on CPU -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ssr_eval/metrics.py:51-132 and ssr_eval/utils.py:7,24-28,43-44,68-92 (torch float32
ops, same upstream torch as the reference) and scikit-image 0.19.x ``structural_similarity``
(third-party, absent here: restated on top of the real ``scipy.ndimage.uniform_filter``).
"""
import numpy as np
import torch
from scipy.ndimage import uniform_filter

from .stft import stft_mag

EPS = 1e-12  # ssr_eval/metrics.py:12, ssr_eval/utils.py:7


def to_log(x):
    """ssr_eval/utils.py:43-44."""
    return torch.log10(x + 1e-12)


def pow_p_norm(signal):
    """ssr_eval/utils.py:68-76: squared 2-norm over every dim but 0 (keepdim)."""
    dims = list(range(1, signal.dim()))
    return torch.norm(signal, p=2, dim=dims, keepdim=True) ** 2


def pow_norm(s1, s2):
    """ssr_eval/utils.py:85-92: sum(s1*s2) over every dim but 0,1 (keepdim)."""
    dims = list(range(2, s1.dim()))
    return torch.sum(s1 * s2, dim=dims, keepdim=True)


def energy_unify(estimated, original):
    """ssr_eval/utils.py:79-82: project the estimate on the target."""
    target = pow_norm(estimated, original) * original
    target = target / (pow_p_norm(original) + EPS)
    return estimated, target


def lsd(est, target):
    """ssr_eval/metrics.py:109-112 (est, target: (1,1,T,F) float32 magnitudes)."""
    ratio = torch.log10(target ** 2 / ((est + EPS) ** 2) + EPS) ** 2
    val = torch.mean(torch.mean(ratio, dim=3) ** 0.5, dim=2)
    return val[..., None, None]


def sispec(est, target):
    """ssr_eval/metrics.py:114-121."""
    output, target = energy_unify(est, target)
    noise = output - target
    sp = 10 * torch.log10(pow_p_norm(target) / (pow_p_norm(noise) + EPS) + EPS)
    return torch.sum(sp) / sp.size()[0]


def ssim_skimage(im1, im2, win_size=7):
    """skimage.metrics.structural_similarity(im1, im2, win_size=7) of scikit-image 0.19.x with
    every other argument at its default (call site ssr_eval/metrics.py:131): uniform 7x7 window,
    sample covariance (NP/(NP-1)), K1=0.01, K2=0.03, data_range from the dtype
    (float -> dtype_range (-1, 1) -> 2.0), image dtype preserved for float32, edge strip of
    (win_size-1)//2 cropped, float64 mean.  PARITY UNPINNED (third-party, not installable)."""
    assert im1.shape == im2.shape and im1.ndim == 2
    if min(im1.shape) < win_size:
        raise ValueError("win_size exceeds image extent.")
    ftype = np.float32 if im1.dtype == np.float32 else np.float64
    im1 = im1.astype(ftype, copy=False)
    im2 = im2.astype(ftype, copy=False)
    data_range = 2.0
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (0.01 * data_range) ** 2
    C2 = (0.03 * data_range) ** 2
    A1, A2 = 2 * ux * uy + C1, 2 * vxy + C2
    B1, B2 = ux ** 2 + uy ** 2 + C1, vx + vy + C2
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return S[pad:-pad, pad:-pad].mean(dtype=np.float64)


def ssim(est, target):
    """ssr_eval/metrics.py:123-132 (loops batch and channel, both 1 here)."""
    t, o = target.numpy(), est.numpy()
    res = np.zeros([o.shape[0], o.shape[1]])
    for b in range(o.shape[0]):
        for c in range(o.shape[1]):
            res[b, c] = ssim_skimage(o[b, c], t[b, c], win_size=7)
    return torch.tensor(res)[..., None, None]


class AudioMetricsOracle:
    """ssr_eval/metrics.py:15-107 for array inputs."""

    def __init__(self, rate, n_fft=None, hop_length=None):
        self.rate = rate
        # ssr_eval/metrics.py:18-19; overridable for BASELINE config 2 (n_fft 2048 / hop 512).
        self.hop_length = int(rate / 100) if hop_length is None else hop_length
        self.n_fft = int(2048 / (44100 / rate)) if n_fft is None else n_fft

    def wav_to_spectrogram(self, wav):
        return torch.tensor(stft_mag(wav, self.n_fft, self.hop_length)[None, None, ...])

    def evaluation(self, est, target, file=None, which=("lsd", "log_sispec", "sispec", "ssim")):
        if type(est) != type(target):
            raise ValueError("The input value should either both be numpy array or strings")
        assert est.ndim == 1 and target.ndim == 1
        assert abs(target.shape[0] - est.shape[0]) < 100
        n = min(target.shape[0], est.shape[0])
        target, est = target[:n], est[:n]
        tsp = self.wav_to_spectrogram(target)
        esp = self.wav_to_spectrogram(est)
        out = {}
        if "lsd" in which:
            out["lsd"] = float(lsd(esp.clone(), tsp.clone()))
        if "log_sispec" in which:
            out["log_sispec"] = float(sispec(to_log(esp.clone()), to_log(tsp.clone())))
        if "sispec" in which:
            out["sispec"] = float(sispec(esp.clone(), tsp.clone()))
        if "ssim" in which:
            out["ssim"] = float(ssim(esp.clone(), tsp.clone()))
        return out


def evaluation(est, target, rate=None, n_fft=None, hop=None, which=("lsd", "log_sispec", "sispec", "ssim")):
    """Functional form: metrics of one (est, target) pair of float32 waveforms."""
    m = AudioMetricsOracle(rate if rate is not None else 44100, n_fft=n_fft, hop_length=hop)
    return m.evaluation(est, target, None, which=which)


def evaluation_exact_reductions(est, target, n_fft, hop):
    """Secondary checker: the SAME float32 spectrograms and float32 element-wise formulas as the
    reference, but every reduction (sum / mean / norm) carried out in float64.  The reference's
    float32 reductions over T*F ~ 5e5 elements are themselves ~1e-5 relative noisy (4.5e-4 dB on
    sispec at L = 240000, DESIGN.md "Numerics"), so this is the tighter statement of what the
    formulas mean; the CUDA path (float64 accumulators) must agree with it to ~1e-6."""
    n = min(len(est), len(target))
    T = torch.tensor(stft_mag(np.asarray(target[:n], np.float32), n_fft, hop))
    E = torch.tensor(stft_mag(np.asarray(est[:n], np.float32), n_fft, hop))
    out = {}
    term = (torch.log10(T ** 2 / ((E + EPS) ** 2) + EPS) ** 2).double()   # float32 element-wise
    out["lsd"] = float(torch.mean(torch.mean(term, dim=1) ** 0.5))

    def _sispec64(e, t):
        s_et, s_tt = (e * t).sum(), (t * t).sum()
        tp = s_et * t / (s_tt + EPS)
        nn = ((e - tp) ** 2).sum()
        return float(10 * torch.log10((tp ** 2).sum() / (nn + EPS) + EPS))
    out["log_sispec"] = _sispec64(torch.log10(E + 1e-12).double(), torch.log10(T + 1e-12).double())
    out["sispec"] = _sispec64(E.double(), T.double())
    return out


def dict_mean(dict_list):
    """ssr_eval/utils.py:24-28 (numpy float64 mean per key)."""
    return {k: np.mean([d[k] for d in dict_list]) for k in dict_list[0].keys()}
