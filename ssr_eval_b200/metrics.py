"""``AudioMetrics`` with the reference's interface (ssr_eval/metrics.py:15-132), computed by the
fused sm_100a STFT+metrics kernels.  No CPU path."""
import numpy as np

from . import _native as N
from .engine import StftMetrics

EPS = 1e-12


def _as_wave(x):
    """float32 or float64 waveform.  float64 stays float64: the reference's IIR low-pass filters return float64
    estimates (scipy sosfiltfilt), soundfile.read returns float64 by default, and librosa / torch keep such a signal in
    float64 end to end (dtype_r2c, type promotion) -- the kernels reproduce that (include/ssr_b200.h,
    ssr_stft_metrics_batched_f64est / _f64).  Any other dtype is cast to float32."""
    a = np.asarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float32)
    return a


class AudioMetrics:
    load_res_type = "kaiser_best"  # resampler of read(): what librosa 0.9's librosa.load(sr=...) uses

    def __init__(self, rate, n_fft=None, hop_length=None):
        """ssr_eval/metrics.py:16-19: hop = int(rate/100), n_fft = int(2048/(44100/rate)).
        ``n_fft`` / ``hop_length`` overrides exist for BASELINE config 2 (2048 / 512)."""
        self.rate = rate
        self.hop_length = int(rate / 100) if hop_length is None else int(hop_length)
        self.n_fft = int(2048 / (44100 / rate)) if n_fft is None else int(n_fft)
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            self._engine = StftMetrics(self.n_fft, self.hop_length)
        return self._engine

    def read(self, est, target):
        """ssr_eval/metrics.py:21-24 (librosa.load(sr=rate, mono=True): resampy's kaiser_best when the file's rate
        differs, see audio_io.load_audio)."""
        from .audio_io import load_audio
        e, _ = load_audio(est, sr=self.rate, res_type=self.load_res_type)
        t, _ = load_audio(target, sr=self.rate, res_type=self.load_res_type)
        return e, t

    def wav_to_spectrogram(self, wav):
        """ssr_eval/metrics.py:26-30: |STFT| as a (1, 1, T, F) float32 torch tensor (on the host)."""
        import torch
        return torch.from_numpy(self.engine.magnitude([np.asarray(wav, dtype=np.float32)])[0])[None, None, ...]

    @staticmethod
    def _check_pair(est, target):
        """Type / shape / length checks and truncation of metrics.py:64-90."""
        if type(est) != type(target):
            raise ValueError("The input value should either both be numpy array or strings")
        assert len(list(est.shape)) == 1 and len(list(target.shape)) == 1, (
            "The input numpy array shape should be [samples,]. Got input shape %s and %s. "
            % (est.shape, target.shape))
        assert abs(target.shape[0] - est.shape[0]) < 100, (
            "Error: Shape mismatch between target and estimation %s and %s"
            % (str(target.shape), str(est.shape)))
        n = min(target.shape[0], est.shape[0])
        return _as_wave(est[:n]), _as_wave(target[:n])

    def evaluation(self, est, target, file=None):
        """Metrics of one (est, target) pair -> {"lsd","log_sispec","sispec","ssim"} floats
        (ssr_eval/metrics.py:51-107).  ``file`` is accepted and unused, as in the reference."""
        if type(est) != type(target):
            raise ValueError("The input value should either both be numpy array or strings")
        if isinstance(est, str):
            est, target = self.read(est, target)
        return self.evaluation_batch([est], [target])[0]

    def evaluation_batch(self, est_list, target_list, flags=N.METRIC_ALL):
        """Batched form used by SSR_Eval_Helper.evaluate: one launch sequence for all pairs."""
        pairs = [self._check_pair(e, t) for e, t in zip(est_list, target_list)]
        if flags & N.METRIC_SSIM:
            for e, _ in pairs:
                # skimage.structural_similarity(win_size=7) raises on images with a side < 7
                if 1 + (len(e) + 2 * (self.n_fft // 2) - self.n_fft) // self.hop_length < 7 or self.n_fft // 2 + 1 < 7:
                    raise ValueError("win_size exceeds image extent. Either ensure that your images are "
                                     "at least 7x7; or pass win_size explicitly in the function call, "
                                     "with an odd value less than or equal to the smaller side of your images.")
        vals = self.engine.metrics([p[0] for p in pairs], [p[1] for p in pairs], flags)
        out = []
        for row in vals:
            out.append({name: float(row[i]) for i, name in enumerate(N.METRIC_NAMES)
                        if flags & (1 << i)})
        return out
