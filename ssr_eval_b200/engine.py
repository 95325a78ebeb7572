"""Batch engine: ragged utterance batches in HBM -> the sm_100a kernels behind the C ABI.

PyTorch is used for device memory, streams and (in dist.py) torch.distributed only; all arithmetic
of the hot path happens in libssr_b200.so.  Everything here needs a CUDA device -- no CPU fallback.
"""
import ctypes
from math import gcd

import numpy as np
import torch

from . import _native as N


def _require_cuda():
    if not torch.cuda.is_available():
        raise N.NativeError("ssr_eval_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def offsets_of(lengths):
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=off[1:])
    return off


def pack_ragged(arrays, pinned=False, dtype=torch.float32):
    """Concatenate 1-D float arrays into one float32 (or float64) host buffer + int64 offsets."""
    off = offsets_of([len(a) for a in arrays])
    flat = torch.empty(int(off[-1]), dtype=dtype, pin_memory=pinned)
    fn = flat.numpy()
    for a, s, e in zip(arrays, off[:-1], off[1:]):
        fn[s:e] = a
    return flat, off


class _Workspace:
    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        return self.buf


class StftMetrics:
    """K1+K2: batched STFT -> {lsd, log_sispec, sispec, ssim} (ssr_eval/metrics.py:92-132)."""

    # cap of the spectrogram workspace per launch sequence (SSIM only).  Measured: L2-sized sub-batches
    # (96 MB) lose more to launch overhead than the HBM round trip of 7.7 MB/pair costs at 6.5 TB/s.
    SSIM_SUBBATCH_BYTES = 8 << 30

    def __init__(self, n_fft, hop, window=None):
        _require_cuda()
        self.n_fft, self.hop = int(n_fft), int(hop)
        self.n_bins = self.n_fft // 2 + 1
        self._plan = ctypes.c_void_p()
        w = None
        if window is not None:
            w = np.ascontiguousarray(window, dtype=np.float64)
            assert w.shape == (self.n_fft,)
        N.check(N.lib().ssr_stft_plan_create(ctypes.byref(self._plan), self.n_fft, self.hop,
                                             _np_ptr(w) if w is not None else None), "ssr_stft_plan_create")
        self._ws = _Workspace()

    def __del__(self):
        try:
            if self._plan:
                N.lib().ssr_stft_plan_destroy(self._plan)
        except Exception:
            pass

    def num_frames(self, length):
        return int(N.lib().ssr_stft_num_frames(self._plan, int(length)))

    def _run(self, est_dev, tgt_dev, off_np, off_dev, out_dev, flags):
        n = len(off_np) - 1
        need = N.lib().ssr_stft_metrics_workspace_bytes(self._plan, _np_ptr(off_np), n, flags)
        if need == 0:
            raise N.NativeError("workspace query failed: " + (N.lib().ssr_last_error() or b"").decode())
        ws = self._ws.get(need, est_dev.device)
        if tgt_dev.dtype == torch.float64:
            fn = N.lib().ssr_stft_metrics_batched_f64
        else:
            fn = N.lib().ssr_stft_metrics_batched_f64est if est_dev.dtype == torch.float64 else N.lib().ssr_stft_metrics_batched
        N.check(fn(self._plan, _ptr(est_dev), _ptr(tgt_dev), _np_ptr(off_np), _ptr(off_dev), n, flags,
                   _ptr(out_dev), _ptr(ws), ws.numel(), _stream()), "ssr_stft_metrics_batched")

    def metrics_device(self, est_dev, tgt_dev, offsets, flags=N.METRIC_ALL, offsets_dev=None, out=None):
        """est_dev/tgt_dev: flat float32 CUDA tensors (ragged, same offsets); est_dev may be float64 (the
        reference's float64-estimate arithmetic, see include/ssr_b200.h). Returns (n,4) float64
        CUDA tensor [lsd, log_sispec, sispec, ssim] (NaN where not requested). Asynchronous."""
        off_np = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(off_np) - 1
        assert est_dev.is_cuda and tgt_dev.is_cuda
        assert est_dev.dtype in (torch.float32, torch.float64) and tgt_dev.dtype in (torch.float32, torch.float64)
        if tgt_dev.dtype == torch.float64 and est_dev.dtype != torch.float64:
            est_dev = est_dev.double()  # a float64 target is scored in float64 together with the (widened) estimate
        assert est_dev.numel() >= off_np[-1] and tgt_dev.numel() >= off_np[-1]
        dev = est_dev.device
        if offsets_dev is None:
            offsets_dev = torch.from_numpy(off_np).to(dev)
        if out is None:
            out = torch.empty((n, 4), dtype=torch.float64, device=dev)
        if not (flags & N.METRIC_SSIM):
            self._run(est_dev, tgt_dev, off_np, offsets_dev, out, flags)
            return out
        # SSIM needs both magnitude spectrograms: run sub-batches whose spectrograms fit in L2
        frames = np.array([self.num_frames(l) for l in np.diff(off_np)], dtype=np.int64)
        bytes_per = frames * self.n_bins * 8
        s = 0
        while s < n:
            e, acc = s, 0
            while e < n and (e == s or acc + bytes_per[e] <= self.SSIM_SUBBATCH_BYTES):
                acc += bytes_per[e]
                e += 1
            base = int(off_np[s])
            sub_off = off_np[s:e + 1] - base
            sub_off_dev = offsets_dev[s:e + 1] - base if s else offsets_dev[:e + 1]
            self._run(est_dev[base:], tgt_dev[base:], np.ascontiguousarray(sub_off), sub_off_dev.contiguous(),
                      out[s:e], flags)
            s = e
        return out

    def metrics(self, est_list, tgt_list, flags=N.METRIC_ALL):
        """Host entry: lists of 1-D float arrays (already truncated to equal length per pair).
        Returns (n,4) float64 numpy.  Pairs whose estimate is a float64 array are scored with the reference's
        float64-estimate arithmetic (one extra launch sequence for them)."""
        assert len(est_list) == len(tgt_list) and len(est_list) > 0
        for a, b in zip(est_list, tgt_list):
            assert len(a) == len(b)
        kinds = [(np.asarray(a).dtype == np.float64, np.asarray(b).dtype == np.float64) for a, b in zip(est_list, tgt_list)]
        if len(set(kinds)) > 1:  # float32 / float64 pairs keep their own arithmetic: one launch sequence per kind
            out = np.empty((len(est_list), 4), dtype=np.float64)
            for want in sorted(set(kinds)):
                idx = [i for i, f in enumerate(kinds) if f == want]
                out[idx] = self.metrics([est_list[i] for i in idx], [tgt_list[i] for i in idx], flags)
            return out
        e64, t64 = kinds[0]
        e_h, off = pack_ragged(est_list, pinned=True, dtype=torch.float64 if (e64 or t64) else torch.float32)
        t_h, _ = pack_ragged(tgt_list, pinned=True, dtype=torch.float64 if t64 else torch.float32)
        e_d = e_h.cuda(non_blocking=True)
        t_d = t_h.cuda(non_blocking=True)
        return self.metrics_device(e_d, t_d, off, flags).cpu().numpy()

    def magnitude(self, wav_list):
        """|STFT| of each utterance: list of (T_i, F) float32 numpy arrays (metrics.py:26-30)."""
        x_h, off = pack_ragged(wav_list, pinned=True)
        x_d = x_h.cuda(non_blocking=True)
        n = len(wav_list)
        frames = [self.num_frames(len(w)) for w in wav_list]
        spec = torch.empty(int(sum(frames)) * self.n_bins, dtype=torch.float32, device=x_d.device)
        off_dev = torch.from_numpy(off).to(x_d.device)
        need = N.lib().ssr_stft_metrics_workspace_bytes(self._plan, _np_ptr(off), n, 0)
        ws = self._ws.get(need, x_d.device)
        N.check(N.lib().ssr_stft_magnitude_batched(self._plan, _ptr(x_d), _np_ptr(off), _ptr(off_dev), n,
                                                   _ptr(spec), _ptr(ws), ws.numel(), _stream()),
                "ssr_stft_magnitude_batched")
        s = spec.cpu().numpy()
        out, pos = [], 0
        for T in frames:
            out.append(s[pos:pos + T * self.n_bins].reshape(T, self.n_bins))
            pos += T * self.n_bins
        return out


def resample_poly_taps(up, down, dtype=np.float32):
    """The FIR scipy.signal.resample_poly designs for window=("kaiser", 5.0): firwin cast to the
    input dtype, then scaled by `up` (scipy/signal/_signaltools.py resample_poly)."""
    from scipy.signal import firwin
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(dtype)
    h *= up
    return h


def kaiser_best_taps(up, down, dtype=np.float32, num_zeros=64, rolloff=0.9475937167399596, beta=14.769656459379492):
    """Prototype FIR of a band-limited (Kaiser-windowed sinc) resampler with resampy's ``kaiser_best`` parameters
    (64 zero crossings, roll-off 0.9476, beta 14.77 -- what librosa 0.9's ``librosa.load(sr=...)`` uses), sampled on
    the polyphase grid of ``resample_poly(x, up, down)`` and already scaled like scipy's ``h * up``:
    taps[m + M] = g(m / up), g(t) = c * sinc(c t) * kaiser(c t / num_zeros), c = min(1, up / down) * rolloff, t in
    input samples.  PARITY UNPINNED: resampy itself interpolates a 512-per-zero-crossing table of this kernel
    linearly; this is the exact kernel (torchaudio documents the same parameters as its kaiser_best equivalent)."""
    up, down = int(up), int(down)
    c = min(1.0, up / down) * rolloff
    half = int(np.ceil(num_zeros / c * up))  # support |c t| <= num_zeros
    t = np.arange(-half, half + 1, dtype=np.float64) / up
    u = np.clip(c * t / num_zeros, -1.0, 1.0)
    win = np.i0(beta * np.sqrt(1.0 - u * u)) / np.i0(beta)
    win[np.abs(c * t) > num_zeros] = 0.0
    return (c * np.sinc(c * t) * win).astype(dtype)


_RESAMPY_TABLE = None


def resampy_kaiser_best_bank(up, down):
    """Polyphase bank of resampy's ``kaiser_best`` resampler -- what librosa 0.9's ``librosa.load(path, sr=...)`` runs
    (ssr_eval/eval.py:242, ssr_eval/metrics.py:22-23) -- for the rational ratio up / down, restated from resampy 0.3 /
    0.4 (``filters.sinc_window``, ``interpn._resample_loop``): a table of the right half of
    ``rolloff * sinc(rolloff t) * kaiser(beta)`` with 512 samples per zero crossing (64 crossings) that is read with
    LINEAR interpolation at ``offset = int(frac * 512)``, ``eta = frac * 512 - offset`` and a TRUNCATED step
    ``index_step = int(scale * 512)`` (scale = min(1, ratio); the table is multiplied by the ratio when down-sampling).
    For a rational ratio the fractional position of output j depends only on (j * down) % up, so the weights form a
    polyphase bank: returns (bank float32 [up][K], K, lead) for ``ssr_resample_plan_create_bank`` -- tap k of a phase
    multiplies x[floor(j * down / up) + lead - k]: k < lead is resampy's right wing (x[n+1], x[n+2], ...), k >= lead
    its left wing (x[n], x[n-1], ...).  Weights are evaluated in float64 like resampy and rounded to float32 once."""
    global _RESAMPY_TABLE
    from scipy.signal.windows import kaiser
    up, down = int(up), int(down)
    if _RESAMPY_TABLE is None:
        num_zeros, num_table, rolloff, beta = 64, 512, 0.9475937167399596, 14.769656459379492
        n = num_table * num_zeros
        _RESAMPY_TABLE = kaiser(2 * n + 1, beta)[n:] * (rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1)))
    ratio = up / down
    win = _RESAMPY_TABLE * ratio if ratio < 1 else _RESAMPY_TABLE
    delta = np.zeros_like(win)
    delta[:-1] = np.diff(win)
    scale = min(1.0, ratio)
    num_table, nwin = 512, win.shape[0]
    index_step = int(scale * num_table)
    wings = []
    for p in range(up):
        frac = scale * (p / up)            # output j = (j * down) // up + p / up input samples
        out = []
        for f in (frac, scale - frac):     # left wing, right wing
            index_frac = f * num_table
            offset = int(index_frac)
            eta = index_frac - offset
            idx = offset + index_step * np.arange((nwin - offset) // index_step)
            out.append(win[idx] + eta * delta[idx])
        wings.append(out)
    lw = max(len(w[0]) for w in wings)
    rw = max(len(w[1]) for w in wings)
    bank = np.zeros((up, lw + rw), dtype=np.float64)
    for p, (left, right) in enumerate(wings):
        bank[p, rw - len(right):rw] = right[::-1]
        bank[p, rw:rw + len(left)] = left
    return np.ascontiguousarray(bank, dtype=np.float32), lw + rw, rw


class PolyphaseResampler:
    """K3: scipy.signal.resample_poly(x, up, down) for float32 batches (dtype=np.float64: float64 batches --
    scipy keeps the input dtype, so a float64 waveform is filtered with float64 taps in float64)."""

    def __init__(self, up, down, dtype=np.float32, taps=None, bank=None):
        """``taps``: optional prototype FIR (odd length, already scaled like scipy's ``h * up``) instead of the
        Kaiser-5.0 ``firwin`` design of resample_poly, e.g. ``kaiser_best_taps(up, down)``.
        ``bank``: "resampy_kaiser_best" -- the explicit polyphase bank of resampy's table-interpolating resampler
        (``resampy_kaiser_best_bank``); such a resampler returns floor(n * up / down) samples like resampy does."""
        _require_cuda()
        g = gcd(int(up), int(down))
        self.up, self.down = int(up) // g, int(down) // g
        self.identity = self.up == 1 and self.down == 1
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float32), np.dtype(np.float64))
        self._tdtype = torch.float64 if self.dtype == np.float64 else torch.float32
        self._plan = ctypes.c_void_p()
        self.floor_len = False
        if bank is not None and not self.identity:
            if bank != "resampy_kaiser_best" or self.dtype != np.float32 or taps is not None:
                raise ValueError("bank must be 'resampy_kaiser_best' (float32, no taps)")
            b, K, lead = resampy_kaiser_best_bank(self.up, self.down)
            N.check(N.lib().ssr_resample_plan_create_bank(ctypes.byref(self._plan), self.up, self.down, _np_ptr(b), K, lead),
                    "ssr_resample_plan_create_bank")
            self.floor_len = True
        elif not self.identity:
            if taps is None:
                taps = resample_poly_taps(self.up, self.down, dtype=self.dtype)
            taps = np.ascontiguousarray(taps, dtype=self.dtype)
            assert taps.ndim == 1 and len(taps) % 2 == 1
            create = N.lib().ssr_resample_plan_create_f64 if self.dtype == np.float64 else N.lib().ssr_resample_plan_create
            N.check(create(ctypes.byref(self._plan), self.up, self.down, _np_ptr(taps), len(taps)),
                    "ssr_resample_plan_create")

    def __del__(self):
        try:
            if self._plan:
                N.lib().ssr_resample_plan_destroy(self._plan)
        except Exception:
            pass

    def out_len(self, n_in):
        t = int(n_in) * self.up
        if self.floor_len:
            return t // self.down
        return t // self.down + (1 if t % self.down else 0)

    def resample_device(self, x_dev, in_offsets, in_offsets_dev=None):
        """Returns (y_dev, out_offsets numpy, out_offsets_dev). Asynchronous."""
        in_off = np.ascontiguousarray(in_offsets, dtype=np.int64)
        if in_offsets_dev is None:
            in_offsets_dev = torch.from_numpy(in_off).to(x_dev.device)
        if self.identity:
            return x_dev, in_off, in_offsets_dev
        n = len(in_off) - 1
        out_off = offsets_of([self.out_len(l) for l in np.diff(in_off)])
        out_off_dev = torch.from_numpy(out_off).to(x_dev.device)
        assert x_dev.dtype == self._tdtype
        y = torch.empty(int(out_off[-1]), dtype=self._tdtype, device=x_dev.device)
        run = N.lib().ssr_resample_poly_batched_f64 if self.dtype == np.float64 else N.lib().ssr_resample_poly_batched
        N.check(run(self._plan, _ptr(x_dev), _np_ptr(in_off), _ptr(in_offsets_dev), _ptr(y), _np_ptr(out_off),
                    _ptr(out_off_dev), n, _stream()), "ssr_resample_poly_batched")
        return y, out_off, out_off_dev

    def resample(self, wav_list):
        x_h, off = pack_ragged(wav_list, pinned=True, dtype=self._tdtype)
        y, out_off, _ = self.resample_device(x_h.cuda(non_blocking=True), off)
        yh = y.cpu().numpy()
        return [yh[s:e].copy() for s, e in zip(out_off[:-1], out_off[1:])]


class HardLowpass:
    """K4: stft_hard_lowpass_v0 (ssr_eval/lowpass.py:17-28) for float32 batches."""

    def __init__(self, n_fft=2048, hop=441):
        _require_cuda()
        self.n_fft, self.hop = int(n_fft), int(hop)
        self.n_bins = self.n_fft // 2 + 1
        self._plan = ctypes.c_void_p()
        N.check(N.lib().ssr_lowpass_plan_create(ctypes.byref(self._plan), self.n_fft, self.hop),
                "ssr_lowpass_plan_create")

    def __del__(self):
        try:
            if self._plan:
                N.lib().ssr_lowpass_plan_destroy(self._plan)
        except Exception:
            pass

    def cut_bin(self, lowpass_ratio):
        return int(self.n_bins * lowpass_ratio)  # lowpass.py:23

    def apply_device(self, x_dev, offsets, cut_bins, offsets_dev=None):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        if offsets_dev is None:
            offsets_dev = torch.from_numpy(off).to(x_dev.device)
        n = len(off) - 1
        cb = torch.as_tensor(np.asarray(cut_bins, dtype=np.int32)).to(x_dev.device)
        assert cb.numel() == n
        y = torch.empty(int(off[-1]), dtype=torch.float32, device=x_dev.device)
        N.check(N.lib().ssr_stft_hard_lowpass_batched(self._plan, _ptr(x_dev), _np_ptr(off), _ptr(offsets_dev), n,
                                                      _ptr(cb), _ptr(y), _stream()),
                "ssr_stft_hard_lowpass_batched")
        return y

    def apply(self, wav_list, lowpass_ratios):
        x_h, off = pack_ragged(wav_list, pinned=True)
        cuts = [self.cut_bin(r) for r in lowpass_ratios]
        y = self.apply_device(x_h.cuda(non_blocking=True), off, cuts).cpu().numpy()
        return [y[s:e].copy() for s, e in zip(off[:-1], off[1:])]


def torchlibrosa_dft_matrices(n_fft):
    """The float32 convolution kernels torchlibrosa's ``STFT`` / ``ISTFT`` modules build (the classes behind
    ssr_eval/dsp.py:21-39), restated from torchlibrosa/stft.py (``DFTBase.dft_matrix`` / ``idft_matrix``,
    ``STFT.__init__``, ``ISTFT.init_real_imag_conv`` / ``init_overlap_add_window``; 0.0.7-0.0.9):
        W  = np.power(exp(-2j*pi/n), x*y)               Wi = np.power(exp(+2j*pi/n), x*y) / n
        stft  conv_real / conv_imag = real / imag (W[:, :n/2+1] * hann[:, None]).T          -> (n/2+1, n)
        istft conv_real / conv_imag = real / imag (Wi * hann[None, :]).T                     -> (n, n) (out, in)
        ola_window = hann ** 2
    with the periodic Hann window, evaluated in float64 (complex128 ``np.power``) and cast to float32 as there.
    Returns (stft_w_real, stft_w_imag, istft_w_real, istft_w_imag, ola_window), C-contiguous float32."""
    n = int(n_fft)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)  # scipy.signal.get_window("hann", n, fftbins=True)
    x, y = np.meshgrid(np.arange(n), np.arange(n))
    W = np.power(np.exp(-2 * np.pi * 1j / n), x * y)
    Wi = np.power(np.exp(2 * np.pi * 1j / n), x * y) / n
    F = n // 2 + 1
    c = np.ascontiguousarray
    fw = W[:, :F] * hann[:, None]
    iw = Wi * hann[None, :]
    return (c(np.real(fw).T, dtype=np.float32), c(np.imag(fw).T, dtype=np.float32),
            c(np.real(iw).T, dtype=np.float32), c(np.imag(iw).T, dtype=np.float32), c(hann ** 2, dtype=np.float32))


class HardLowpassDense:
    """K4d: stft_hard_lowpass_v0 in the reference's own arithmetic -- dense float32 DFT / IDFT products in the
    accumulation order of torch's CPU convolutions (csrc/stft_lowpass_dense.cu).  ~100x the arithmetic of
    ``HardLowpass``; use it when the noise floor above the cutoff has to be the reference's (LSD / log-sispec of an
    unprocessed ``proc_fft_*`` input)."""

    WORKSPACE_BYTES = 3 << 30  # per launch sequence; the batch is chunked to fit

    def __init__(self, n_fft=2048, hop=441):
        _require_cuda()
        self.n_fft, self.hop = int(n_fft), int(hop)
        self.n_bins = self.n_fft // 2 + 1
        self._plan = ctypes.c_void_p()
        mats = torchlibrosa_dft_matrices(self.n_fft)
        N.check(N.lib().ssr_lowpass_dense_plan_create(ctypes.byref(self._plan), self.n_fft, self.hop,
                                                      *[_np_ptr(m) for m in mats]), "ssr_lowpass_dense_plan_create")
        self._ws = _Workspace()

    def __del__(self):
        try:
            if self._plan:
                N.lib().ssr_lowpass_dense_plan_destroy(self._plan)
        except Exception:
            pass

    def cut_bin(self, lowpass_ratio):
        return int(self.n_bins * lowpass_ratio)  # lowpass.py:23

    def apply_device(self, x_dev, offsets, cut_bins, offsets_dev=None):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        if offsets_dev is None:
            offsets_dev = torch.from_numpy(off).to(x_dev.device)
        n = len(off) - 1
        cb = torch.as_tensor(np.asarray(cut_bins, dtype=np.int32)).to(x_dev.device)
        assert cb.numel() == n and x_dev.dtype == torch.float32
        y = torch.empty(int(off[-1]), dtype=torch.float32, device=x_dev.device)
        need = N.lib().ssr_stft_hard_lowpass_dense_workspace_bytes(self._plan, _np_ptr(off), n)
        longest = N.lib().ssr_stft_hard_lowpass_dense_workspace_bytes(
            self._plan, _np_ptr(np.array([0, int(np.diff(off).max())], dtype=np.int64)), 1)
        ws = self._ws.get(max(min(need, self.WORKSPACE_BYTES), longest + 8 * (n + 64)), x_dev.device)
        N.check(N.lib().ssr_stft_hard_lowpass_dense_batched(self._plan, _ptr(x_dev), _np_ptr(off), _ptr(offsets_dev), n,
                                                            _ptr(cb), _ptr(y), _ptr(ws), ws.numel(), _stream()),
                "ssr_stft_hard_lowpass_dense_batched")
        return y

    def apply(self, wav_list, lowpass_ratios):
        x_h, off = pack_ragged(wav_list, pinned=True)
        cuts = [self.cut_bin(r) for r in lowpass_ratios]
        y = self.apply_device(x_h.cuda(non_blocking=True), off, cuts).cpu().numpy()
        return [y[s:e].copy() for s, e in zip(off[:-1], off[1:])]


class SpliceIstft:
    """K6: the STFT splice + ISTFT of BasicTestee.postprocessing (ssr_eval/eval.py:28-41)."""

    def __init__(self, n_fft=2048, hop=512):
        _require_cuda()
        self.n_fft, self.hop = int(n_fft), int(hop)
        self._plan = ctypes.c_void_p()
        N.check(N.lib().ssr_splice_plan_create(ctypes.byref(self._plan), self.n_fft, self.hop),
                "ssr_splice_plan_create")
        self._stft = StftMetrics(self.n_fft, self.hop)

    def __del__(self):
        try:
            if self._plan:
                N.lib().ssr_splice_plan_destroy(self._plan)
        except Exception:
            pass

    def cutoff_indices(self, x_list):
        """BasicTestee._get_cutoff_index (eval.py:21-31) per utterance: |STFT| summed over the frames,
        cumulative sum over the bins, last bin whose cumulative energy is below 97 % of the total."""
        mags = self._stft.magnitude(x_list)
        out = []
        for m in mags:
            energy = np.cumsum(np.sum(np.ascontiguousarray(m.T), axis=-1))  # (F,), as the reference's numpy ops
            level = energy[-1] * 0.97
            below = np.nonzero(energy[:0:-1] < level)[0]  # scans x[-1], x[-2], ..., x[1]
            out.append(int(energy.shape[0] - (below[0] + 1)) if len(below) else 0)
        return out

    def apply(self, x_list, out_list, cut_bins):
        assert len(x_list) == len(out_list) == len(cut_bins) and len(x_list) > 0
        for a, b in zip(x_list, out_list):
            if len(a) != len(b):
                raise ValueError("postprocessing: input and output need the same length (%d vs %d)" % (len(a), len(b)))
        x_h, off = pack_ragged(x_list, pinned=True)
        o_h, _ = pack_ragged(out_list, pinned=True)
        x_d, o_d = x_h.cuda(non_blocking=True), o_h.cuda(non_blocking=True)
        off_d = torch.from_numpy(off).to(x_d.device)
        cb = torch.as_tensor(np.asarray(cut_bins, dtype=np.int32)).to(x_d.device)
        y = torch.empty_like(o_d)
        N.check(N.lib().ssr_stft_splice_istft_batched(self._plan, _ptr(x_d), _ptr(o_d), _np_ptr(off), _ptr(off_d),
                                                      len(x_list), _ptr(cb), _ptr(y), _stream()),
                "ssr_stft_splice_istft_batched")
        yh = y.cpu().numpy()
        return [yh[s:e].copy() for s, e in zip(off[:-1], off[1:])]


def sosfiltfilt_batch(sos, wav_list):
    """K7: scipy.signal.sosfiltfilt(sos, x) (default odd padding) for a batch of float32 utterances ->
    list of float64 arrays.  Filter design, sosfilt_zi and the pad length are scipy's, on the host."""
    from scipy.signal import sosfilt_zi
    _require_cuda()
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    assert sos.ndim == 2 and sos.shape[1] == 6
    n_sections = sos.shape[0]
    ntaps = 2 * n_sections + 1
    ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    edge = 3 * ntaps  # scipy's default padlen
    for w in wav_list:
        if len(w) <= edge:
            raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % edge)
    zi = np.ascontiguousarray(sosfilt_zi(sos), dtype=np.float64)
    x_h, off = pack_ragged(wav_list, pinned=True)
    x_d = x_h.cuda(non_blocking=True)
    off_d = torch.from_numpy(off).to(x_d.device)
    y = torch.empty(int(off[-1]), dtype=torch.float64, device=x_d.device)
    need = N.lib().ssr_sosfiltfilt_workspace_bytes(_np_ptr(off), len(wav_list), edge)
    ws = torch.empty(max(int(need), 8), dtype=torch.uint8, device=x_d.device)
    N.check(N.lib().ssr_sosfiltfilt_batched(_np_ptr(sos), n_sections, _np_ptr(zi), edge, _ptr(x_d), _np_ptr(off),
                                            _ptr(off_d), len(wav_list), _ptr(y), _ptr(ws), ws.numel(), _stream()),
            "ssr_sosfiltfilt_batched")
    yh = y.cpu().numpy()
    return [yh[s:e].copy() for s, e in zip(off[:-1], off[1:])]


def xcorr_argmax_batch(a_list, x_list, workspace_bytes=2 << 30):
    """K8: ``np.argmax(scipy.signal.correlate(a, x))`` for a batch of equal-length float32 pairs (the alignment of the
    mp3 path, ssr_eval/eval.py:319) -> list of ints (scipy's 'full' index; the lag is index - (len(x) - 1))."""
    _require_cuda()
    assert len(a_list) == len(x_list) and len(a_list) > 0
    for a, x in zip(a_list, x_list):
        if len(a) != len(x):
            raise ValueError("xcorr: the two signals of a pair need the same length (%d vs %d)" % (len(a), len(x)))
        if 2 * len(x) - 1 > (1 << 20):
            raise ValueError("xcorr: utterances longer than 524288 samples are not supported")
    a_h, off = pack_ragged([np.asarray(a, dtype=np.float32) for a in a_list], pinned=True)
    x_h, _ = pack_ragged([np.asarray(x, dtype=np.float32) for x in x_list], pinned=True)
    a_d, x_d = a_h.cuda(non_blocking=True), x_h.cuda(non_blocking=True)
    off_d = torch.from_numpy(off).to(a_d.device)
    n = len(a_list)
    out = torch.empty(n, dtype=torch.int64, device=a_d.device)
    need = N.lib().ssr_xcorr_workspace_bytes(_np_ptr(off), n)
    longest = N.lib().ssr_xcorr_workspace_bytes(_np_ptr(np.array([0, int(np.diff(off).max())], dtype=np.int64)), 1)
    ws = torch.empty(max(min(int(need), int(workspace_bytes)), int(longest) + 4 * n + 256), dtype=torch.uint8, device=a_d.device)
    N.check(N.lib().ssr_xcorr_argmax_batched(_ptr(a_d), _ptr(x_d), _np_ptr(off), _ptr(off_d), n, _ptr(out), _ptr(ws),
                                             ws.numel(), _stream()), "ssr_xcorr_argmax_batched")
    return [int(v) for v in out.cpu().numpy()]


def pcm16_to_float_device(src_dev, out=None):
    """K0: int16 CUDA tensor -> float32 CUDA tensor, x / 32768 (what librosa.load / soundfile.read give the
    reference for a 16-bit wav).  Asynchronous on the current stream."""
    assert src_dev.is_cuda and src_dev.dtype == torch.int16 and src_dev.is_contiguous()
    if out is None:
        out = torch.empty(src_dev.numel(), dtype=torch.float32, device=src_dev.device)
    assert out.dtype == torch.float32 and out.numel() >= src_dev.numel()
    N.check(N.lib().ssr_pcm16_to_float(_ptr(src_dev), _ptr(out), src_dev.numel(), _stream()), "ssr_pcm16_to_float")
    return out


class HostPipeline:
    """Host-buffer entry point of K1/K2: (pinned) host batches are streamed to the GPU in chunks on a
    copy stream, double-buffered against the kernels, and the (n, 4) float64 result is read back.
    This is the path timed as ``e2e`` in bench.py.

    Either host buffer may be float32 or **int16** (16-bit PCM, the sample format of the wav files the
    reference loads): int16 chunks are uploaded as they are -- half the PCIe bytes -- and converted on the
    device by K0 (x / 32768, bit-identical to what librosa.load returns for such a file)."""

    def __init__(self, engine, max_pairs, max_len, chunk_pairs=64):
        _require_cuda()
        self.engine = engine
        self.chunk_samples = int(min(max_pairs, chunk_pairs)) * int(max_len)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.slots = [(torch.empty(self.chunk_samples, dtype=torch.float32, device=dev),
                       torch.empty(self.chunk_samples, dtype=torch.float32, device=dev)) for _ in range(2)]
        self.raw = [[None, None], [None, None]]  # int16 staging, allocated on first use
        self.copy_stream = torch.cuda.Stream()
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]

    def _raw(self, slot, which):
        if self.raw[slot][which] is None:
            self.raw[slot][which] = torch.empty(self.chunk_samples, dtype=torch.int16, device=self.slots[0][0].device)
        return self.raw[slot][which]

    def run(self, est_host, tgt_host, offsets, flags=N.METRIC_ALL):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(off) - 1
        for h in (est_host, tgt_host):
            if h.dtype not in (torch.float32, torch.int16):
                raise TypeError("host batches must be float32 or int16 (16-bit PCM), got %s" % h.dtype)
        dev = self.slots[0][0].device
        out = torch.empty((n, 4), dtype=torch.float64, device=dev)
        compute = torch.cuda.current_stream()
        # chunk boundaries and ALL rebased offset tables first, uploaded in one copy: a per-chunk upload from pageable
        # memory is a synchronous copy on the compute stream, i.e. the host would wait for the previous chunk's H2D
        # before it can queue the next one and the copy engine would idle in between (-8 % end to end, measured)
        chunks, s = [], 0
        while s < n:
            e = s + 1
            while e < n and off[e + 1] - off[s] <= self.chunk_samples:
                e += 1
            if int(off[e] - off[s]) > self.chunk_samples:
                raise ValueError("utterance longer than the pipeline slot")
            chunks.append((s, e))
            s = e
        rebased = [off[s:e + 1] - off[s] for s, e in chunks]
        starts = np.cumsum([0] + [len(r) for r in rebased])
        all_dev = torch.from_numpy(np.concatenate(rebased)).to(dev)
        for c, (s, e) in enumerate(chunks):
            a, b = int(off[s]), int(off[e])
            slot = c % 2
            dst = []
            with torch.cuda.stream(self.copy_stream):
                if c >= 2:
                    self.copy_stream.wait_event(self.free[slot])
                for which, host in enumerate((est_host, tgt_host)):
                    d = self._raw(slot, which) if host.dtype == torch.int16 else self.slots[slot][which]
                    d[:b - a].copy_(host[a:b], non_blocking=True)
                    dst.append(d)
                self.ready[slot].record(self.copy_stream)
            compute.wait_event(self.ready[slot])
            for which in range(2):
                if dst[which].dtype == torch.int16:
                    pcm16_to_float_device(dst[which][:b - a], out=self.slots[slot][which])
            se, st = self.slots[slot]
            self.engine.metrics_device(se, st, rebased[c], flags, offsets_dev=all_dev[starts[c]:starts[c + 1]], out=out[s:e])
            self.free[slot].record(compute)
        return out.cpu().numpy()
