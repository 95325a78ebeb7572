"""resampy's ``kaiser_best`` resampler restated on CPU -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``librosa.load(path, sr=...)`` (ssr_eval/eval.py:242, ssr_eval/metrics.py:22-23) resamples with
``librosa.resample(..., res_type="kaiser_best")`` in librosa 0.9.x, which is ``resampy.resample(filter="kaiser_best")``
followed by ``librosa.util.fix_length(size=ceil(n * ratio))``.  resampy (third-party, BSD/ISC, NOT installed in this
image, un-vendored by the reference: setup.py lists only ``librosa``) is restated here from its published algorithm,
resampy 0.3 / 0.4 ``resampy/filters.py`` (``sinc_window``) and ``resampy/interpn.py`` (``_resample_loop``):

* the ``kaiser_best`` table is ``sinc_window(num_zeros=64, precision=9, window=kaiser(beta=14.769656459379492),
  rolloff=0.9475937167399596)``: 512 samples per zero crossing, 32769 float64 values of the right half of
  ``rolloff * sinc(rolloff * t) * kaiser(2n+1, beta)``  (resampy ships it precomputed as ``kaiser_best.npz``);
* an output sample at time t = i / ratio reads the table at ``frac * 512`` (+ ``k * index_step``) with LINEAR
  interpolation between neighbouring entries, ``index_step = int(scale * 512)`` -- truncated, so a down-sampling
  filter is not exactly a dilated copy of the prototype -- ``scale = min(1, ratio)``, left wing then right wing;
* the output has ``int(n * ratio)`` samples and the dtype of the input; the accumulation ``y[t] += weight * x[..]``
  adds a float64 product into the float32 output element (one rounding per tap).

PARITY UNPINNED: there is no resampy here to run against and the reference has no fixture for it; the restatement is
cross-checked against an exact windowed-sinc evaluation (tests/test_oracle.py).
"""
from functools import lru_cache

import numpy as np
from scipy.signal.windows import kaiser

KAISER_BEST = dict(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492)


@lru_cache(maxsize=2)
def sinc_window(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492):
    """resampy.filters.sinc_window with a Kaiser taper -> (half window, samples per zero crossing, rolloff)."""
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits, rolloff


def resample(x, sr_orig, sr_new):
    """resampy.resample(x, sr_orig, sr_new, filter="kaiser_best") for a 1-D signal (interpn._resample_loop)."""
    x = np.asarray(x)
    if sr_orig == sr_new:
        return x.copy()
    sample_ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[0] * sample_ratio)
    dtype = x.dtype if np.issubdtype(x.dtype, np.floating) else np.float32
    interp_win, num_table, _ = sinc_window(**KAISER_BEST)
    if sample_ratio < 1:
        interp_win = sample_ratio * interp_win
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    scale = min(1.0, sample_ratio)
    time_increment = 1.0 / sample_ratio
    t_out = np.arange(n_out) * time_increment
    index_step = int(scale * num_table)
    nwin = interp_win.shape[0]
    n_orig = x.shape[0]
    y = np.zeros(n_out, dtype=dtype)
    xd = x.astype(np.float64)
    for t in range(n_out):
        time_register = t_out[t]
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        acc = y.dtype.type(0)
        i_max = min(n + 1, (nwin - offset) // index_step)
        idx = offset + index_step * np.arange(i_max)
        w = interp_win[idx] + eta * interp_delta[idx]
        for i in range(i_max):  # the output element is rounded to its dtype after every tap
            acc = y.dtype.type(np.float64(acc) + w[i] * xd[n - i])
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_orig - n - 1, (nwin - offset) // index_step)
        idx = offset + index_step * np.arange(k_max)
        w = interp_win[idx] + eta * interp_delta[idx]
        for k in range(k_max):
            acc = y.dtype.type(np.float64(acc) + w[k] * xd[n + k + 1])
        y[t] = acc
    return y


def librosa_load_resample(x, sr_orig, sr_new):
    """librosa.resample(x, sr_orig, sr_new, res_type="kaiser_best") of librosa 0.9.x: resampy, then
    fix_length to ceil(n * ratio) (zero padding; resampy returns int(n * ratio) samples)."""
    x = np.asarray(x)
    if sr_orig == sr_new:
        return x
    ratio = float(sr_new) / sr_orig
    n_samples = int(np.ceil(x.shape[-1] * ratio))
    y = resample(x, sr_orig, sr_new)
    if len(y) < n_samples:
        y = np.pad(y, (0, n_samples - len(y)))
    return np.asarray(y[:n_samples], dtype=x.dtype if np.issubdtype(x.dtype, np.floating) else np.float32)
