"""CPU, build container only (skipped where /root/reference is not mounted): the oracle's restatement of the in-repo
formulas against the REFERENCE'S OWN CODE executed live through oracle/shims, on random cases beyond the committed
golden fixtures.  The shims forward the third-party calls (librosa.stft, skimage SSIM, torchlibrosa) to the oracle, so
what is compared here is everything the reference itself implements: evaluation's checks and truncation, lsd, sispec,
to_log, energy_unify, the lowpass dispatcher, stft_hard_lowpass_v0, subsampling, the IIR wrappers, dict_mean."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ssr_eval")), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    names = ("ssr_eval", "ssr_eval.utils", "ssr_eval.dsp", "ssr_eval.metrics", "ssr_eval.lowpass")
    saved = {k: sys.modules.get(k) for k in names}
    pkg = types.ModuleType("ssr_eval")
    pkg.__path__ = [os.path.join(REF, "ssr_eval")]
    sys.modules["ssr_eval"] = pkg
    mods = {}
    for name in ("utils", "dsp", "metrics", "lowpass"):
        spec = importlib.util.spec_from_file_location("ssr_eval." + name, os.path.join(REF, "ssr_eval", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["ssr_eval." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    yield mods
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_metrics_random_pairs(ref):
    import oracle
    from ssr_eval_b200.synth import speech_like
    rng = np.random.default_rng(2026)
    for case in range(10):
        rate = int(rng.choice([16000, 24000, 44100, 48000]))
        L = int(rng.integers(rate // 4, rate))
        tgt = speech_like(L, rate, seed=int(rng.integers(1 << 30)))
        kind = case % 4
        if kind == 0:
            est = ref["lowpass"].lowpass(tgt, int(rng.integers(1000, rate // 2 - 500)), rate, order=1, _type="stft_hard")
        elif kind == 1:
            est = ref["lowpass"].lowpass(tgt, int(rng.integers(1000, rate // 2 - 500)), rate, order=int(rng.integers(2, 9)),
                                         _type=str(rng.choice(["butter", "cheby1", "ellip", "bessel"])))  # float64
        elif kind == 2:
            est = (tgt + 10.0 ** rng.uniform(-4, -1) * rng.standard_normal(L)).astype(np.float32)
        else:
            est = tgt[: L - int(rng.integers(1, 99))] * np.float32(0.7)   # length mismatch < 100 -> truncation
        want = ref["metrics"].AudioMetrics(rate).evaluation(est, tgt, "none")
        got = oracle.evaluation(est, tgt, rate=rate)
        for k in ("lsd", "log_sispec", "sispec", "ssim"):
            assert got[k] == pytest.approx(want[k], rel=1e-6, abs=1e-7), (case, rate, kind, k, got[k], want[k])
    m = ref["metrics"].AudioMetrics(44100)
    with pytest.raises(ValueError):
        m.evaluation(np.zeros(1000, np.float32), "a.wav", None)
    with pytest.raises(AssertionError):
        m.evaluation(np.zeros(1000, np.float32), np.zeros(1200, np.float32), None)


def test_lowpass_dispatch_random_settings(ref):
    import oracle
    from ssr_eval_b200.synth import speech_like
    rng = np.random.default_rng(7)
    for case in range(12):
        fs = int(rng.choice([16000, 44100, 48000]))
        x = speech_like(int(rng.integers(4000, 20000)), fs, seed=case)
        cutoff = int(rng.integers(500, fs // 2 - 500))
        _type = ["stft_hard", "subsampling", "butter", "cheby1", "ellip", "bessel", "stft", "sub", "but"][case % 9]
        order = int(rng.integers(0, 14))
        want = ref["lowpass"].lowpass(x, cutoff, fs, order=order, _type=_type)
        got = oracle.lowpass(x, cutoff, fs, order=order, _type=_type)
        assert got.dtype == want.dtype and got.shape == want.shape, (case, _type)
        assert np.array_equal(got, want), (case, _type, cutoff, fs, order, np.abs(got - want).max())
    with pytest.raises(ValueError):
        ref["lowpass"].lowpass(x[:, None], 1000, 44100, _type="butter")
    with pytest.raises(ValueError):
        oracle.lowpass(x[:, None], 1000, 44100, _type="butter")
