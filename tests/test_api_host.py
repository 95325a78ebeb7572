"""CPU: host logic of the drop-in API (no kernel launches)."""
import os

import numpy as np
import pytest

from ssr_eval_b200 import AudioMetrics, BasicTestee, SSR_Eval_Helper, lowpass
from ssr_eval_b200.engine import offsets_of, pack_ragged, resample_poly_taps
from ssr_eval_b200.lowpass import align_length, limit
from ssr_eval_b200.utils import dict_mean


def test_audio_metrics_parameters():
    # ssr_eval/metrics.py:18-19
    for rate, hop, n_fft in ((48000, 480, 2229), (44100, 441, 2048), (24000, 240, 1114), (16000, 160, 743)):
        m = AudioMetrics(rate)
        assert (m.hop_length, m.n_fft) == (hop, n_fft)
    m = AudioMetrics(48000, n_fft=2048, hop_length=512)
    assert (m.hop_length, m.n_fft) == (512, 2048)


def test_pair_checks():
    a = np.zeros(1000, np.float32)
    with pytest.raises(ValueError):
        AudioMetrics(44100).evaluation(a, "x.wav", None)
    with pytest.raises(AssertionError):
        AudioMetrics._check_pair(a, np.zeros(1100, np.float32))
    with pytest.raises(AssertionError):
        AudioMetrics._check_pair(a[:, None], a)
    e, t = AudioMetrics._check_pair(np.ones(950, np.float64), a)
    # a float64 estimate stays float64 (the reference scores it in float64), the target is float32
    assert len(e) == len(t) == 950 and e.dtype == np.float64 and t.dtype == np.float32


def test_lowpass_dispatch_errors_and_iir():
    x = np.random.default_rng(0).standard_normal(4000).astype(np.float32)
    with pytest.raises(ValueError):
        lowpass(x[:, None], 1000, 44100, _type="stft_hard")
    with pytest.raises(ValueError):
        lowpass(x, 1000, 44100, _type="nope")
    assert limit(1, 10, 2) == 2 and limit(11, 10, 2) == 10 and limit(5.0, 10, 2) == 5
    assert len(align_length(np.zeros(10), np.zeros(7))) == 10
    assert len(align_length(np.zeros(10), np.zeros(17))) == 10


def test_cutoff_doubling_mutates_callers_dict(tmp_path):
    setting = {"cutoff_freq": [4000, 12000]}
    h = SSR_Eval_Helper(BasicTestee(), 44100, 44100, evaluation_sr=48000, setting_fft=setting,
                        test_data_root=str(tmp_path))
    assert setting["cutoff_freq"] == [8000, 24000]  # eval.py:121-126
    assert h.setting_fft is setting
    with pytest.raises(AssertionError):
        SSR_Eval_Helper(BasicTestee(), 44100, 44100, evaluation_sr=96000, test_data_root=str(tmp_path))
    with pytest.raises(FileNotFoundError):
        SSR_Eval_Helper(BasicTestee(), 44100, 44100, test_data_root=str(tmp_path / "missing"))


def test_file_listing(tmp_path):
    d = tmp_path / "p360"
    d.mkdir()
    for n in ("a.wav", "b.flac", "c_proc_x.wav", ".DS_Store.wav", "d.txt"):
        (d / n).write_bytes(b"")
    h = SSR_Eval_Helper(BasicTestee(), 44100, 44100, test_data_root=str(tmp_path))
    assert sorted(h.get_test_file_list(str(d))) == ["a.wav", "b.flac"]
    (tmp_path / "x1").mkdir()
    assert h._speakers(-1) == ["p360"]


def test_ragged_packing_and_taps():
    off = offsets_of([3, 0, 5])
    assert off.tolist() == [0, 3, 3, 8]
    flat, off = pack_ragged([np.arange(3), np.arange(5)])
    assert flat.tolist() == [0, 1, 2, 0, 1, 2, 3, 4] and off.tolist() == [0, 3, 8]
    h = resample_poly_taps(160, 147)
    assert h.dtype == np.float32 and len(h) == 3201 and abs(h.sum() - 160) < 1e-2


def test_dict_mean():
    out = dict_mean([{"a": 1.0, "b": 2.0}, {"a": 3.0, "b": 6.0}])
    assert out == {"a": 2.0, "b": 4.0}


def test_helper_length_and_naming_helpers(tmp_path):
    """shift / pad / unify_length / cache_file_name of the mp3 path (ssr_eval/eval.py:272-300, 327-332): same results
    as the reference's code on the same inputs (values below were produced by the reference's own methods)."""
    h = SSR_Eval_Helper(BasicTestee(), 44100, 44100, test_data_root=str(tmp_path))
    x = np.arange(1, 7, dtype=np.float32)
    assert h.shift(x, 2).tolist() == [3, 4, 5, 6, 0, 0]
    assert h.shift(x, -2).tolist() == [0, 0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        h.shift(x, 0)  # ret[:-0] is empty: the reference raises here too
    a, b = h.pad(np.ones(3), np.ones(5))
    assert a.tolist() == [1, 1, 1, 0, 0] and b.tolist() == [1] * 5
    a, b = h.pad(np.ones(5), np.ones(2))
    assert a.tolist() == [1] * 5 and b.tolist() == [1, 1, 0, 0, 0]
    a, b = h.unify_length(np.arange(6.0), np.arange(4.0))
    assert a.tolist() == [0, 1, 2, 3] and len(b) == 4
    a, b = h.unify_length(np.arange(3.0), np.arange(5.0))
    assert a.tolist() == [0, 1, 2, 0, 0]
    assert h.cache_file_name("proc_mp3_32_44100", "/d/p360/p360_001.wav") == "/d/p360/p360_001_proc_mp3_32_44100.flac"
    assert h.cache_file_name("k", "a/b.wav", suffix=".mp3") == "a/b_k.mp3"
    with pytest.raises(NotImplementedError):
        h.mp3_encoding("f.wav", x, 44100)


def test_reference_methods_live_when_the_reference_is_mounted(tmp_path):
    """Build-container only: the same helper methods executed from /root/reference/ssr_eval/eval.py itself."""
    import importlib.util
    import sys
    import types
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "ssr_eval")):
        pytest.skip("reference not mounted (GPU box)")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(repo, "oracle", "shims"))
    saved = {k: sys.modules.get(k) for k in ("ssr_eval", "ssr_eval.utils", "ssr_eval.dsp", "ssr_eval.metrics",
                                             "ssr_eval.lowpass", "ssr_eval.eval")}
    try:
        pkg = types.ModuleType("ssr_eval")
        pkg.__path__ = [os.path.join(ref_root, "ssr_eval")]
        sys.modules["ssr_eval"] = pkg
        mods = {}
        for name in ("utils", "dsp", "metrics", "lowpass", "eval"):
            spec = importlib.util.spec_from_file_location("ssr_eval." + name, os.path.join(ref_root, "ssr_eval", name + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules["ssr_eval." + name] = m
            spec.loader.exec_module(m)
            mods[name] = m
        R = mods["eval"].SSR_Eval_Helper
        ours = SSR_Eval_Helper(BasicTestee(), 44100, 44100, test_data_root=str(tmp_path))
        rng = np.random.default_rng(0)
        x, y = rng.standard_normal(50).astype(np.float32), rng.standard_normal(37).astype(np.float32)
        for sh in (-7, -1, 1, 9):
            assert np.array_equal(R.shift(None, x, sh), ours.shift(x, sh))
        for p, q in ((x, y), (y, x), (x, x)):
            for f in ("pad", "unify_length"):
                ra, rb = getattr(R, f)(None, p, q)
                oa, ob = getattr(ours, f)(p, q)
                assert np.array_equal(ra, oa) and np.array_equal(rb, ob)
        assert R.cache_file_name(None, "key", "/a/b/c.wav") == ours.cache_file_name("key", "/a/b/c.wav")
        d1, d2 = {"cutoff_freq": [1000, 12000]}, {"cutoff_freq": [1000, 12000]}
        assert R._cutoff2sr(None, d1) == ours._cutoff2sr(d2) and d1 == d2
        (tmp_path / "p1").mkdir()
        for n in ("a.wav", "b.flac", "c_proc.wav", "x.DS_Store.wav", "d.mp3"):
            (tmp_path / "p1" / n).write_bytes(b"")
        assert sorted(R.get_test_file_list(None, str(tmp_path / "p1"))) == sorted(ours.get_test_file_list(str(tmp_path / "p1")))
        # dict_mean (utils.py:24-28) and the lowpass dispatcher's `limit`
        dm = mods["utils"].dict_mean
        rows = [{"a": float(i), "b": float(i * i)} for i in range(5)]
        assert dm(rows) == dict_mean(rows)
        for v in (-3, 2, 5.7, 10, 99):
            assert mods["lowpass"].limit(v, 10, 2) == limit(v, 10, 2)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_per_family_degradation_methods_keys_and_calls(tmp_path, monkeypatch):
    """lowpass_butterworth / _chebyshev / _ellip / _bessel / _subsampling / _stft_hard (ssr_eval/eval.py:334-421):
    key naming, the `low_rate == sr -> low_rate - 1` rule and the arguments handed to the filters, with the GPU
    entry points replaced by recorders (host logic only)."""
    import ssr_eval_b200.eval as ev
    calls = []

    def fake_lowpass_batch(waves, highcut, fs, order=5, _type="butter"):
        calls.append(("lowpass", len(waves), highcut, fs, order, _type))
        return [np.asarray(w) * 0.5 for w in waves]

    def fake_hard(waves, ratios, mode=None):
        calls.append(("stft_hard", len(waves), [round(r, 6) for r in ratios]))
        return [np.asarray(w) * 0.25 for w in waves]

    monkeypatch.setattr(ev, "lowpass_batch", fake_lowpass_batch)
    monkeypatch.setattr(ev, "stft_hard_lowpass_batch", fake_hard)
    h = SSR_Eval_Helper(BasicTestee(), 44100, 44100, test_data_root=str(tmp_path),
                        setting_fft={"cutoff_freq": [4000, 22050]}, setting_subsampling={"cutoff_freq": [8000]},
                        setting_lowpass_filtering={"filter": ["butter", "cheby", "ellip", "bessel"],
                                                   "cutoff_freq": [6000], "filter_order": [2, 8]})
    x = np.linspace(-1, 1, 1000).astype(np.float32)
    d = h.lowpass_butterworth("f.wav", x, 44100)
    assert list(d) == ["proc_bw_12000_2_44100", "proc_bw_12000_8_44100"]
    assert calls == [("lowpass", 1, 6000, 44100, 2, "butter"), ("lowpass", 1, 6000, 44100, 8, "butter")]
    calls.clear()
    assert list(h.lowpass_chebyshev("f.wav", x, 44100)) == ["proc_ch_12000_2_44100", "proc_ch_12000_8_44100"]
    assert [c[5] for c in calls] == ["cheby1", "cheby1"]
    assert list(h.lowpass_ellip("f.wav", x, 44100)) == ["proc_el_12000_2_44100", "proc_el_12000_8_44100"]
    assert list(h.lowpass_bessel("f.wav", x, 44100)) == ["proc_bessel_12000_2_44100", "proc_bessel_12000_8_44100"]
    calls.clear()
    assert list(h.lowpass_subsampling("f.wav", x, 44100)) == ["proc_subsampling_16000_44100"]
    assert calls == [("lowpass", 1, 8000, 44100, 1, "subsampling")]
    calls.clear()
    d = h.lowpass_stft_hard("f.wav", x, 44100)
    # doubled cutoff 44100 == sr -> 44099 (eval.py:404-405); ratio = (low_rate // 2) / int(sr / 2)
    assert list(d) == ["proc_fft_8000_44100", "proc_fft_44099_44100"]
    assert calls == [("stft_hard", 2, [round(4000 / 22050, 6), round(22049 / 22050, 6)])]
    # preprocess-style fan-out: every family present, reference order (filters, subsampling, fft)
    calls.clear()
    full = h._degrade_batch([x, x[:500]], 44100)
    assert [list(o) for o in full] == [list(full[0])] * 2
    assert list(full[0]) == ["proc_bw_12000_2_44100", "proc_bw_12000_8_44100", "proc_ch_12000_2_44100",
                             "proc_ch_12000_8_44100", "proc_el_12000_2_44100", "proc_el_12000_8_44100",
                             "proc_bessel_12000_2_44100", "proc_bessel_12000_8_44100",
                             "proc_subsampling_16000_44100", "proc_fft_8000_44100", "proc_fft_44099_44100"]
    assert all(c[1] == 2 for c in calls if c[0] == "lowpass") and calls[-1][1] == 4
