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
