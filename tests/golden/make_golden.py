"""Generate tests/golden/golden_v1.npz by RUNNING THE REFERENCE'S OWN CODE.

Runs only in the build container (needs /root/reference, read-only).  The reference package
cannot be imported as-is (librosa / scikit-image / torchlibrosa / soundfile are absent and not
installable), so the missing third-party modules are replaced by the shims in
``oracle/shims`` (built from the oracle's restatement); everything in-repo -- AudioMetrics.evaluation,
lsd, sispec, to_log, energy_unify, lowpass dispatch, stft_hard_lowpass_v0, subsampling,
FDomainHelper, dict_mean -- is the reference's own source, executed unmodified.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))

from ssr_eval_b200.synth import speech_like  # noqa: E402


def _load_reference():
    """Import ssr_eval.{utils,dsp,metrics,lowpass} from /root/reference without executing
    ssr_eval/__init__.py (which pulls eval.py -> file I/O stack)."""
    pkg = types.ModuleType("ssr_eval")
    pkg.__path__ = [os.path.join(REF, "ssr_eval")]
    sys.modules["ssr_eval"] = pkg
    mods = {}
    for name in ("utils", "dsp", "metrics", "lowpass", "eval"):
        spec = importlib.util.spec_from_file_location(
            "ssr_eval." + name, os.path.join(REF, "ssr_eval", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["ssr_eval." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def main():
    from scipy.signal import resample_poly
    ref = _load_reference()
    AudioMetrics = ref["metrics"].AudioMetrics
    lowpass = ref["lowpass"].lowpass
    out = {}
    meta = []

    def metric_case(name, rate, est, tgt):
        res = AudioMetrics(rate).evaluation(est, tgt, "none")
        out[f"{name}/est"] = est.astype(np.float32)
        out[f"{name}/tgt"] = tgt.astype(np.float32)
        out[f"{name}/rate"] = np.int64(rate)
        out[f"{name}/metrics"] = np.array(
            [res["lsd"], res["log_sispec"], res["sispec"], res["ssim"]], dtype=np.float64)
        meta.append(name)
        print(name, rate, len(est), len(tgt), res)

    # A: the reference's own test() flow (ssr_eval/test.py:21-38) on one synthetic utterance:
    #    48k target -> 44.1k model input -> stft_hard cutoff 12 kHz -> identity testee ->
    #    polyphase 44.1k -> 48k (eval.py:144-150 via the librosa shim) -> metrics @ 48k (n_fft 2229).
    import librosa  # the shim
    tgt48 = speech_like(24000, sr=48000, seed=1)
    x44 = resample_poly(tgt48, 147, 160).astype(np.float32)
    lp44 = lowpass(x44, 24000 // 2, 44100, order=1, _type="stft_hard")
    est48 = librosa.resample(lp44, 44100, 48000, res_type="polyphase")
    out["A/x44"] = x44
    out["A/lp44"] = lp44
    metric_case("A", 48000, est48, tgt48)

    # B: evaluation at 44.1k (n_fft 2048 / hop 441), subsampling degradation, cutoff 8 kHz
    tgt44 = speech_like(22050, sr=44100, seed=2)
    est44 = lowpass(tgt44, 16000 // 2, 44100, order=1, _type="subsampling").astype(np.float32)
    metric_case("B", 44100, est44, tgt44)

    # C: 16 kHz (n_fft 743 -- prime), benign additive-noise estimate
    tgt16 = speech_like(12000, sr=16000, seed=3)
    est16 = (tgt16 + 1e-3 * np.random.default_rng(33).standard_normal(12000)).astype(np.float32)
    metric_case("C", 16000, est16, tgt16)

    # D: length mismatch < 100 -> truncation to the shorter (metrics.py:82-90), stft_hard 4 kHz @ 48k
    tgt48b = speech_like(24000, sr=48000, seed=4)
    est48b = lowpass(tgt48b, 4000, 48000, order=1, _type="stft_hard")[:-57]
    metric_case("D", 48000, est48b, tgt48b)

    # E: 24 kHz (n_fft 1114 = 2*557), stft_hard cutoff 6 kHz
    tgt24 = speech_like(12000, sr=24000, seed=5)
    est24 = lowpass(tgt24, 6000, 24000, order=1, _type="stft_hard")
    metric_case("E", 24000, est24, tgt24)

    out["metric_cases"] = np.array(meta)

    # lowpass goldens (stft_hard / subsampling / IIR passthrough) on one 44.1k utterance
    x = speech_like(22050, sr=44100, seed=6)
    out["LP/x"] = x
    lp_meta = []
    for fs in (44100, 48000):
        for cutoff in (4000, 8000, 12000, 16000):
            y = lowpass(x, cutoff, fs, order=1, _type="stft_hard")
            out[f"LP/stft_hard_{cutoff}_{fs}"] = np.asarray(y, dtype=np.float32)
            lp_meta.append(f"stft_hard_{cutoff}_{fs}")
    for cutoff in (4000, 12000):
        y = lowpass(x, cutoff, 44100, order=1, _type="subsampling")
        out[f"LP/subsampling_{cutoff}_44100"] = np.asarray(y, dtype=np.float32)
        lp_meta.append(f"subsampling_{cutoff}_44100")
    y = lowpass(x, 8000, 44100, order=8, _type="butter")
    out["LP/butter_8000_44100"] = np.asarray(y, dtype=np.float64)
    lp_meta.append("butter_8000_44100")
    out["lp_cases"] = np.array(lp_meta)

    # aggregation: reference dict_mean (utils.py:24-28) twice = mean of speaker means (eval.py:200-216)
    rng = np.random.default_rng(7)
    per_spk = [rng.random((n, 4)) for n in (3, 5, 2)]
    dm = ref["utils"].dict_mean
    keys = ["lsd", "log_sispec", "sispec", "ssim"]
    spk_means = [dm([dict(zip(keys, row)) for row in a]) for a in per_spk]
    avg = dm(spk_means)
    # BasicTestee.postprocessing (eval.py:33-41) run by the reference's own class (librosa stft/istft shims)
    BasicTestee = ref["eval"].BasicTestee
    xin = lowpass(speech_like(22050, sr=44100, seed=8), 4000, 44100, order=1, _type="stft_hard")
    model_out = (xin + 0.02 * np.random.default_rng(9).standard_normal(22050)).astype(np.float32)
    bt = BasicTestee()
    out["PP/x"] = xin.astype(np.float32)
    out["PP/out"] = model_out
    out["PP/cutoff"] = np.int64(bt._get_cutoff_index(xin))
    out["PP/renewed"] = np.asarray(bt.postprocessing(xin, model_out), dtype=np.float32)
    print("postprocessing cutoff index", int(out["PP/cutoff"]))

    out["AGG/values"] = np.concatenate(per_spk)
    out["AGG/counts"] = np.array([len(a) for a in per_spk])
    out["AGG/averaged"] = np.array([avg[k] for k in keys])

    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
