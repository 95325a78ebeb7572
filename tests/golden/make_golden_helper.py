"""Generate tests/golden/helper_reference_runs.json by RUNNING THE REFERENCE'S OWN ORCHESTRATOR.

`SSR_Eval_Helper.evaluate()` of /root/reference/ssr_eval/eval.py is executed unmodified on a small synthetic
"VCTK" tree.  What is not available in this container is stood in for as narrowly as possible:
  * librosa / skimage / torchlibrosa / soundfile  -> oracle/shims (restatement; see oracle/shims/README.md);
  * librosa.load / soundfile.write                -> scipy.io.wavfile (float32 WAV, native rate only -- the
    data set is written at the rate the run loads it at, so NO resampler that the reference leaves to
    librosa / sox is involved);
  * os.system("sox file -r SR temp.wav")          -> a file copy when SR is the file's rate; for the run that is
    ssr_eval/test.py:24-36 verbatim (files at 44.1 kHz, evaluation_sr = 48000) the target is produced by
    scipy.signal.resample_poly(x, 160, 147) in float32 -- a declared stand-in for the sox binary; what the run
    pins is everything AFTER the target exists (K4 -> polyphase 44.1k->48k -> n_fft 2229 / hop 480 metrics).
The orchestration itself -- distortion fan-out and key naming (eval.py:334-421), the plugin call and its
(wav, extra_metrics) form, the polyphase resampling of the output (eval.py:144-150), AudioMetrics.evaluation,
per-speaker means and the mean of means (eval.py:200-216), the result schema -- is the reference's code.

    python tests/golden/make_golden_helper.py      (build container only: needs /root/reference)
"""
import importlib.util
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))

from ssr_eval_b200.synth import speech_like  # noqa: E402

# (speaker, [(file name, samples, seed)]) -- shared with tests/test_gpu_parity.py through the JSON
DATASET = [("p360", [("u0.wav", 22050, 11), ("u1.wav", 26460, 12)]),
           ("s5", [("a.wav", 19845, 21), ("b.wav", 24255, 22), ("c.wav", 30870, 23)])]
RATE = 44100


def write_dataset(root):
    from scipy.io import wavfile
    for spk, files in DATASET:
        os.makedirs(os.path.join(root, spk), exist_ok=True)
        for name, n, seed in files:
            wavfile.write(os.path.join(root, spk, name), RATE, speech_like(n, RATE, seed=seed))


def load_reference():
    import librosa
    import soundfile
    from scipy.io import wavfile

    def load(path, sr=None, **kw):
        native, data = wavfile.read(path)
        assert data.dtype == np.float32 and data.ndim == 1
        assert sr is None or int(sr) == native, "the golden run never resamples at load time"
        return data.copy(), native

    def write(path, data, samplerate, **kw):
        wavfile.write(path, int(samplerate), np.asarray(data, dtype=np.float32))

    librosa.load = load
    soundfile.write = write
    pkg = types.ModuleType("ssr_eval")
    pkg.__path__ = [os.path.join(REF, "ssr_eval")]
    sys.modules["ssr_eval"] = pkg
    mods = {}
    for name in ("utils", "dsp", "metrics", "lowpass", "eval"):
        spec = importlib.util.spec_from_file_location("ssr_eval." + name, os.path.join(REF, "ssr_eval", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["ssr_eval." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m

    def fake_system(cmd):
        parts = cmd.split()
        if parts[0] == "sox" and "-r" in parts:  # "sox <file> -r <sr> temp.wav"
            src, sr, dst = parts[1], int(parts[3]), parts[4]
            native, data = wavfile.read(src)
            if native == sr:
                shutil.copyfile(src, dst)
            else:  # stand-in for sox's rate conversion (see the module docstring)
                from math import gcd
                from scipy.signal import resample_poly as rp
                g = gcd(sr, native)
                wavfile.write(dst, sr, rp(data.astype(np.float32), sr // g, native // g).astype(np.float32))
        return 0

    mods["eval"].os.system = fake_system
    return mods


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    return o


def main():
    from scipy.signal import resample_poly
    ref = load_reference()
    Helper, Testee = ref["eval"].SSR_Eval_Helper, ref["eval"].BasicTestee

    class Upsampler(Testee):  # a "model" with a 48 kHz output; also exercises the (wav, extra metrics) return form
        def infer(self, x):
            return resample_poly(x, 160, 147).astype(np.float32), {"n_in": float(len(x))}

    runs = {}
    cwd = os.getcwd()
    for name, testee, kwargs in (
        ("identity_all_settings", Testee(), dict(
            input_sr=RATE, output_sr=RATE, evaluation_sr=RATE,
            setting_fft={"cutoff_freq": [4000, 12000]}, setting_subsampling={"cutoff_freq": [8000]},
            setting_lowpass_filtering={"filter": ["butter", "cheby"], "cutoff_freq": [6000], "filter_order": [4]})),
        ("upsampling_testee_output_48k", Upsampler(), dict(
            input_sr=RATE, output_sr=48000, evaluation_sr=RATE, setting_fft={"cutoff_freq": [8000]})),
        # ssr_eval/test.py:24-36 verbatim: identity testee, 44.1 kHz in / out, scored at 48 kHz (n_fft 2229, hop 480)
        ("reference_test_py", Testee(), dict(
            input_sr=44100, output_sr=44100, evaluation_sr=48000, setting_fft={"cutoff_freq": [12000]},
            save_processed_result=True)),
    ):
        tmp = tempfile.mkdtemp()
        try:
            root = os.path.join(tmp, "vctk")
            write_dataset(root)
            os.chdir(tmp)
            h = Helper(testee, test_name=name, test_data_root=root, **kwargs)
            res = h.evaluate(limit_test_nums=-1, limit_test_speaker=-1)
            saved = [f for f in os.listdir(os.path.join(tmp, "results")) if f.endswith(name + ".json")]
            assert len(saved) == 1
            on_disk = json.load(open(os.path.join(tmp, "results", saved[0])))
            assert list(on_disk) == list(res)
            runs[name] = {"kwargs": jsonable(kwargs), "result": jsonable(res)}
        finally:
            os.chdir(cwd)
            shutil.rmtree(tmp, ignore_errors=True)
        print(name, json.dumps(runs[name]["result"]["averaged"], indent=1)[:600])
    out = {"rate": RATE, "dataset": DATASET, "runs": runs}
    path = os.path.join(ROOT, "tests", "golden", "helper_reference_runs.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
