#!/usr/bin/env python
"""Tiny invocation of every kernel family, meant to run under compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  --error-exitcode 9 python tools/sanitize_small.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py quick
    compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_small.py quick

Sizes are small (ragged utterances of 0.1 .. 0.5 s, edge frames included) so that racecheck's
~100x slowdown stays within a minute.  Prints one line per family; results are not checked here
(tests/test_gpu_parity.py does that) -- the sanitizer's exit code is the verdict."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ssr_eval_b200 import _native as N  # noqa: E402
from ssr_eval_b200.engine import (StftMetrics, PolyphaseResampler, HardLowpass, HardLowpassDense, SpliceIstft,  # noqa: E402
                                  sosfiltfilt_batch, xcorr_argmax_batch, pcm16_to_float_device)


def main():
    quick = "quick" in sys.argv[1:]
    rng = np.random.default_rng(7)
    lens = [4800, 9613, 23000] if quick else [4800, 9613, 23000, 2049, 31111]
    tgt = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    est = [(t + 1e-3 * rng.standard_normal(len(t))).astype(np.float32) for t in tgt]
    sizes = [(2048, 512), (2229, 480)] if quick else [(2048, 512), (2229, 480), (1114, 240), (743, 160), (1024, 256),
                                                      (1031, 300)]
    for n_fft, hop in sizes:
        eng = StftMetrics(n_fft, hop)
        for flags in ((N.METRIC_ALL,) if quick else (N.METRIC_LSD, 7, N.METRIC_ALL)):
            r = eng.metrics(est, tgt, flags)
        print("K1/K2 n_fft %d hop %d ok" % (n_fft, hop), r[0], flush=True)
    for up, down in ([(160, 147), (147, 160)] if quick else [(160, 147), (147, 160), (441, 160), (147, 80), (80, 147), (3, 1)]):
        y = PolyphaseResampler(up, down).resample(tgt)
        print("K3 %d/%d ok" % (up, down), len(y[0]), flush=True)
    lp = HardLowpass(2048, 441)
    y = lp.apply(tgt, [0.25, 0.5, 0.9, 0.1, 1.0][:len(tgt)])
    print("K4 2048 ok", float(np.abs(y[0]).max()), flush=True)
    if not quick:
        y = HardLowpass(1024, 256).apply(tgt, [0.5] * len(tgt))
        print("K4 generic ok", float(np.abs(y[0]).max()), flush=True)
    sp = SpliceIstft(2048, 512)
    cuts = sp.cutoff_indices(est)
    y = sp.apply(est, tgt, cuts)
    print("K6 ok", cuts, flush=True)
    from scipy.signal import butter
    sos = butter(8, 0.3, output="sos")
    y = sosfiltfilt_batch(sos, tgt)
    print("K7 ok", float(np.abs(y[0]).max()), flush=True)
    # round-2 kernels: K0 (PCM16), K3 with TMA-staged interior tiles (long utterance), K4d (dense), K8 (alignment)
    pcm = torch.from_numpy(rng.integers(-32768, 32768, size=10001, dtype=np.int16)).cuda()
    print("K0 ok", float(pcm16_to_float_device(pcm[1:]).abs().max()), flush=True)
    long = (0.1 * rng.standard_normal(40000)).astype(np.float32)
    y = PolyphaseResampler(160, 147).resample([long, tgt[0]])
    print("K3 bulk ok", len(y[0]), flush=True)
    y = PolyphaseResampler(44100, 48000, bank="resampy_kaiser_best").resample([tgt[0]])
    print("K3 resampy bank ok", len(y[0]), flush=True)
    y = HardLowpassDense(2048, 441).apply(tgt[:2], [0.25, 0.5])
    print("K4d ok", float(np.abs(y[0]).max()), flush=True)
    k = xcorr_argmax_batch([np.roll(t, 7) for t in tgt], tgt)
    print("K8 ok", k, flush=True)
    torch.cuda.synchronize()
    print("sanitize_small done; kernel launches:", N.launch_count(), flush=True)


if __name__ == "__main__":
    main()
