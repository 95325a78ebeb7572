"""Generate tests/golden/dense_lowpass_v1.npz: stft_hard_lowpass_v0 of three synthetic utterances computed by the oracle
(oracle/lowpass.py: torch's CPU conv1d with torchlibrosa's kernels = what ssr_eval/lowpass.py:17-28 runs) IN THE BUILD
CONTAINER.  The bins above the cutoff of such a waveform are the rounding noise of the convolutions, so the result
depends on the accumulation order torch / oneDNN pick for the host CPU; the dense GPU mode (K4d) reproduces the order of
an AVX-512 host (this container: bit-identical STFT / ISTFT frames, profiles/r02_dense_dft_study.md).  A GPU box with
another CPU would give the live oracle another noise floor, so the bit-level test uses this fixture.

    python tests/golden/make_golden_dense.py        (build container)
"""
import os
import platform
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from ssr_eval_b200.synth import speech_like  # noqa: E402

CASES = [(30000, 0.3, 181), (22050, 12000 / 22050, 185), (1025, 0.5, 180)]  # (samples, lowpass_ratio, seed) at 44.1 kHz


def main():
    out = {}
    for i, (n, ratio, seed) in enumerate(CASES):
        x = speech_like(n, 44100, seed=seed)
        out["y%d" % i] = oracle.stft_hard_lowpass_v0(x, ratio)
    flags = ""
    try:
        flags = " ".join(sorted(set(f for f in open("/proc/cpuinfo").read().split() if f.startswith("avx"))))
    except Exception:
        pass
    out["cases"] = np.array(CASES, dtype=np.float64)
    out["host"] = np.array("torch %s; %s; %s" % (torch.__version__, platform.processor() or platform.machine(), flags))
    path = os.path.join(ROOT, "tests", "golden", "dense_lowpass_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", out["host"])


if __name__ == "__main__":
    main()
