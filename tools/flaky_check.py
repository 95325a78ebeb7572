#!/usr/bin/env python
"""Repeatability of the all-four-metrics path on a ragged batch with the workspace poisoned between runs
(catches reads of uninitialised workspace memory and races).  Prints per-column max differences."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssr_eval_b200.engine import StftMetrics  # noqa: E402


def main():
    rng = np.random.default_rng(3)
    lens = [int(x) for x in rng.integers(3000, 90000, size=97)] + [2049, 1025, 240000]
    tgt = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    est = [(0.7 * t + 1e-2 * rng.standard_normal(len(t))).astype(np.float32) for t in tgt]
    for n_fft, hop in ((2048, 512), (2229, 480)):
        eng = StftMetrics(n_fft, hop)
        runs = []
        for i, (variant, poison) in enumerate(((0, None), (0, 0xFF), (0, 0x7F), (1, 0x00), (1, 0xFF), (0, 0x3C))):
            os.environ["SSR_K1_VARIANT"] = str(variant)
            if poison is not None and eng._ws.buf is not None:
                eng._ws.buf.fill_(poison)
                torch.cuda.synchronize()
            runs.append((variant, poison, eng.metrics(est, tgt, 15)))
        ref = runs[0][2]
        for variant, poison, r in runs[1:]:
            d = np.abs(np.nan_to_num(r, nan=-1.0) - np.nan_to_num(ref, nan=-1.0))
            worst = int(np.argmax(d.max(axis=1)))
            print(json.dumps({"n_fft": n_fft, "variant": variant, "poison": poison, "max_diff_per_col": d.max(axis=0).tolist(),
                              "worst_pair": worst, "worst_len": lens[worst], "ref_row": ref[worst].tolist(),
                              "got_row": r[worst].tolist()}), flush=True)
    os.environ["SSR_K1_VARIANT"] = "0"


if __name__ == "__main__":
    main()
