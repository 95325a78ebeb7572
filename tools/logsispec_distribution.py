#!/usr/bin/env python
"""How far do the reference's own float32 reductions move log_sispec / sispec / lsd at full size?

64 BASELINE-size pairs (L = 240000, n_fft 2048 / hop 512; hard-low-passed, noisy and scaled estimates) are scored
  (a) by oracle.evaluation -- the reference arithmetic: float32 torch.sum / torch.norm / torch.mean over
      T*F = 4.8e5 elements -- at 1 torch thread and at all host threads (torch chunks its reductions per thread),
  (b) by oracle.evaluation_exact_reductions -- the same float32 element-wise formulas, float64 reductions,
  (c) on a GPU box, by the CUDA path (K1, float64 accumulators).
Prints a markdown table of |a - b|, |a(1 thread) - a(N threads)| and, with a GPU, |c - a| and |c - b|: the
tolerance the full-size parity tests use for log_sispec is the measured bound of |c - a|, not a guess.

    python tools/logsispec_distribution.py [--pairs 64] [--out gpurun_out/logsispec_distribution.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (tools/ is test infrastructure)
from ssr_eval_b200.synth import speech_like  # noqa: E402

L, N_FFT, HOP = 240000, 2048, 512
KEYS = ("lsd", "log_sispec", "sispec")


def make_pairs(n):
    est, tgt, kind = [], [], []
    rng = np.random.default_rng(2026)
    for i in range(n):
        t = speech_like(L, 48000, seed=4000 + i)
        k = i % 4
        if k == 0:
            e = oracle.lowpass(t, (4000, 8000, 12000, 16000)[(i // 4) % 4], 48000, order=1, _type="stft_hard").astype(np.float32)
        elif k == 1:
            e = (t + 10.0 ** rng.uniform(-4, -1.5) * rng.standard_normal(L)).astype(np.float32)
        elif k == 2:
            e = oracle.lowpass(t, (4000, 8000, 12000)[(i // 4) % 3], 48000, order=1, _type="subsampling").astype(np.float32)[:L]
        else:
            e = (float(rng.uniform(0.2, 1.5)) * t + 1e-3 * rng.standard_normal(L)).astype(np.float32)
        est.append(e)
        tgt.append(t)
        kind.append(("stft_hard", "noise", "subsampling", "scaled+noise")[k])
    return est, tgt, kind


def stats(d):
    d = np.abs(np.asarray(d, dtype=np.float64))
    return {"max": float(d.max()), "p95": float(np.percentile(d, 95)), "median": float(np.median(d)), "mean": float(d.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "logsispec_distribution.json"))
    args = ap.parse_args()
    est, tgt, kind = make_pairs(args.pairs)
    n_threads = len(os.sched_getaffinity(0))
    rows = {}
    for name, th in (("ref_1thread", 1), ("ref_Nthreads", n_threads)):
        torch.set_num_threads(th)
        rows[name] = [oracle.evaluation(e, t, n_fft=N_FFT, hop=HOP, which=KEYS) for e, t in zip(est, tgt)]
    rows["exact"] = [oracle.evaluation_exact_reductions(e, t, N_FFT, HOP) for e, t in zip(est, tgt)]
    if torch.cuda.is_available():
        from ssr_eval_b200.engine import StftMetrics
        from ssr_eval_b200 import _native as N
        got = StftMetrics(N_FFT, HOP).metrics(est, tgt, N.METRIC_LSD | N.METRIC_LOG_SISPEC | N.METRIC_SISPEC)
        rows["gpu"] = [dict(zip(KEYS, g[:3])) for g in got]
    pairs = [("ref_1thread", "exact"), ("ref_Nthreads", "exact"), ("ref_1thread", "ref_Nthreads")]
    if "gpu" in rows:
        pairs += [("gpu", "ref_1thread"), ("gpu", "ref_Nthreads"), ("gpu", "exact")]
    summary = {}
    print("| |a - b| over %d full-size pairs | metric | max | p95 | median |" % args.pairs)
    print("|---|---|---|---|---|")
    for a, b in pairs:
        for k in KEYS:
            s = stats([ra[k] - rb[k] for ra, rb in zip(rows[a], rows[b])])
            summary["%s-%s/%s" % (a, b, k)] = s
            print("| %s vs %s | %s | %.2e | %.2e | %.2e |" % (a, b, k, s["max"], s["p95"], s["median"]))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"pairs": args.pairs, "length": L, "n_fft": N_FFT, "hop": HOP, "host_threads": n_threads, "kinds": kind,
               "values": {k: [{m: float(v[m]) for m in KEYS} for v in r] for k, r in rows.items()},
               "summary": summary}, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
