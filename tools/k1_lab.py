#!/usr/bin/env python
"""K1 laboratory: A/B timing of kernel variants and ablations of the 2048-point STFT+LSD kernel on the
BASELINE configs[1] workload (1024 pairs x 240000 samples, n_fft 2048 / hop 512).  One JSON line per
variant.  Variants are selected through SSR_K1_VARIANT (read by the library on every launch):
  0        production kernel
  1        phase-staggered kernel (k1_2048s.cuh)
  100+ABL  ablations of the production kernel (timing only, results wrong): 1 no epilogue math,
           2 no global loads / conversions, 4 no CTA barriers, 8 no shared-memory exchanges
SSR_K1_EXTRA_SMEM pads the dynamic shared memory of the 100+ variants (CTAs per SM sweep)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssr_eval_b200 import _native as N  # noqa: E402
from ssr_eval_b200.engine import StftMetrics, offsets_of  # noqa: E402

dev = torch.device("cuda", 0)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def setvar(v, extra=0):
    os.environ["SSR_K1_VARIANT"] = str(v)
    os.environ["SSR_K1_EXTRA_SMEM"] = str(extra)


def main():
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    eng = StftMetrics(2048, 512)
    # ---- correctness of the non-ablation variants on a ragged batch (edge frames, all flag sets)
    rng = np.random.default_rng(3)
    lens = [int(x) for x in rng.integers(3000, 90000, size=97)] + [2049, 1025, 240000]
    tgt = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    est = [(0.7 * t + 1e-2 * rng.standard_normal(len(t))).astype(np.float32) for t in tgt]
    for flags in (1, 7, 15):
        setvar(0)
        ref = eng.metrics(est, tgt, flags)
        for v in ((1, 2, 3, 4, 5, 6, 7, 8, 9) if flags == 1 else (1, 4, 5, 8, 9)):
            setvar(v)
            got = eng.metrics(est, tgt, flags)
            same = np.array_equal(np.nan_to_num(ref, nan=-1.0), np.nan_to_num(got, nan=-1.0))
            diff = float(np.nanmax(np.abs(ref - got)))
            print(json.dumps({"check": "variant %d vs 0" % v, "flags": flags, "bit_identical": bool(same),
                              "max_abs_diff": diff}), flush=True)
    # ---- timing on the bench workload
    n, L = 1024, 240000
    tg = 0.1 * torch.randn(n * L, generator=g, device=dev)
    es = tg + 1e-3 * torch.randn(n * L, generator=g, device=dev)
    off = offsets_of([L] * n)
    off_d = torch.from_numpy(off).to(dev)
    out = torch.empty((n, 4), dtype=torch.float64, device=dev)
    cases = [("production (3 CTAs/SM)", 0, 0, 1), ("staggered 3x128", 1, 0, 1),
             ("production, flags 7", 0, 0, 7), ("staggered, flags 7", 1, 0, 7),
             ("production, flags 15 (K1+K2)", 0, 0, 15), ("staggered, flags 15 (K1+K2)", 1, 0, 15),
             ("two-level twiddles, 3 CTAs/SM", 2, 0, 1), ("two-level twiddles, 4 CTAs/SM", 3, 0, 1),
             ("TMEM twiddles, 3 CTAs/SM", 4, 0, 1), ("TMEM twiddles, 4 CTAs/SM", 5, 0, 1),
             ("TMEM twiddles + sample ring, 3 CTAs/SM", 6, 0, 1), ("TMEM twiddles + sample ring, 4 CTAs/SM", 7, 0, 1),
             ("TMEM ring + early loads, 3 CTAs/SM", 8, 0, 1), ("TMEM ring + early loads, 4 CTAs/SM", 9, 0, 1),
             ("ABL16 no input F2F", 116, 0, 1), ("ABL32 constant window", 132, 0, 1),
             ("ABL48 no input F2F, constant window", 148, 0, 1),
             ("1 CTA/SM: ABL1 no epilogue math", 101, 72 * 1024, 1), ("1 CTA/SM: ABL2 no loads/conversions", 102, 72 * 1024, 1),
             ("1 CTA/SM: ABL4 no barriers", 104, 72 * 1024, 1), ("1 CTA/SM: ABL8 no smem exchange", 108, 72 * 1024, 1),
             ("1 CTA/SM: ABL15 FP64 butterflies only", 115, 72 * 1024, 1),
             ("ABL0 3 CTAs/SM", 100, 0, 1), ("ABL0 2 CTAs/SM", 100, 34 * 1024, 1), ("ABL0 1 CTA/SM", 100, 72 * 1024, 1),
             ("ABL1 no epilogue math", 101, 0, 1), ("ABL2 no loads/conversions", 102, 0, 1),
             ("ABL3 no epilogue, no loads", 103, 0, 1), ("ABL4 no barriers", 104, 0, 1),
             ("ABL8 no smem exchange", 108, 0, 1), ("ABL12 no barriers, no exchange", 112, 0, 1),
             ("ABL15 FP64 butterflies only", 115, 0, 1)]
    if "modes" in sys.argv[1:]:
        cases = []
        for flags in (1, 7, 15):
            for name, v in (("production", 0), ("TMEM tw, 3/SM", 4), ("TMEM tw, 4/SM", 5), ("TMEM tw+ring+early, 3/SM", 8),
                            ("TMEM tw+ring+early, 4/SM", 9)):
                cases.append(("flags %d: %s" % (flags, name), v, 0, flags))
        sys.argv = [a for a in sys.argv if a != "modes"]
    only = set(sys.argv[1:])
    for name, v, extra, flags in cases:
        if only and str(v) not in only:
            continue
        setvar(v, extra)
        try:
            ms = timeit(lambda: eng.metrics_device(es, tg, off, flags, offsets_dev=off_d, out=out))
            print(json.dumps({"case": name, "variant": v, "extra_smem": extra, "flags": flags, "ms": round(ms, 4),
                              "pairs_per_s": round(n / (ms * 1e-3), 1)}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"case": name, "variant": v, "error": str(e)}), flush=True)
    # ---- n_fft 2048 / hop 441 (evaluation at 44.1 kHz): no sample ring
    del tg, es
    n, L = 1024, 220500
    eng2 = StftMetrics(2048, 441)
    tg = 0.1 * torch.randn(n * L, generator=g, device=dev)
    es = tg + 1e-3 * torch.randn(n * L, generator=g, device=dev)
    off = offsets_of([L] * n)
    off_d = torch.from_numpy(off).to(dev)
    for flags in (1, 15):
        ref = None
        for name, v in (("production", 0), ("TMEM tw, 3/SM", 4), ("TMEM tw, 4/SM", 5)):
            setvar(v)
            ms = timeit(lambda: eng2.metrics_device(es, tg, off, flags, offsets_dev=off_d, out=out))
            r = out.cpu().numpy().copy()
            ref = r if ref is None else ref
            print(json.dumps({"case": "hop 441 flags %d: %s" % (flags, name), "ms": round(ms, 4),
                              "pairs_per_s": round(n / (ms * 1e-3), 1),
                              "bit_identical_to_production": bool(np.array_equal(np.nan_to_num(r), np.nan_to_num(ref)))}), flush=True)
    setvar(0)


if __name__ == "__main__":
    main()
