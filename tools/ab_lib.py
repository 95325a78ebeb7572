#!/usr/bin/env python
"""A/B timing of several builds of the library in ONE process (dev tool):

    python tools/ab_lib.py [--nfft 2048 --hop 512 --flags 1,7,15] lib_a.so lib_b.so ...   ("-" = the in-tree build)

The same device-resident batch (1024 pairs x 240000 samples) is scored by every build; prints one JSON line per
(build, flag set) with the time per launch sequence and the mean of each metric (must agree between builds)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssr_eval_b200 import _native as N  # noqa: E402
from ssr_eval_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nfft", type=int, default=2048)
ap.add_argument("--hop", type=int, default=512)
ap.add_argument("--flags", default="1,7,15")
ap.add_argument("--pairs", type=int, default=1024)
ap.add_argument("--k4", action="store_true", help="also time K4 (stft_hard 2048/441, 1024 x 220500) and K3 (160/147)")
ap.add_argument("libs", nargs="+")
a = ap.parse_args()
default_path = N._LIB_PATH
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
n, L = a.pairs, 240000
tg = 0.1 * torch.randn(n * L, generator=g, device=dev)
es = tg + 1e-3 * torch.randn(n * L, generator=g, device=dev)
off = engine.offsets_of([L] * n)
off_d = torch.from_numpy(off).to(dev)
out = torch.empty((n, 4), dtype=torch.float64, device=dev)
for lib in a.libs:
    N._lib = None
    N._LIB_PATH = default_path if lib == "-" else os.path.abspath(lib)
    eng = engine.StftMetrics(a.nfft, a.hop)
    for flags in [int(f) for f in a.flags.split(",")]:
        for _ in range(2):
            eng.metrics_device(es, tg, off, flags, offsets_dev=off_d, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.metrics_device(es, tg, off, flags, offsets_dev=off_d, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"lib": os.path.basename(N._LIB_PATH), "n_fft": a.nfft, "hop": a.hop, "flags": flags,
                          "ms": round(ms, 4), "pairs_per_s": round(n / ms * 1e3, 1),
                          "mean": [float(x) for x in out.nanmean(dim=0).cpu()]}), flush=True)
    del eng
    if a.k4:
        L4 = 220500
        x4 = tg[:n * L4]
        off4 = engine.offsets_of([L4] * n)
        off4_d = torch.from_numpy(off4).to(dev)
        lp = engine.HardLowpass(2048, 441)
        cuts = [lp.cut_bin(12000 / 22050)] * n
        rs = engine.PolyphaseResampler(160, 147)
        for name, fn in (("k4", lambda: lp.apply_device(x4, off4, cuts, off4_d)),
                         ("k3_160_147", lambda: rs.resample_device(x4, off4, off4_d)[0])):
            for _ in range(2):
                y4 = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y4 = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(json.dumps({"lib": os.path.basename(N._LIB_PATH), "kernel": name, "ms": round(ms, 4),
                              "utt_per_s": round(n / ms * 1e3, 1), "checksum": float(y4.double().abs().sum().cpu()),
                              "sum": float(y4.double().sum().cpu())}), flush=True)
        del lp, rs, y4
