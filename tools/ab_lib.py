#!/usr/bin/env python
"""A/B timing of two builds of the library (dev tool): python tools/ab_lib.py <path to .so> [n_fft hop]."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssr_eval_b200 import _native as N  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] != "-":
    N._LIB_PATH = os.path.abspath(sys.argv[1])
from ssr_eval_b200.engine import StftMetrics, offsets_of  # noqa: E402

n_fft, hop = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2048, 512)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
n, L = 1024, 240000
tg = 0.1 * torch.randn(n * L, generator=g, device=dev)
es = tg + 1e-3 * torch.randn(n * L, generator=g, device=dev)
off = offsets_of([L] * n)
off_d = torch.from_numpy(off).to(dev)
out = torch.empty((n, 4), dtype=torch.float64, device=dev)
eng = StftMetrics(n_fft, hop)
for flags in ((1,) if os.environ.get("AB_LSD_ONLY") else (1, 7, 15)):
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
    print(json.dumps({"lib": os.path.basename(N._LIB_PATH), "n_fft": n_fft, "hop": hop, "flags": flags, "ms": round(ms, 4),
                      "pairs_per_s": round(n / ms * 1e3, 1), "mean": [float(x) for x in out.nanmean(dim=0).cpu()]}), flush=True)
