#!/usr/bin/env python
"""Locate spectrogram elements K1 (store mode) leaves unwritten: poison the workspace, run, scan."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssr_eval_b200 import _native as N  # noqa: E402
from ssr_eval_b200.engine import StftMetrics, pack_ragged, _np_ptr  # noqa: E402


def main():
    rng = np.random.default_rng(3)
    lens = [int(x) for x in rng.integers(3000, 90000, size=97)] + [2049, 1025, 240000]
    tgt = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    est = [(0.7 * t + 1e-2 * rng.standard_normal(len(t))).astype(np.float32) for t in tgt]
    eng = StftMetrics(2048, 512)
    F = 1025
    e_h, off = pack_ragged(est)
    t_h, _ = pack_ragged(tgt)
    e_d, t_d = e_h.cuda(), t_h.cuda()
    need = N.lib().ssr_stft_metrics_workspace_bytes(eng._plan, _np_ptr(off), len(lens), 15)
    eng._ws.get(need, e_d.device).fill_(0xFF)
    torch.cuda.synchronize()
    out = eng.metrics_device(e_d, t_d, off, 15)
    torch.cuda.synchronize()
    frames = np.array([eng.num_frames(l) for l in lens])
    total = int(frames.sum())
    reg = (4 * total * F + 255) // 256 * 256
    ws = eng._ws.buf[:need].cpu().numpy()
    starts = np.concatenate([[0], np.cumsum(frames)])
    for name, o in (("spec_e", need - 2 * reg), ("spec_t", need - reg)):
        a = ws[o:o + 4 * total * F].view(np.uint32).reshape(total, F)
        bad = np.argwhere(a == 0xFFFFFFFF)
        print(name, "unwritten elements:", len(bad))
        rows = sorted(set(int(b[0]) for b in bad))
        for r in rows[:40]:
            p = int(np.searchsorted(starts, r, side="right") - 1)
            cols = bad[bad[:, 0] == r][:, 1]
            print("  pair", p, "len", lens[p], "T", int(frames[p]), "frame", r - int(starts[p]), "n_cols", len(cols),
                  "cols", cols[:6].tolist(), "..", cols[-3:].tolist())
    print("nan ssim pairs:", [(i, lens[i], int(frames[i])) for i in np.argwhere(np.isnan(out[:, 3].cpu().numpy())).ravel()])


if __name__ == "__main__":
    main()
