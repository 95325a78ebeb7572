#!/usr/bin/env python
"""Device-resident runs of BASELINE.json configs 3, 4 and 5 (synthetic data), one JSON line each.

  cfg3: 4096 pairs x cutoff sweep {4k,8k,12k,16k,24k}: K4 low-pass at 48 kHz + full metric suite
  cfg4: 8192 utterances: K3 16k->44.1k, K3 44.1k->48k, K4 stft_hard (n_fft 2048 / hop 441)
  cfg5: VCTK-shaped set (8 speakers x 300 ragged 2-8 s utterances), reference STFT setting at 48 kHz
        (n_fft 2229 / hop 480), utterance-sharded over the ranks + the metric-table all-reduce
        (run under torchrun for N > 1)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssr_eval_b200 import _native as N, dist  # noqa: E402
from ssr_eval_b200.engine import StftMetrics, PolyphaseResampler, HardLowpass, offsets_of  # noqa: E402


def sync_time(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def cfg3(dev, n_pairs=4096, L=240000):
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    tgt = 0.1 * torch.randn(n_pairs * L, generator=g, device=dev)
    off = offsets_of([L] * n_pairs)
    off_d = torch.from_numpy(off).to(dev)
    lp, eng = HardLowpass(2048, 441), StftMetrics(2048, 512)
    out = torch.empty((n_pairs, 4), dtype=torch.float64, device=dev)
    cut_hz = [4000, 8000, 12000, 16000, 24000]
    # lowpass(x, low_rate // 2, sr): low_rate = 2*cutoff, minus 1 when it equals sr (eval.py:404-405)
    ratios = [((2 * c - (1 if 2 * c == 48000 else 0)) // 2) / 24000 for c in cut_hz]

    def run():
        res = []
        for r in ratios:
            est = lp.apply_device(tgt, off, [lp.cut_bin(r)] * n_pairs, off_d)
            eng.metrics_device(est, tgt, off, N.METRIC_ALL, offsets_dev=off_d, out=out)
            res.append(out.mean(dim=0).cpu().numpy())
        return res
    run()
    ms, res = sync_time(run)
    evals = n_pairs * len(cut_hz)
    print(json.dumps({"config": "cfg3", "pairs": n_pairs, "cutoffs_hz": cut_hz, "ms": round(ms, 2),
                      "pair_evaluations_per_s": round(evals / (ms * 1e-3), 1),
                      "mean_metrics_per_cutoff": [[round(float(v), 4) for v in r] for r in res]}), flush=True)


def cfg4(dev, n=8192):
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    x16 = 0.1 * torch.randn(n * 80000, generator=g, device=dev)
    off16 = offsets_of([80000] * n)
    off16_d = torch.from_numpy(off16).to(dev)
    r1, r2, lp = PolyphaseResampler(44100, 16000), PolyphaseResampler(48000, 44100), HardLowpass(2048, 441)

    def run():
        x44, o44, o44d = r1.resample_device(x16, off16, off16_d)
        x48, o48, o48d = r2.resample_device(x44, o44, o44d)
        y = lp.apply_device(x48, o48, [lp.cut_bin(0.5)] * n, o48d)
        return int(o44[1]), int(o48[1]), float(y.abs().mean())
    run()
    ms, (l44, l48, _) = sync_time(run)
    alg = 4 * n * (80000 + l44) + 4 * n * (l44 + l48) + 8 * n * l48
    print(json.dumps({"config": "cfg4", "utterances": n, "lengths": [80000, l44, l48], "ms": round(ms, 2),
                      "utterances_per_s": round(n / (ms * 1e-3), 1),
                      "algorithmic_GBps": round(alg / (ms * 1e-3) / 1e9, 1)}), flush=True)


def cfg5(dev, speakers=8, utts=300):
    rank, world = dist.rank_world()
    rng = np.random.default_rng(5)
    lengths = (rng.uniform(2.0, 8.0, size=speakers * utts) * 48000).astype(np.int64)
    speaker_of = np.repeat(np.arange(speakers), utts)
    ids = dist.shard_indices(len(lengths), rank, world)
    mine = lengths[ids]
    # the SAME synthetic set on every rank / world size (fixed seed), then this rank's shard of it, so the
    # aggregated result must not depend on the number of GPUs
    g = torch.Generator(device=dev)
    g.manual_seed(50)
    all_off = offsets_of(lengths)
    full_t = 0.1 * torch.randn(int(all_off[-1]), generator=g, device=dev)
    full_e = full_t + 1e-3 * torch.randn(int(all_off[-1]), generator=g, device=dev)
    tgt = torch.cat([full_t[all_off[i]:all_off[i + 1]] for i in ids])
    est = torch.cat([full_e[all_off[i]:all_off[i + 1]] for i in ids])
    del full_t, full_e
    off = offsets_of(mine)
    off_d = torch.from_numpy(off).to(dev)
    eng = StftMetrics(2229, 480)  # AudioMetrics(48000): metrics.py:18-19

    def run():
        vals = eng.metrics_device(est, tgt, off, N.METRIC_ALL, offsets_dev=off_d).cpu().numpy()
        sums = np.zeros((speakers, 1, 4))
        counts = np.zeros((speakers, 1))
        np.add.at(sums[:, 0, :], speaker_of[ids], vals)
        np.add.at(counts[:, 0], speaker_of[ids], 1.0)
        sums, counts = dist.allreduce_table(sums, counts)
        return dist.mean_of_means(sums, counts)[1][0]
    run()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    avg = run()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"config": "cfg5", "world": world, "pairs": int(len(lengths)), "n_fft": 2229, "hop": 480,
                          "wall_ms_incl_d2h_and_allreduce": round(dt * 1e3, 2),
                          "pairs_per_s": round(len(lengths) / dt, 1),
                          "averaged": [round(float(v), 10) for v in avg]}), flush=True)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    which = set(sys.argv[1:]) or {"cfg3", "cfg4", "cfg5"}
    if world == 1:
        if "cfg3" in which:
            cfg3(dev)
        if "cfg4" in which:
            cfg4(dev)
    if "cfg5" in which:
        cfg5(dev)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
