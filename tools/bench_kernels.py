#!/usr/bin/env python
"""Per-kernel device timings (CUDA events) at BASELINE-shaped sizes; prints one JSON line per case.
Not the contract benchmark (that is bench.py) -- used to fill DESIGN.md and to steer optimisation."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssr_eval_b200 import _native as N  # noqa: E402
from ssr_eval_b200.engine import (StftMetrics, PolyphaseResampler, HardLowpass, HardLowpassDense, offsets_of,  # noqa: E402
                                  pcm16_to_float_device)

dev = torch.device("cuda", 0)
PEAK = 6582.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


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


def report(name, ms, units, unit_name, alg_bytes):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "ms": round(ms, 4), unit_name + "_per_s": round(units / (ms * 1e-3), 1),
                      "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / PEAK, 4)}), flush=True)


def main():
    which = set(sys.argv[1:]) or {"k0", "k1", "k1_441", "k1blue", "k3", "k4", "k4d", "k6", "k7", "k8"}
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    if "k1" in which:
        n, L = 1024, 240000
        tgt = 0.1 * torch.randn(n * L, generator=g, device=dev)
        est = tgt + 1e-3 * torch.randn(n * L, generator=g, device=dev)
        off = offsets_of([L] * n)
        off_d = torch.from_numpy(off).to(dev)
        eng = StftMetrics(2048, 512)
        out = torch.empty((n, 4), dtype=torch.float64, device=dev)
        for flags, name in ((1, "K1 2048/512 LSD"), (7, "K1 2048/512 LSD+log_sispec+sispec"),
                            (15, "K1+K2 2048/512 all four")):
            ms = timeit(lambda: eng.metrics_device(est, tgt, off, flags, offsets_dev=off_d, out=out))
            report(name, ms, n, "pairs", n * (8 * L + 32))
        del tgt, est
    if "k1_441" in which:
        # evaluation at 44.1 kHz (metrics.py:18-19: n_fft 2048, hop 441): the 2048 kernel without the sample ring
        n, L = 1024, 220500
        tgt = 0.1 * torch.randn(n * L, generator=g, device=dev)
        est = tgt + 1e-3 * torch.randn(n * L, generator=g, device=dev)
        off = offsets_of([L] * n)
        off_d = torch.from_numpy(off).to(dev)
        eng = StftMetrics(2048, 441)
        out = torch.empty((n, 4), dtype=torch.float64, device=dev)
        for flags, name in ((1, "K1 2048/441 LSD (5 s @ 44.1k)"), (15, "K1+K2 2048/441 all four (5 s @ 44.1k)")):
            ms = timeit(lambda: eng.metrics_device(est, tgt, off, flags, offsets_dev=off_d, out=out))
            report(name, ms, n, "pairs", n * (8 * L + 32))
        del tgt, est
    if "k1_f64est" in which:
        # float64 estimates (what the IIR low-pass keys of setting_lowpass_filtering hand to the metrics): generic kernel
        for n_fft, hop, L, n in ((2048, 441, 220500, 256), (2229, 480, 240000, 128)):
            tgt = 0.1 * torch.randn(n * L, generator=g, device=dev)
            est = (tgt + 1e-3 * torch.randn(n * L, generator=g, device=dev)).double()
            off = offsets_of([L] * n)
            off_d = torch.from_numpy(off).to(dev)
            eng = StftMetrics(n_fft, hop)
            out = torch.empty((n, 4), dtype=torch.float64, device=dev)
            for flags, name in ((1, "LSD"), (15, "all four")):
                ms = timeit(lambda: eng.metrics_device(est, tgt, off, flags, offsets_dev=off_d, out=out), iters=2, warm=1)
                report("K1 float64-estimate %d/%d %s" % (n_fft, hop, name), ms, n, "pairs", n * (12 * L + 32))
            del tgt, est
    if "k0" in which:
        n = 1024 * 240000 * 2
        pcm = torch.randint(-32768, 32767, (n,), dtype=torch.int16, device=dev)
        dst = torch.empty(n, dtype=torch.float32, device=dev)
        ms = timeit(lambda: pcm16_to_float_device(pcm, out=dst))
        report("K0 pcm16 -> float32 (1024 pairs x 5 s @ 48k, both signals)", ms, 1024, "pairs", 6 * n)
        del pcm, dst
    if "k1blue" in which:
        n, L = 256, 240000
        tgt = 0.1 * torch.randn(n * L, generator=g, device=dev)
        est = tgt + 1e-3 * torch.randn(n * L, generator=g, device=dev)
        off = offsets_of([L] * n)
        off_d = torch.from_numpy(off).to(dev)
        eng = StftMetrics(2229, 480)
        out = torch.empty((n, 4), dtype=torch.float64, device=dev)
        for flags, name in ((1, "K1 2229/480 (PFA 3x743) LSD"), (15, "K1+K2 2229/480 all four")):
            ms = timeit(lambda: eng.metrics_device(est, tgt, off, flags, offsets_dev=off_d, out=out), iters=3, warm=1)
            report(name, ms, n, "pairs", n * (8 * L + 32))
        del tgt, est
    if "k3" in which:
        for up, down, L_in, name in ((160, 147, 220500, "K3 44.1k->48k (160/147, 3201 taps)"),
                                     (441, 160, 80000, "K3 16k->44.1k (441/160, 8821 taps)")):
            n = 1024
            x = 0.1 * torch.randn(n * L_in, generator=g, device=dev)
            rs = PolyphaseResampler(up, down)
            off = offsets_of([L_in] * n)
            off_d = torch.from_numpy(off).to(dev)
            ms = timeit(lambda: rs.resample_device(x, off, off_d))
            report(name, ms, n, "utterances", 4 * n * (L_in + rs.out_len(L_in)))
            del x
    if "k4" in which:
        n, L = 1024, 220500
        x = 0.1 * torch.randn(n * L, generator=g, device=dev)
        lp = HardLowpass(2048, 441)
        off = offsets_of([L] * n)
        off_d = torch.from_numpy(off).to(dev)
        cuts = [lp.cut_bin(12000 / 22050)] * n
        ms = timeit(lambda: lp.apply_device(x, off, cuts, off_d))
        report("K4 stft_hard 2048/441 (5 s @ 44.1k)", ms, n, "utterances", 8 * n * L)


    if "k4d" in which:
        n, L = 64, 220500
        x = 0.1 * torch.randn(n * L, generator=g, device=dev)
        lp = HardLowpassDense(2048, 441)
        off = offsets_of([L] * n)
        off_d = torch.from_numpy(off).to(dev)
        cuts = [lp.cut_bin(12000 / 22050)] * n
        ms = timeit(lambda: lp.apply_device(x, off, cuts, off_d), iters=2, warm=1)
        frames = n * (L // 441 + 1)
        flop = 2.0 * frames * 2048 * (2 * 1025 + 2 * 2048)
        print(json.dumps({"case": "K4d dense stft_hard 2048/441 (5 s @ 44.1k, 64 utterances)", "ms": round(ms, 3),
                          "utterances_per_s": round(n / (ms * 1e-3), 1), "fp32_TFLOPs": round(flop / (ms * 1e-3) / 1e12, 2)}),
              flush=True)
        del x
    if "k8" in which:
        import ctypes
        n, L = 256, 240000
        a = 0.1 * torch.randn(n * L, generator=g, device=dev)
        x = torch.roll(a, 1105) + 1e-3 * torch.randn(n * L, generator=g, device=dev)
        off = offsets_of([L] * n)
        off_d = torch.from_numpy(off).to(dev)
        out = torch.empty(n, dtype=torch.int64, device=dev)
        need = N.lib().ssr_xcorr_workspace_bytes(ctypes.c_void_p(off.ctypes.data), n)
        ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
        vp = ctypes.c_void_p

        def run8():
            N.check(N.lib().ssr_xcorr_argmax_batched(vp(a.data_ptr()), vp(x.data_ptr()), vp(off.ctypes.data), vp(off_d.data_ptr()),
                                                     n, vp(out.data_ptr()), vp(ws.data_ptr()), ws.numel(),
                                                     vp(torch.cuda.current_stream().cuda_stream)), "ssr_xcorr_argmax_batched")
        ms = timeit(run8, iters=3, warm=1)
        report("K8 xcorr argmax (256 pairs x 5 s @ 48k, FFT 2^19)", ms, n, "pairs", n * 8 * L)
        del a, x, ws
    if "k6" in which:
        import ctypes
        from ssr_eval_b200.engine import SpliceIstft
        n, L = 1024, 220500
        x = 0.1 * torch.randn(n * L, generator=g, device=dev)
        o = x + 0.01 * torch.randn(n * L, generator=g, device=dev)
        sp = SpliceIstft(2048, 512)
        off = offsets_of([L] * n)
        off_d = torch.from_numpy(off).to(dev)
        cb = torch.full((n,), 372, dtype=torch.int32, device=dev)
        y = torch.empty_like(o)
        vp = ctypes.c_void_p

        def run6():
            N.check(N.lib().ssr_stft_splice_istft_batched(sp._plan, vp(x.data_ptr()), vp(o.data_ptr()), vp(off.ctypes.data),
                                                          vp(off_d.data_ptr()), n, vp(cb.data_ptr()), vp(y.data_ptr()),
                                                          vp(torch.cuda.current_stream().cuda_stream)),
                    "ssr_stft_splice_istft_batched")
        ms = timeit(run6, iters=3, warm=1)
        report("K6 postprocessing splice + ISTFT 2048/512 (5 s @ 44.1k)", ms, n, "utterances", 12 * n * L)
        del x, o, y
    if "k7" in which:
        from scipy.signal import butter, sosfilt_zi
        import ctypes
        # the recursion is sequential in time: a launch takes ~ (samples x 4 dependent FP64 operations) however few
        # utterances it holds, so throughput grows with the batch until the schedulers are full
        for n in (4096, 16384):
            L = 220500
            x = 0.1 * torch.randn(n * L, generator=g, device=dev)
            off = offsets_of([L] * n)
            off_d = torch.from_numpy(off).to(dev)
            sos = np.ascontiguousarray(butter(8, 8000 / 22050, output="sos"))
            zi = np.ascontiguousarray(sosfilt_zi(sos))
            edge = 3 * (2 * sos.shape[0] + 1)
            y = torch.empty(n * L, dtype=torch.float64, device=dev)
            need = N.lib().ssr_sosfiltfilt_workspace_bytes(ctypes.c_void_p(off.ctypes.data), n, edge)
            ws = torch.empty(int(need), dtype=torch.uint8, device=dev)

            def run():
                N.check(N.lib().ssr_sosfiltfilt_batched(
                    ctypes.c_void_p(sos.ctypes.data), sos.shape[0], ctypes.c_void_p(zi.ctypes.data), edge,
                    ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(off.ctypes.data), ctypes.c_void_p(off_d.data_ptr()), n,
                    ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "ssr_sosfiltfilt_batched")
            ms = timeit(run, iters=2, warm=1)
            report("K7 sosfiltfilt butter order 8 (4 sections), %d x 5 s @ 44.1k" % n, ms, n, "utterances", n * L * (4 + 8))
            del x, y, ws


if __name__ == "__main__":
    main()
