#!/usr/bin/env python
"""bench.py -- utterance-pairs/s of the fused STFT+LSD hot path (BASELINE.json config[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

Workload (N=1): 1024 synthetic 48 kHz pairs, L = 240000 samples (5 s), n_fft 2048 / hop 512, LSD
(BASELINE.json configs[1]).  A step = one pass of the hot path over that batch.  For N>1 every rank
owns its own 1024 pairs (weak scaling) and the step ends with ONE NCCL all-reduce of the scalar
metric accumulators.  Inputs (1.97 GB per rank) are far larger than the 126 MB L2, so no flush is
needed between iterations.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_FFT, HOP, SR, LENGTH = 2048, 512, 48000, 240000
METRIC = "utterance_pairs_per_sec_stft_lsd_48k_5s"
UNIT = "pairs/s"


def algorithmic_bytes(n_pairs, length):
    """every sample of est and target read once + 32 B of results per pair."""
    return n_pairs * (2 * length * 4 + 32)


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference algorithm) on the host cores
# --------------------------------------------------------------------------------------------
_cpu_data = None


_cpu_barrier = None


def _cpu_init(barrier):
    global _cpu_data, _cpu_barrier
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)
    import oracle  # noqa: F401
    _cpu_barrier = barrier
    rng = np.random.default_rng(os.getpid())
    waves = []
    for s in range(2):  # the CPU cost of STFT+LSD does not depend on the sample values
        t = (0.1 * rng.standard_normal(LENGTH)).astype(np.float32)
        e = (t + 1e-3 * rng.standard_normal(LENGTH)).astype(np.float32)
        waves.append((e, t))
    _cpu_data = waves


def _cpu_warm(_):
    _cpu_work(1)
    _cpu_barrier.wait()  # every worker takes exactly one warm-up task
    return 0


def _cpu_work(n):
    import oracle
    acc = 0.0
    for i in range(n):
        e, t = _cpu_data[i % len(_cpu_data)]
        acc += oracle.evaluation(e, t, n_fft=N_FFT, hop=HOP, which=("lsd",))["lsd"]
    return acc


def usable_cores():
    """Host threads this process may actually use: the affinity mask, capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    env = os.environ.get("SSR_BENCH_CPU_CORES")
    return int(env) if env else n


class CpuArm:
    """Persistent process pool, one worker per host core, each running the oracle's
    STFT(float64 pocketfft)+LSD(torch float32) on 5 s pairs."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or usable_cores()
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init, initargs=(ctx.Barrier(self.cores),))
        self.pool.map(_cpu_warm, range(self.cores), chunksize=1)  # imports + caches warm in EVERY worker

    def step(self, pairs_per_core):
        t0 = time.perf_counter()
        self.pool.map(_cpu_work, [pairs_per_core] * self.cores, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    per_core = args.cpu_pairs_per_core
    for _ in range(args.warmup):
        arm.step(per_core)
    t = sum(arm.step(per_core) for _ in range(args.steps))
    arm.close()
    pairs = per_core * arm.cores * args.steps
    value = pairs / t
    sample = "%d pairs/step (%d per core) of the N=1 workload, oracle STFT(f64)+LSD(f32)" % (per_core * arm.cores, per_core)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 48kHz 5s pairs, n_fft 2048 hop 512, STFT+LSD", "sample_rate": SR,
                   "length": LENGTH, "n_fft": N_FFT, "hop": HOP},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        from datetime import datetime
        parsed = []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                parsed.append((ts, float(f[1]), float(f[2]), f[4:8]))
            except ValueError:
                continue
        # samples taken inside the timed regions (device-resident + e2e); 0.15 s slack for the 100 ms period
        inside = [p for p in parsed if self.t0 is not None and self.t0 - 0.15 <= p[0] <= (self.t1 or 1e30) + 0.15]
        use = inside if inside else parsed
        for _, a, b, flags in use:
            sm.append(a)
            mx.append(b)
            for name, v in zip(names, flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_total": len(parsed), "reasons": sorted(reasons)}


def make_gpu_batch(n_pairs, device, rank):
    """Synthetic batch built ON the device from 16 seeded CPU utterances: target = shifted / scaled
    base utterance + a little noise; est = the repo's own STFT hard low-pass (cutoff 12 kHz) of it --
    the parity-critical kind of estimate."""
    import torch
    from ssr_eval_b200.engine import HardLowpass, offsets_of
    from ssr_eval_b200.synth import speech_like
    base = torch.from_numpy(np.stack([speech_like(LENGTH, sr=SR, seed=5000 + 16 * rank + i) for i in range(16)])).to(device)
    g = torch.Generator(device=device)
    g.manual_seed(1234 + rank)
    tgt = torch.empty(n_pairs * LENGTH, dtype=torch.float32, device=device)
    for i in range(n_pairs):
        shift = int(torch.randint(0, LENGTH, (1,), generator=g, device=device))
        gain = 0.5 + 0.5 * float(torch.rand(1, generator=g, device=device))
        seg = torch.roll(base[i % 16], shift) * gain
        seg = seg + 1e-4 * torch.randn(LENGTH, generator=g, device=device)
        tgt[i * LENGTH:(i + 1) * LENGTH] = seg
    off = offsets_of([LENGTH] * n_pairs)
    lp = HardLowpass(2048, 441)
    est = lp.apply_device(tgt, off, [lp.cut_bin(12000 / (SR / 2))] * n_pairs)
    torch.cuda.synchronize()
    return est, tgt, off


def run_gpu(args):
    import torch
    import torch.distributed as td
    from ssr_eval_b200 import _native as N
    from ssr_eval_b200.engine import StftMetrics
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    n_pairs = args.pairs
    est, tgt, off = make_gpu_batch(n_pairs, dev, rank)
    off_dev = torch.from_numpy(off).to(dev)
    eng = StftMetrics(N_FFT, HOP)
    out = torch.empty((n_pairs, 4), dtype=torch.float64, device=dev)
    flags = N.METRIC_LSD
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    count_t = torch.full((1,), float(n_pairs), dtype=torch.float64, device=dev)
    pending = []

    def step():
        eng.metrics_device(est, tgt, off, flags, offsets_dev=off_dev, out=out)
        if world > 1:
            # the single all-reduce of the scalar metric accumulators (sum of LSD, pair count) over
            # NCCL; enqueued asynchronously so ranks are not lock-stepped, completed inside the timed region
            a = torch.cat((out[:, 0].sum(dim=0, keepdim=True), count_t))
            pending.append((a, td.all_reduce(a, async_op=True)))

    def drain():
        for a, w in pending:
            w.wait()
        if pending:
            acc.copy_(pending[-1][0])
        pending.clear()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()
    # ---- device-resident timing: exactly K steps, CUDA events on the launching stream
    sampler.mark_begin()
    N.timing_enable(True)
    N.timing_collect()
    launches0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    drain()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = N.launch_count() - launches0
    k1_ms, k1_n = N.timing_collect()
    N.timing_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n_pairs * args.steps / (ms * 1e-3)
    lsd_mean = float(out[:, 0].mean().item())
    timed_first = out[:2].cpu().numpy()  # values of the timed launches (out is reused by the context runs below)

    # ---- end to end through the public host API: pinned host buffers -> H2D -> kernels -> D2H.
    # Headline e2e: the host batch is 16-bit PCM -- the sample format of the wav files the reference reads
    # (VCTK, and everything it writes with sf.write); librosa.load turns a sample s into float32(s) / 32768, which
    # is what K0 does on the device after a 2-byte-per-sample upload.  The same batch as float32 host buffers
    # (4 bytes per sample over PCIe) is timed too and reported as e2e_f32_host.
    from ssr_eval_b200.engine import HostPipeline
    pipe = HostPipeline(eng, n_pairs, LENGTH)

    def to_pcm16(x):
        return torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def time_e2e(est_h, tgt_h):
        pipe.run(est_h, tgt_h, off, flags)  # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r = pipe.run(est_h, tgt_h, off, flags)
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(tt, op=td.ReduceOp.MAX)
        return world * n_pairs * e2e_steps / float(tt.item()), r

    est_q, tgt_q = to_pcm16(est), to_pcm16(tgt)
    # device-resident result of the very same (dequantised) pairs, for the equality check below
    want_q = eng.metrics_device(est_q.float() / 32768.0, tgt_q.float() / 32768.0, off, flags, offsets_dev=off_dev)
    want_q = want_q.cpu().numpy()
    est_qh, tgt_qh = est_q.cpu().pin_memory(), tgt_q.cpu().pin_memory()
    del est_q, tgt_q
    e2e_value, res = time_e2e(est_qh, tgt_qh)
    assert np.array_equal(res[:, 0], want_q[:, 0]), "e2e (PCM16 host buffers) differs from the device-resident run"
    del est_qh, tgt_qh
    est_h = est.cpu().pin_memory()
    tgt_h = tgt.cpu().pin_memory()
    e2e_f32_value, res32 = time_e2e(est_h, tgt_h)
    assert abs(float(np.mean(res32[:, 0])) - lsd_mean) < 1e-9
    del est_h, tgt_h
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    # ---- context numbers (not part of the contract metric): the reference's own per-pair call computes all four
    # metrics, and at 48 kHz its STFT is n_fft 2229 / hop 480 (metrics.py:18-19); same device-resident batch
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        def rate(engine, fl, iters):
            for _ in range(2):
                engine.metrics_device(est, tgt, off, fl, offsets_dev=off_dev, out=out)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                engine.metrics_device(est, tgt, off, fl, offsets_dev=off_dev, out=out)
            b.record()
            torch.cuda.synchronize()
            return n_pairs * iters / (a.elapsed_time(b) * 1e-3)
        eng48 = StftMetrics(2229, 480)
        peak_h = 6650.0
        try:
            peak_h = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass

        def entry(pairs_per_s):  # same algorithmic bytes per pair as the contract metric
            gbs = pairs_per_s * algorithmic_bytes(1, LENGTH) / 1e9
            return {"pairs_per_s": pairs_per_s, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak_h}
        extras = {"unit": UNIT, "note": "device-resident, same batch; context only (the reference's per-pair call computes "
                                        "all four metrics; at 48 kHz its STFT is n_fft 2229 / hop 480, metrics.py:18-19)",
                  "n_fft2048_hop512_all_four_metrics": entry(rate(eng, N.METRIC_ALL, 5)),
                  "n_fft2229_hop480_lsd": entry(rate(eng48, N.METRIC_LSD, 2)),
                  "n_fft2229_hop480_all_four_metrics": entry(rate(eng48, N.METRIC_ALL, 2))}
        del eng48

    check = {"mean_lsd": lsd_mean}
    if rank == 0 and not args.no_oracle_check:
        # parity inside the run: two pairs of the timed batch (pair 0: hard-low-passed estimate) scored by the CPU
        # oracle -- the checker, outside every timed region -- against the values the timed kernels produced
        import oracle
        got_all = eng.metrics_device(est[:2 * LENGTH], tgt[:2 * LENGTH], off[:3], N.METRIC_ALL).cpu().numpy()
        timed = timed_first
        diffs = {m: 0.0 for m in N.METRIC_NAMES}
        for i in range(2):
            e_i = est[i * LENGTH:(i + 1) * LENGTH].cpu().numpy()
            t_i = tgt[i * LENGTH:(i + 1) * LENGTH].cpu().numpy()
            want = oracle.evaluation(e_i, t_i, n_fft=N_FFT, hop=HOP)
            # (another batch size groups the frames into other work items: float64 partial sums regrouped, ~1e-13)
            assert abs(got_all[i, 0] - timed[i, 0]) < 1e-9, "LSD of the timed launch and of the all-metrics launch differ"
            for j, m in enumerate(N.METRIC_NAMES):
                diffs[m] = max(diffs[m], abs(float(got_all[i, j]) - float(want[m])))
        check["oracle_pairs"] = 2
        check["max_abs_diff"] = diffs
        check["tolerance"] = {"lsd": 1e-4, "log_sispec": 1e-4, "ssim": 1e-3}
        check["ok"] = bool(diffs["lsd"] <= 1e-4 and diffs["ssim"] <= 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k1_avg_ms = k1_ms / max(k1_n, 1)
        achieved = algorithmic_bytes(n_pairs, LENGTH) / (k1_avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "kernel": "k_stft_metrics_2048", "kernel_ms": k1_avg_ms,
                    "kernel_share_of_step": k1_ms / ms if ms else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    "note": "float64 FFT: the FP64 pipe binds, not HBM (see DESIGN.md roofline section)"}
        # secondary roofline: the pipe that actually binds.  K1's hot loop issues 595 FP64 instructions per thread and
        # frame (376 DADD + 111 DMUL + 108 DFMA, execution counts of the ncu source page, profiles/r02_ncu_summary.md),
        # 128 threads per frame; the peak is measured in this run (DFMA chains, ssr_probe_fp64_rate)
        roofline_fp64 = None
        try:
            fp64_peak = N.probe_fp64_rate(None)
            frames = 1 + LENGTH // HOP
            fp64_instr = 595.0 * 128.0 * frames * n_pairs
            ach = fp64_instr / (k1_avg_ms * 1e-3)
            roofline_fp64 = {"bound": "fp64_pipe", "achieved": ach / 1e9, "peak": fp64_peak / 1e9, "unit": "G thread-instr/s",
                             "frac": ach / fp64_peak, "fp64_instr_per_thread_per_frame": 595,
                             "peak_source": "measured in this run (ssr_probe_fp64_rate: 8 independent DFMA chains per thread)"}
        except Exception as ex:  # the probe is measurement support only
            roofline_fp64 = {"error": str(ex)}
        traffic_file = os.path.join(ROOT, "profiles", "k1_traffic_bytes.json")
        if os.path.exists(traffic_file):
            try:
                # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, scaled per pair
                tj = json.load(open(traffic_file))
                roofline["traffic"] = tj["dram_bytes_per_pair"] * n_pairs
                roofline["traffic_source"] = "ncu capture at commit %s (%s), scaled per pair; not re-measured in this run" % (
                    tj.get("captured_at_commit", "?"), tj.get("source", "?"))
            except Exception:
                pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            arm = CpuArm()
            per_core = args.cpu_pairs_per_core
            arm.step(1)
            tt = arm.step(per_core)
            arm.close()
            cpu = {"value": per_core * arm.cores / tt, "unit": UNIT, "cores": arm.cores, "kind": "port",
                   "sample": "%d pairs (%d per core) of the same workload, oracle STFT(f64)+LSD(f32), one process per core"
                             % (per_core * arm.cores, per_core)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: %d pairs/GPU, 48kHz 5s (L=240000), n_fft 2048 hop 512, fused STFT+LSD" % n_pairs,
                       "pairs_per_gpu": n_pairs, "sample_rate": SR, "length": LENGTH, "n_fft": N_FFT, "hop": HOP,
                       "l2_policy": "inputs (1.97 GB/GPU) larger than L2, no flush", "parallelism": "pairs sharded, 1 all-reduce"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * n_pairs * LENGTH * 2),
                    "d2h_bytes_per_step": int(n_pairs * 4 * 8), "steps": e2e_steps,
                    "host_format": "int16 PCM (wav sample format), converted on the device (K0)"},
            "e2e_f32_host": {"value": e2e_f32_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * n_pairs * LENGTH * 4),
                             "d2h_bytes_per_step": int(n_pairs * 4 * 8), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_fp64": roofline_fp64,
            "cpu_baseline": cpu,
            "check": check,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def run_cfg5(args):
    """BASELINE.json configs[4]: VCTK-shaped synthetic set (8 speakers x 300 ragged 2-8 s utterances at 48 kHz), all
    four metrics at the reference's own 48 kHz STFT setting (n_fft 2229 / hop 480, metrics.py:18-19), utterances
    sharded i % world over the ranks, ONE all-reduce of the per-speaker sum / count table (eval.py:200-216) per step.
    STRONG scaling: the 2400 pairs are the whole job whatever N is.  A step = scoring the rank's shard + D2H of the
    (n, 4) result + the table all-reduce + the mean of speaker means."""
    import torch
    import torch.distributed as td
    from ssr_eval_b200 import _native as N, dist
    from ssr_eval_b200.engine import StftMetrics, offsets_of
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    speakers, utts = 8, 300
    rng = np.random.default_rng(5)
    lengths = (rng.uniform(2.0, 8.0, size=speakers * utts) * 48000).astype(np.int64)
    speaker_of = np.repeat(np.arange(speakers), utts)
    ids = dist.shard_indices(len(lengths), rank, world)
    g = torch.Generator(device=dev)
    g.manual_seed(50)
    all_off = offsets_of(lengths)
    # the SAME synthetic set for every world size (fixed seed), then this rank's shard of it
    full_t = 0.1 * torch.randn(int(all_off[-1]), generator=g, device=dev)
    full_e = full_t + 1e-3 * torch.randn(int(all_off[-1]), generator=g, device=dev)
    tgt = torch.cat([full_t[all_off[i]:all_off[i + 1]] for i in ids])
    est = torch.cat([full_e[all_off[i]:all_off[i + 1]] for i in ids])
    del full_t, full_e
    off = offsets_of(lengths[ids])
    off_d = torch.from_numpy(off).to(dev)
    eng = StftMetrics(2229, 480)

    def step():
        vals = eng.metrics_device(est, tgt, off, N.METRIC_ALL, offsets_dev=off_d).cpu().numpy()
        sums = np.zeros((speakers, 1, 4))
        counts = np.zeros((speakers, 1))
        np.add.at(sums[:, 0, :], speaker_of[ids], vals)
        np.add.at(counts[:, 0], speaker_of[ids], 1.0)
        sums, counts = dist.allreduce_table(sums, counts)
        return dist.mean_of_means(sums, counts)[1][0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        avg = step()
    launches0 = N.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        avg = step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(dt, op=td.ReduceOp.MAX)
    dt = float(dt.item())
    if rank == 0:
        print(json.dumps({
            "metric": "utterance_pairs_per_sec_all_metrics_vctk_shaped_48k", "value": len(lengths) * args.steps / dt,
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "gpu_launches": int(N.launch_count() - launches0),
            "config": {"workload": "configs[4]: 8 speakers x 300 ragged 2-8 s utterances at 48 kHz, n_fft 2229 hop 480, "
                                   "lsd+log_sispec+sispec+ssim, utterance-sharded, 1 all-reduce of the speaker table",
                       "pairs": int(len(lengths)), "n_fft": 2229, "hop": 480,
                       "timing": "wall clock around K steps incl. D2H of the results and the all-reduce, max over ranks"},
            "averaged": [float(v) for v in avg]}), flush=True)
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2 = BASELINE configs[1] (the contract metric, default); cfg5 = configs[4], strong scaling")
    ap.add_argument("--pairs", type=int, default=1024, help="pairs per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-pairs-per-core", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the two-pair CPU-oracle parity check")
    ap.add_argument("--no-extras", action="store_true", help="skip the context measurements (all four metrics, n_fft 2229)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "cfg5":
        run_cfg5(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
