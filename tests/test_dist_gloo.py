"""CPU, world_size 2 over gloo: sharding + the one all-reduce / all-gather of the metric tables
reproduce the reference's mean-of-speaker-means aggregation (eval.py:200-216)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, values, speaker_of, n_spk, q):
    import torch.distributed as td
    from ssr_eval_b200 import dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    ids = dist.shard_indices(len(values), rank, world)
    sums = np.zeros((n_spk, 1, 4))
    counts = np.zeros((n_spk, 1))
    for i in ids:
        sums[speaker_of[i], 0] += values[i]
        counts[speaker_of[i], 0] += 1
    sums, counts = dist.allreduce_table(sums, counts)
    each, avg = dist.mean_of_means(sums, counts)
    full = dist.allgather_rows(values[ids], ids, len(values))
    merged = dist.gather_results({int(i): {"v": float(values[i][0])} for i in ids}, world)
    q.put((rank, avg, full, sorted(merged.keys())))
    td.destroy_process_group()


def test_sharded_aggregation_matches_reference_mean_of_means():
    rng = np.random.default_rng(3)
    counts = [5, 3, 7]
    values = rng.random((sum(counts), 4))
    speaker_of = np.repeat(np.arange(3), counts)
    keys = ("lsd", "log_sispec", "sispec", "ssim")
    spk_means = [oracle.dict_mean([dict(zip(keys, r)) for r in values[speaker_of == s]]) for s in range(3)]
    want = oracle.dict_mean(spk_means)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, values, speaker_of, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, avg, full, merged_keys in got:
        np.testing.assert_allclose(avg[0], [want[k] for k in keys], rtol=1e-13)
        np.testing.assert_array_equal(full, values)
        assert merged_keys == list(range(len(values)))
    # the global mean differs from the mean of means (unequal speaker sizes) -- the table must carry counts
    assert abs(values[:, 0].mean() - want["lsd"]) > 1e-6
