"""Multi-GPU plumbing: utterance pairs shard embarrassingly across ranks (one
process per GPU); the only cross-rank step is the aggregation of eval.py:200-216, done as ONE
all-reduce(SUM) of the float64 per-(speaker, distortion) sum / count table plus one all-gather of the
per-item 4-tuples so rank 0 can emit the reference's full JSON.  NCCL over NVLink on GPUs; the same
code runs over gloo on CPU for the host-logic tests."""
import numpy as np
import torch
import torch.distributed as td


def rank_world():
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def _comm_device():
    if td.is_initialized() and td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def shard_indices(n_items, rank, world):
    """Round-robin shard: item i belongs to rank i % world."""
    return list(range(rank, n_items, world))


def allreduce_table(sums, counts):
    """One all-reduce of [sums | counts] (float64).  sums: (..., M), counts: (...)."""
    sums = np.asarray(sums, dtype=np.float64)
    counts = np.asarray(counts, dtype=np.float64)
    rank, world = rank_world()
    if world == 1:
        return sums.copy(), counts.copy()
    flat = torch.from_numpy(np.concatenate([sums.ravel(), counts.ravel()])).to(_comm_device())
    td.all_reduce(flat, op=td.ReduceOp.SUM)
    flat = flat.cpu().numpy()
    return flat[:sums.size].reshape(sums.shape), flat[sums.size:].reshape(counts.shape)


def mean_of_means(sums, counts):
    """eval.py:200-216: per-speaker mean over files, then mean over speakers (NOT a global mean).
    sums: (S, D, M), counts: (S, D) -> (each_speaker (S, D, M), averaged (D, M))."""
    each = sums / counts[..., None]
    return each, each.mean(axis=0)


def allgather_rows(local_rows, local_ids, n_total):
    """All-gather per-item rows (n_local, M) with their global ids into a dense (n_total, M)."""
    local_rows = np.asarray(local_rows, dtype=np.float64).reshape(len(local_ids), -1)
    rank, world = rank_world()
    M = local_rows.shape[1] if local_rows.size else 4
    full = np.full((n_total, M), np.nan)
    if world == 1:
        full[np.asarray(local_ids, dtype=np.int64)] = local_rows
        return full
    cap = (n_total + world - 1) // world
    buf = torch.full((cap, M + 1), -1.0, dtype=torch.float64)
    if len(local_ids):
        buf[:len(local_ids), 0] = torch.as_tensor(np.asarray(local_ids, dtype=np.float64))
        buf[:len(local_ids), 1:] = torch.from_numpy(local_rows)
    buf = buf.to(_comm_device())
    out = [torch.empty_like(buf) for _ in range(world)]
    td.all_gather(out, buf)
    for t in out:
        a = t.cpu().numpy()
        ok = a[:, 0] >= 0
        full[a[ok, 0].astype(np.int64)] = a[ok, 1:]
    return full


def gather_results(local, world):
    """Merge {key: nested-dict} results of all ranks (host objects; used by SSR_Eval_Helper.evaluate,
    whose per-item dicts may carry user-defined extra metrics)."""
    if world == 1:
        return dict(local)
    parts = [None] * world
    td.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged
