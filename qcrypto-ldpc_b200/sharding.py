"""Frame sharding across GPUs (one process per GPU) and the host-side statistics reduction.

Frames are independent (AFF3CT's n_frames semantics; ML/BPSK_nrldpc_sim_RM_FP.m:27 parfor), so the batch is
cut into contiguous ranges, one per rank, with NO collective on the decode path.  The only reduction is the sum
of the per-GPU {frames, failures, iteration sum, iteration histogram} and the max of the device times, done on
CPU tensors (gloo) so NCCL never touches the hot path.
"""
import numpy as np

STAT_KEYS = ("frames", "failures", "iter_sum", "kernel_launches")


def frame_range(total_frames, rank, world):
    """contiguous range [lo, hi) of rank `rank`: sizes differ by at most one frame"""
    base, rem = divmod(int(total_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stats_to_vector(stats, hist_bins=64):
    v = [int(stats[k]) for k in STAT_KEYS] + [int(x) for x in stats["iter_hist"][:hist_bins]]
    return np.asarray(v, dtype=np.int64)


def vector_to_stats(vec):
    vec = [int(x) for x in vec]
    out = {k: vec[i] for i, k in enumerate(STAT_KEYS)}
    out["iter_hist"] = vec[len(STAT_KEYS):]
    out["fer"] = out["failures"] / max(1, out["frames"])
    out["mean_iters"] = out["iter_sum"] / max(1, out["frames"])
    return out


def reduce_stats(stats, times_ms, dist=None):
    """sum the statistics and take the max of the times over all ranks; `dist` is torch.distributed (initialised)
    or None for a single process.  Works on CPU tensors only."""
    import torch
    cnt = torch.from_numpy(stats_to_vector(stats))
    t = torch.tensor([float(x) for x in times_ms], dtype=torch.float64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return vector_to_stats(cnt.numpy()), [float(x) for x in t]
