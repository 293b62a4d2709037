"""N>1 host logic on CPU: frame partition and statistics reduction over gloo, world_size 2."""
import importlib
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    sh = importlib.import_module("qcrypto-ldpc_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sh.frame_range(total, rank, world)
    n = hi - lo
    # fake per-rank decoder statistics: every frame takes (index % 3) + 1 iterations, every 7th fails
    idx = np.arange(lo, hi)
    iters = idx % 3 + 1
    hist = np.bincount(iters, minlength=64).tolist()
    stats = {"frames": n, "failures": int((idx % 7 == 0).sum()), "iter_sum": int(iters.sum()), "kernel_launches": 1,
             "iter_hist": hist}
    red, times = sh.reduce_stats(stats, [10.0 + rank, 5.0 - rank], dist)
    q.put((rank, lo, hi, red, times))
    dist.destroy_process_group()


def test_frame_range_partition():
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("qcrypto-ldpc_b200.sharding")
    for total in (0, 1, 7, 64, 65536, 382966):
        for world in (1, 2, 3, 4, 8):
            r = [sh.frame_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_stats_reduction_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    total = 1001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(total)
    for rank, lo, hi, red, times in res:
        assert red["frames"] == total and red["failures"] == int((idx % 7 == 0).sum())
        assert red["iter_sum"] == int((idx % 3 + 1).sum()) and red["kernel_launches"] == 2
        assert red["iter_hist"][1:4] == np.bincount(idx % 3 + 1, minlength=4)[1:4].tolist()
        assert times == [11.0, 5.0]              # max over ranks
        assert abs(red["mean_iters"] - red["iter_sum"] / total) < 1e-12
    assert res[0][2] == res[1][1]                # contiguous ranges
