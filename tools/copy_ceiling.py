#!/usr/bin/env python3
"""Platform ceiling of the host-pointer call at N GPUs: the bytes `qldpc_decode_bits` moves per step (packed key bits in,
packed bits / ok / iteration counts out, BASELINE config 2 sizes) copied between pinned host memory and the device with NO
decoder at all -- H2D and D2H concurrently on two streams, one process per GPU, every rank at the same time.

    python tools/copy_ceiling.py                                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/copy_ceiling.py

Rank 0 prints one JSON line: ms per step (max over ranks) and the information-bit rate a decoder of zero cost would show
through this call.  If that rate at 8 GPUs is below 7 x the one-GPU e2e figure of bench.py, the host side of the box, not the
library, bounds the e2e scaling."""
import json
import os
import sys

import torch
import torch.distributed as dist

F, N, K = 65536, 26112, 8448
IN_BYTES = F * (N // 8)                      # packed sifted-key bits
OUT_BYTES = F * (K // 8) + F + 2 * F         # packed info bits + ok flags + iteration counts


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    steps, warm = 10, 3
    h_in = torch.empty(IN_BYTES, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(OUT_BYTES, dtype=torch.uint8).pin_memory()
    h_in.random_(0, 256)
    d_in = torch.empty(IN_BYTES, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(OUT_BYTES, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d+d2h", "h2d", "d2h"):
        def step():
            if "h2d" in mode:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if "d2h" in mode:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = float(t[0])
    if rank == 0:
        ms = res["h2d+d2h"]
        print(json.dumps({"n_gpus": world, "h2d_bytes_per_gpu_step": IN_BYTES, "d2h_bytes_per_gpu_step": OUT_BYTES,
                          "ms_per_step_max_over_ranks": res, "h2d_GBps_per_gpu": IN_BYTES / res["h2d"] / 1e6,
                          "d2h_GBps_per_gpu": OUT_BYTES / res["d2h"] / 1e6,
                          "info_mbps_if_decoding_were_free": world * F * K / ms / 1e3}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
