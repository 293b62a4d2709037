#!/usr/bin/env python3
"""BASELINE config 4: frame-sharded decode of a synthetic 10 Gbit sifted-key stream with early termination and a
host-side reduction of the FER / iteration statistics.  STRONG scaling: the stream is fixed, every rank decodes a
contiguous range of frames (qcrypto-ldpc_b200/sharding.py), no collective on the decode path.

    python tools/stream_10gbit.py                                             # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/stream_10gbit.py --gpus N [--gbit 10] [--e2e]

Frames are generated on the device chunk by chunk (not timed); the decode of every chunk is timed with CUDA events on
the launch stream; --e2e also pushes every chunk through the host-pointer call qldpc_decode_bits (pinned buffers).
Prints ONE JSON line on rank 0."""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402  (workload constants and the frame synthesiser)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--gbit", type=float, default=10.0, help="information bits in the stream, in Gbit")
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    q = importlib.import_module("qcrypto-ldpc_b200")
    sh = importlib.import_module("qcrypto-ldpc_b200.sharding")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    code = q.Code.from_qc_file(q.data_path(B.CODE_FILE))
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=B.MAX_ITER, early_stop=True,
                    norm_factor=B.NORM, out_mode=q.OUT_INFO, device=local_rank)
    N, K = code.n, code.k
    total_frames = int(-(-args.gbit * 1e9 // K))
    lo, hi = sh.frame_range(total_frames, rank, world)
    st = torch.cuda.current_stream().cuda_stream
    kw = K // 32
    L = q.lib()
    dec_ms, e2e_s, n_chunks, wrong = 0.0, 0.0, 0, 0
    out = torch.empty((args.chunk, dec.out_words), dtype=torch.int32, device=dev)
    ok = torch.empty(args.chunk, dtype=torch.uint8, device=dev)
    it = torch.empty(args.chunk, dtype=torch.int16, device=dev)
    if args.e2e:
        h_bits = torch.empty((args.chunk, dec.cw_words), dtype=torch.int32).pin_memory()
        h_out = torch.empty((args.chunk, dec.out_words), dtype=torch.int32).pin_memory()
        h_ok = torch.empty(args.chunk, dtype=torch.uint8).pin_memory()
        h_it = torch.empty(args.chunk, dtype=torch.int16).pin_memory()
    # warm-up (not part of the stream)
    msg, noisy, known, llr = B.synth_frames_device(torch, dec, min(args.chunk, 4096), K, N, 7, dev, st)
    dec.decode_device(llr.data_ptr(), 0, llr.shape[0], out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
    torch.cuda.synchronize()
    dec.reset_stats()
    if world > 1:
        dist.barrier()
    wall0 = time.perf_counter()
    f0 = lo
    while f0 < hi:
        F = min(args.chunk, hi - f0)
        msg, noisy, known, llr = B.synth_frames_device(torch, dec, F, K, N, 1000003 * (f0 // args.chunk + 1) + 17, dev, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        dec.decode_device(llr.data_ptr(), 0, F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
        e1.record()
        torch.cuda.synchronize()
        dec_ms += e0.elapsed_time(e1)
        wrong += int((~((out[:F, :kw] == msg).all(dim=1) & ok[:F].bool())).sum())
        if args.e2e:
            h_bits[:F].copy_(noisy)
            h_known = known.cpu()
            t0 = time.perf_counter()
            rc = L.qldpc_decode_bits(dec.h, h_bits.data_ptr(), h_known.data_ptr(), None, B.LLR_NOISY, B.LLR_KNOWN, None, F,
                                     h_out.data_ptr(), h_ok.data_ptr(), h_it.data_ptr())
            e2e_s += time.perf_counter() - t0
            assert rc == 0
        del msg, noisy, known, llr
        n_chunks += 1
        f0 += F
    wall = time.perf_counter() - wall0
    stats = dec.stats()
    red, tmax = sh.reduce_stats(stats, [dec_ms, e2e_s * 1e3, wall * 1e3, float(wrong)], dist if world > 1 else None)
    if world > 1:
        import torch as _t
        w = _t.tensor([wrong], dtype=_t.int64); dist.all_reduce(w); wrong = int(w[0])
    if rank == 0:
        frames_expected = total_frames * (2 if args.e2e else 1)
        line = {"metric": "reconciled info Mbit/s", "unit": "Mbit/s", "n_gpus": world, "scaling": "strong",
                "value": total_frames * K / (tmax[0] * 1e-3) / 1e6, "decode_ms_max_over_ranks": tmax[0],
                "config": {"workload": "10 Gbit synthetic sifted-key stream, BG1 Z=384 int8 layered NMS 6/8, max 10 iterations, early "
                                       "termination, QBER 3 %, send-parity formulation; frame-sharded, host-side stats reduction",
                           "stream_info_bits": total_frames * K, "frames": total_frames, "chunk_frames": args.chunk,
                           "frames_rank0": hi - lo},
                "fer": wrong / total_frames, "frames_not_reconciled": wrong,
                "decoder_stats": {"frames": red["frames"], "failures": red["failures"], "mean_iters": red["mean_iters"],
                                  "iter_hist": red["iter_hist"][:12], "frames_expected": frames_expected},
                "wall_s_incl_generation_max": tmax[2] * 1e-3, "data": "synthetic", "dtype": "i8"}
        if args.e2e:
            line["e2e"] = {"value": total_frames * K / (tmax[1] * 1e-3) / 1e6, "unit": "Mbit/s", "ms_max_over_ranks": tmax[1],
                           "api": "qldpc_decode_bits, pinned host buffers, per chunk"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
