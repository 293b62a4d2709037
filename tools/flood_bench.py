#!/usr/bin/env python3
"""BASELINE config 3 throughput probe (run on the GPU box from the repository root):
flooding decoders on the long QKD block, QBER 3 %, syndrome formulation, device-resident inputs, CUDA events.

    python tools/flood_bench.py [--code qkd_psdpeg_n65536.qc] [--frames 2368] [--qber 0.03] [--out gpurun_out/flood.json]

Per decoder: key Mbit/s (N basis), mean sweeps, and the roofline of SURVEY.md 8d for off-chip flooding -- algorithmic bytes
per frame = 16 B x E x sweeps (var-to-check write + read, check-to-var write + read at 4 bytes; 2- and 1-byte messages for
the int16 / int8 tiers scale it to 8 / 4 B per edge) against the measured HBM peak (MEASURED_PEAKS.json)."""
import argparse
import importlib
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="qkd_psdpeg_n65536.qc")
    ap.add_argument("--frames", type=int, default=2368)
    ap.add_argument("--qber", type=float, default=0.03)
    ap.add_argument("--max-iter", type=int, default=50)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="", help="run only the decoders whose name contains this text")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    q = importlib.import_module("qcrypto-ldpc_b200")
    dev = torch.device("cuda", 0)
    code = q.Code.from_alist(q.data_path(args.code)) if args.code.endswith(".alist") else q.Code.from_qc_file(q.data_path(args.code))
    N, E, F, qber = code.n, code.edges, args.frames, args.qber
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    Np = (N + 31) // 32 * 32                                   # frames are packed into whole words
    x = torch.zeros((F, Np), dtype=torch.uint8, device=dev)
    x[:, :N] = torch.randint(0, 2, (F, N), dtype=torch.uint8, device=dev, generator=g)
    e = torch.zeros_like(x)
    e[:, :N] = (torch.rand((F, N), device=dev, generator=g) < qber).to(torch.uint8)
    w = (2 ** torch.arange(31, -1, -1, device=dev, dtype=torch.int64))

    def pack(b):
        v = (b.view(F, -1, 32).to(torch.int64) * w).sum(dim=-1)
        return torch.where(v >= 2**31, v - 2**32, v).to(torch.int32).contiguous()

    xb, yb = pack(x), pack(x ^ e)
    del x, e
    st = torch.cuda.current_stream().cuda_stream
    mag = math.log((1 - qber) / qber)
    results = []
    cases = (("SPA f32", q.RULE_SPA, q.DTYPE_F32, 1.0, mag, torch.float32, 16, 0),
             ("SPA f32 fast (QLDPC_FLAG_FAST_SPA)", q.RULE_SPA, q.DTYPE_F32, 1.0, mag, torch.float32, 16, q.FLAG_FAST_SPA),
             ("NMS 13/16 f32", q.RULE_NMS, q.DTYPE_F32, 0.8125, mag, torch.float32, 16, 0),
             ("NMS 6/8 i16", q.RULE_NMS, q.DTYPE_I16, 0.75, round(mag * 64), torch.int16, 8, 0),
             ("NMS 6/8 i8", q.RULE_NMS, q.DTYPE_I8, 0.75, round(mag * 4), torch.int8, 4, 0))
    for name, rule, dt, norm, m, tdt, bytes_per_edge, flags in cases:
        if args.only and args.only not in name:
            continue
        dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=rule, dtype=dt, max_iter=args.max_iter, early_stop=True,
                        norm_factor=norm, out_mode=q.OUT_ALL, flags=flags)
        syn = torch.empty((F, dec.syn_words), dtype=torch.int32, device=dev)
        dec.syndrome_device(xb.data_ptr(), F, syn.data_ptr(), st)
        llr = torch.empty((F, N), dtype=tdt, device=dev)
        dec.make_llr_device(yb.data_ptr(), 0, 0, float(m), 0.0, F, llr.data_ptr(), st)
        out = torch.empty((F, dec.cw_words), dtype=torch.int32, device=dev)
        ok = torch.empty(F, dtype=torch.uint8, device=dev)
        it = torch.empty(F, dtype=torch.int16, device=dev)
        for _ in range(args.warmup):
            dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
        ev[0].record()
        for r in range(args.reps):
            dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
            ev[r + 1].record()
        torch.cuda.synchronize()
        ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(args.reps))
        sweeps = float(it.float().mean())
        alg = bytes_per_edge * E * sweeps * F
        res = {"decoder": name, "kernel": dec.kernel_name, "frames": F, "qber": qber, "ok_frac": float(ok.float().mean()),
               "mean_sweeps": sweeps, "ms": ms, "key_mbps": F * N / ms / 1e3, "algorithmic_GBps": alg / ms / 1e6,
               "frac_of_hbm_peak": alg / ms / 1e6 / peak, "hbm_peak_GBps": peak, "all_reconciled": bool((out[:, :N // 32] == xb[:, :N // 32]).all())}
        print(json.dumps(res), flush=True)
        results.append(res)
        dec.close()
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
