#!/usr/bin/env python3
"""Throughput probe of the layered decoders that are NOT the headline int8 kernel (run on the GPU box from the repository
root): float / int16 layered min-sum on QC codes, the int8 kernel family on other lifting sizes, the layered schedule on an
arbitrary (.alist) H.  Syndrome formulation, BSC at --qber, device-resident inputs, CUDA events.

    python tools/layered_bench.py [--frames 4096] [--iters 10] [--out gpurun_out/layered.json]

Per case: kernel name, codeword Mbit/s (N basis), edge updates per second and the on-chip byte rate SURVEY.md 8d prescribes
for layered decoding (I x E x 4 state bytes per frame: read L, read R, write L, write R at the tier's width)."""
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
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--qber", type=float, default=0.03)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--early-stop", action="store_true", help="syndrome test after every iteration (the usual mode of operation)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    q = importlib.import_module("qcrypto-ldpc_b200")
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    mag = math.log((1 - args.qber) / args.qber)
    w = (2 ** torch.arange(31, -1, -1, device=dev, dtype=torch.int64))
    cases = (("BG1 Z=384 f32 NMS 0.75", "NR_1_1_384.qc", q.RULE_NMS, q.DTYPE_F32, 0.75, mag, torch.float32, 4),
             ("BG1 Z=384 f32 OMS 0.5", "NR_1_1_384.qc", q.RULE_OMS, q.DTYPE_F32, 1.0, mag, torch.float32, 4),
             ("BG1 Z=384 i16 NMS 6/8", "NR_1_1_384.qc", q.RULE_NMS, q.DTYPE_I16, 0.75, round(mag * 8), torch.int16, 2),
             ("BG1 Z=384 i8 NMS 6/8 (fixed iterations)", "NR_1_1_384.qc", q.RULE_NMS, q.DTYPE_I8, 0.75, 14, torch.int8, 1),
             ("BG2 Z=256 i8 NMS 6/8", "NR_2_0_256.qc", q.RULE_NMS, q.DTYPE_I8, 0.75, 14, torch.int8, 1),
             ("BG2 Z=112 i8 NMS 6/8", "NR_2_3_112.qc", q.RULE_NMS, q.DTYPE_I8, 0.75, 14, torch.int8, 1),
             ("BG2 Z=52 i8 NMS 6/8", "NR_2_6_52.qc", q.RULE_NMS, q.DTYPE_I8, 0.75, 14, torch.int8, 1),
             ("802.11n N=1944 Z=81 f32 NMS 0.75", "wifi_n1944_r12.qc", q.RULE_NMS, q.DTYPE_F32, 0.75, mag, torch.float32, 4),
             ("802.11n N=1944 Z=81 i8 NMS 6/8", "wifi_n1944_r12.qc", q.RULE_NMS, q.DTYPE_I8, 0.75, 14, torch.int8, 1),
             ("PEGReg504x1008 (alist) f32 NMS 0.75", "PEGReg504x1008.alist", q.RULE_NMS, q.DTYPE_F32, 0.75, mag, torch.float32, 4))
    results = []
    for name, fname, rule, dt, norm, m, tdt, width in cases:
        if args.only and args.only not in name:
            continue
        path = q.data_path(fname)
        if not os.path.exists(path):
            continue
        code = q.Code.from_alist(path) if fname.endswith(".alist") else q.Code.from_qc_file(path)
        # the layered schedule on a general H is parallel ACROSS frames (one thread per frame): give it a GPU-sized batch
        N, E, F = code.n, code.edges, (args.frames * 16 if fname.endswith(".alist") else args.frames)
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        Np = (N + 31) // 32 * 32
        x = torch.zeros((F, Np), dtype=torch.uint8, device=dev)
        x[:, :N] = torch.randint(0, 2, (F, N), dtype=torch.uint8, device=dev, generator=g)
        e = torch.zeros_like(x)
        e[:, :N] = (torch.rand((F, N), device=dev, generator=g) < args.qber).to(torch.uint8)

        def pack(b):
            v = (b.view(F, -1, 32).to(torch.int64) * w).sum(dim=-1)
            return torch.where(v >= 2**31, v - 2**32, v).to(torch.int32).contiguous()

        xb, yb = pack(x), pack(x ^ e)
        dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=rule, dtype=dt, max_iter=args.iters, early_stop=args.early_stop,
                        norm_factor=norm, offset=0.5, out_mode=q.OUT_ALL)
        syn = torch.empty((F, dec.syn_words), dtype=torch.int32, device=dev)
        dec.syndrome_device(xb.data_ptr(), F, syn.data_ptr(), st)
        llr = torch.empty((F, N), dtype=tdt, device=dev)
        dec.make_llr_device(yb.data_ptr(), 0, 0, float(m), 0.0, F, llr.data_ptr(), st)
        out = torch.empty((F, dec.cw_words), dtype=torch.int32, device=dev)
        ok = torch.empty(F, dtype=torch.uint8, device=dev)
        it = torch.empty(F, dtype=torch.int16, device=dev)
        dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
        ev[0].record()
        for r in range(args.reps):
            dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
            ev[r + 1].record()
        torch.cuda.synchronize()
        ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(args.reps))
        iters = float(it.float().mean())
        res = {"case": name, "kernel": dec.kernel_name, "frames": F, "N": N, "edges": E, "iterations": iters,
               "ok_frac": float(ok.float().mean()), "ms": ms, "codeword_mbps": F * N / ms / 1e3,
               "edge_updates_per_s": F * E * iters / ms * 1e3, "state_GBps": 4 * width * E * iters * F / ms / 1e6,
               "all_reconciled": bool((out[:, :N // 32] == xb[:, :N // 32]).all())}
        print(json.dumps(res), flush=True)
        results.append(res)
        dec.close()
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
