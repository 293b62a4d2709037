#!/usr/bin/env python3
"""BASELINE config 3 measurement: QBER sweep 1..8 % on a long-block QKD code (N = 65536), float flooding SPA vs normalised
min-sum, rate adaptation by puncturing (p) / shortening (s) a public pseudo-random position set.
Runs on the GPU box; writes a markdown table (FER, mean iterations, efficiency f = leak / h(QBER), decode Mbit/s of key).

    python tools/qber_sweep.py [--code qkd_irregular_n65536_r34.qc] [--frames 592] [--f-targets 1.1,1.15,...] [--out gpurun_out/qber_sweep.md]

Without --f-targets the fixed (p, s) grid of round 1 is used (rate-1/2 codes); with it, every QBER is decoded at the
puncturing / shortening fraction that lands on each target efficiency: leak per key bit = (M - p) / (N - p - s) = f h(QBER)
-- puncturing when that is below M / N, shortening above.
"""
import argparse
import importlib
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def h2(p):
    return -p * math.log2(p) - (1 - p) * math.log2(1 - p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=192)
    ap.add_argument("--max-iter", type=int, default=50)
    ap.add_argument("--code", default="qkd_psdpeg_n65536.qc")
    ap.add_argument("--f-targets", default="")
    ap.add_argument("--fast-spa", action="store_true", help="QLDPC_FLAG_FAST_SPA for the SPA column")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "qber_sweep.md"))
    args = ap.parse_args()
    import torch
    q = importlib.import_module("qcrypto-ldpc_b200")
    dev = torch.device("cuda", 0)
    code = q.Code.from_qc_file(q.data_path(args.code))
    N, M = code.n, code.m
    F = args.frames
    decs = {"SPA": q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=args.max_iter,
                             early_stop=True, out_mode=q.OUT_ALL, flags=q.FLAG_FAST_SPA if args.fast_spa else 0),
            "NMS 13/16": q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_NMS, dtype=q.DTYPE_F32, max_iter=args.max_iter,
                                   early_stop=True, norm_factor=0.8125, out_mode=q.OUT_ALL)}
    any_dec = decs["SPA"]
    cw, sw = any_dec.cw_words, any_dec.syn_words
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=dev); g.manual_seed(99)
    perm = torch.from_numpy(np.random.default_rng(7).permutation(N)).to(dev)
    weights = (2 ** torch.arange(31, -1, -1, device=dev, dtype=torch.int64))

    def pack(bits):   # [.., N] uint8 on device -> MSB-first int32 words
        v = (bits.view(*bits.shape[:-1], -1, 32).to(torch.int64) * weights).sum(dim=-1)
        return torch.where(v >= 2**31, v - 2**32, v).to(torch.int32).contiguous()

    rows = []
    f_targets = [float(x) for x in args.f_targets.split(",") if x]

    def grid(qber):
        if not f_targets:
            return ((0.0, 0.0), (0.05, 0.0), (0.10, 0.0), (0.15, 0.0), (0.20, 0.0), (0.25, 0.0), (0.0, 0.05), (0.10, 0.05))
        pts = []
        for ft in f_targets:
            leak = ft * h2(qber)
            if leak >= 1.0:
                continue
            if leak < M / N:
                pts.append(((M - leak * N) / (1 - leak) / N, 0.0))      # puncture
            else:
                pts.append((0.0, (N - M / leak) / N))                    # shorten
        return pts

    for qber in (0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08):
        for pf, sf in grid(qber):
            n_p, n_s = int(pf * N), int(sf * N)
            punct = torch.zeros(N, dtype=torch.uint8, device=dev); punct[perm[:n_p]] = 1
            short = torch.zeros(N, dtype=torch.uint8, device=dev); short[perm[n_p:n_p + n_s]] = 1
            x = torch.randint(0, 2, (F, N), dtype=torch.uint8, device=dev, generator=g)
            e = (torch.rand((F, N), device=dev, generator=g) < qber).to(torch.uint8)
            e[:, (punct | short) == 1] = 0
            xb, yb = pack(x), pack(x ^ e)
            syn = torch.empty((F, sw), dtype=torch.int32, device=dev)
            any_dec.syndrome_device(xb.data_ptr(), F, syn.data_ptr(), st)
            pm, km = pack(punct), pack(short)
            llr = torch.empty((F, N), dtype=torch.float32, device=dev)
            mag = math.log((1 - qber) / qber)
            any_dec.make_llr_device(yb.data_ptr(), km.data_ptr(), pm.data_ptr(), mag, 23.02585, F, llr.data_ptr(), st)
            n_key = N - n_p - n_s
            leak = (M - n_p) / n_key            # syndrome bits minus the filler bits they are spent on, per key bit
            eff = leak / h2(qber)
            cell = [("%.0f %%" % (100 * qber)), "%.3f" % pf, "%.3f" % sf, "%.3f" % leak, "%.3f" % eff]
            for name, dec in decs.items():
                out = torch.empty((F, cw), dtype=torch.int32, device=dev)
                ok = torch.empty(F, dtype=torch.uint8, device=dev)
                it = torch.empty(F, dtype=torch.int16, device=dev)
                dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                good = ok.bool() & (out == xb).all(dim=1)      # converged AND equal to Alice's word
                fer = 1.0 - float(good.float().mean())
                undet = int((ok.bool() & ~(out == xb).all(dim=1)).sum())
                cell += ["%.3f" % fer, "%.1f" % float(it.float().mean()), "%.0f" % (F * n_key / dt / 1e6), str(undet)]
            rows.append(cell)
            print(" | ".join(cell), flush=True)
    hdr = ["QBER", "p/N", "s/N", "leak/key bit", "f"]
    for name in decs:
        hdr += ["FER %s" % name, "iters", "key Mbit/s", "undetected"]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("# QBER sweep, %s (N = %d, M = %d, check degree <= %d), flooding fp32 (%s), max %d iterations, %d frames per cell, B200\n\n" %
                (args.code, N, M, code.max_chk_degree, "SPA on the SFUs, QLDPC_FLAG_FAST_SPA" if args.fast_spa else "SPA in double", args.max_iter, F))
        f.write("Syndrome decoding with a public pseudo-random modulation pattern: p punctured (filler, LLR 0), s shortened (known, LLR 23.03).\n")
        f.write("FER counts frames that did not converge to Alice's word; `undetected` = converged to a different word.\n\n")
        f.write("| " + " | ".join(hdr) + " |\n|" + "---|" * len(hdr) + "\n")
        for r in rows:
            f.write("| " + " | ".join(r) + " |\n")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
