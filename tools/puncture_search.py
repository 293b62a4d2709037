#!/usr/bin/env python3
"""Puncture-pattern search, the reference driver's loop batched on the GPU.

The reference (BOOT/src/main.cpp:305-409, strategy PARITY_BITS_SIMULATED_RANDOM) shuffles the parity positions, punctures the
first `bits_to_puncture` of them (LLR 0; the other parity bits are "sent" and get the confirmed-bit LLR), pushes
`iterations_per_BER` frames through encode -> BSC -> decode one at a time, and keeps the first pattern that shows no frame error.
Here one candidate pattern = ONE decoder call over `--frames` frames (the masks of qldpc_decode_bits are per call), and the
candidates are tried until one is error free (or all of `--patterns`, with `--all`).

    python tools/puncture_search.py [--code NR_1_1_384.qc] [--qber 0.02] [--puncture 3000] [--frames 2048] [--patterns 16]

`--ref-quirk` reproduces the driver's off-by-one (the confirmed-bit override starts at K+1, so parity
position K keeps its noisy channel LLR).  Prints one JSON line per pattern and a summary line; runs on the GPU box.
"""
import argparse
import importlib
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="NR_1_1_384.qc")
    ap.add_argument("--qber", type=float, default=0.02)
    ap.add_argument("--puncture", type=int, default=3000, help="parity bits to puncture (bits_to_puncture)")
    ap.add_argument("--frames", type=int, default=2048, help="frames per candidate (iterations_per_BER)")
    ap.add_argument("--patterns", type=int, default=16, help="candidates to try (simulation_iters)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--decoder", choices=("layered-i8", "flooding-spa"), default="layered-i8")
    ap.add_argument("--max-iter", type=int, default=20)
    ap.add_argument("--ref-quirk", action="store_true")
    ap.add_argument("--all", action="store_true", help="evaluate every candidate instead of stopping at the first clean one")
    args = ap.parse_args()

    q = importlib.import_module("qcrypto-ldpc_b200")
    code = q.Code.from_qc_file(q.data_path(args.code))
    N, K = code.n, code.k
    if args.decoder == "layered-i8":
        dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=args.max_iter,
                        early_stop=True, norm_factor=0.75, out_mode=q.OUT_INFO)
        # int8 scale 2^2: |LLR| = round(4 ln((1-q)/q)) capped to the 6-bit message range, confirmed bits saturate
        mag, known = float(min(31, round(4 * math.log((1 - args.qber) / args.qber)))), 31.0
    else:
        dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=args.max_iter,
                        early_stop=True, out_mode=q.OUT_INFO)
        mag, known = math.log((1 - args.qber) / args.qber), 23.02585      # BOOT/src/main.cpp: CONFIRMED_BIT_LLR
    rng = np.random.default_rng(args.seed)
    F = args.frames
    msg = rng.integers(0, 2, (F, K), dtype=np.uint8)
    msg_p = q.pack_bits(msg)
    cw_p = dec.encode_nr(msg_p)                                           # Alice: parity bits are what she transmits
    noise = np.zeros((F, N), np.uint8)
    noise[:, :K] = rng.random((F, K), dtype=np.float32) < args.qber       # BSC on the sifted key only
    bob_p = cw_p ^ q.pack_bits(noise)
    parity_pos = np.arange(K, N)
    known_bits = np.zeros(N, np.uint8)
    known_bits[(K + 1 if args.ref_quirk else K):] = 1

    found, rows = None, []
    t0 = time.perf_counter()
    for c in range(args.patterns):
        pat = rng.permutation(parity_pos)[:args.puncture]                 # std::shuffle + the first bits_to_puncture
        punct = np.zeros(N, np.uint8)
        punct[pat] = 1
        out, ok, iters = dec.decode_bits(bob_p, mag, known, known_mask=q.pack_bits(known_bits), punct_mask=q.pack_bits(punct))
        frame_err = int((out != msg_p).any(axis=1).sum())
        row = {"pattern": c, "punctured": int(args.puncture), "frames": F, "frame_errors": frame_err,
               "syndrome_fail": int((~ok).sum()), "mean_iters": float(iters.mean())}
        rows.append(row)
        print(json.dumps(row), flush=True)
        if frame_err == 0 and found is None:
            found = (c, np.sort(pat))
            if not args.all:
                break
    dt = time.perf_counter() - t0
    rate = (K) / (N - args.puncture)                                       # transmitted: N - K - punctured parity bits
    leak = (N - K - args.puncture) / K
    h = -args.qber * math.log2(args.qber) - (1 - args.qber) * math.log2(1 - args.qber)
    print(json.dumps({"code": args.code, "decoder": args.decoder, "qber": args.qber, "punctured": args.puncture,
                      "effective_rate": rate, "leak_per_key_bit": leak, "efficiency_f": leak / h,
                      "patterns_tried": len(rows), "clean_pattern": None if found is None else int(found[0]),
                      "pattern_head": None if found is None else [int(x) for x in found[1][:16]],
                      "seconds": dt, "frames_decoded": len(rows) * F,
                      "reference_loop": "BOOT/src/main.cpp:305-409 (one frame per decode_siho call)"}), flush=True)


if __name__ == "__main__":
    main()
