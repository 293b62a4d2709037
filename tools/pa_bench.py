#!/usr/bin/env python3
"""Privacy amplification / confirmation CRC through the host-pointer entry points (run on the GPU box):
wall time per call of qldpc_privacy_amplify (65 535-bit blocks -> 40 000 final bits, priv_amp.c:213-218) and of
qldpc_crc32_frames, copies inside the call.

    python tools/pa_bench.py [--out gpurun_out/postproc.md]"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "postproc.md"))
    args = ap.parse_args()
    q = importlib.import_module("qcrypto-ldpc_b200")
    rng = np.random.default_rng(0)
    wb, fb = 65535, 40000
    words = (wb + 31) // 32
    lines = ["# Privacy amplification and confirmation CRC on the GPU, host pointers, copies inside the call (B200, round 2)", "",
             "`tools/pa_bench.py`: 65 535-bit blocks (ecd2's cap) -> 40 000 final bits each, one launch per call, bit-exact with",
             "`priv_amp.c:213-218` / `rnd.c:118-127` (`tests/golden/pa_golden.json`); best of 5 calls after one warm-up call.", "",
             "| call | blocks / frames per call | wall ms | output Mbit/s | input Mbit/s |", "|---|---|---|---|---|"]
    for nb in (1, 8, 64, 512):
        keys = rng.integers(0, 2**32, (nb, words), dtype=np.uint64).astype(np.uint32)
        seeds = rng.integers(1, 2**32, nb, dtype=np.uint64).astype(np.uint32)
        best = 1e9
        for rep in range(6):
            t0 = time.perf_counter()
            out = q.privacy_amplify(keys, np.full(nb, wb, np.int32), np.full(nb, fb, np.int32), seeds)
            dt = time.perf_counter() - t0
            if rep:
                best = min(best, dt)
        lines.append("| `qldpc_privacy_amplify` | %d | %.3f | %.1f | %.1f |" % (nb, best * 1e3, nb * fb / best / 1e6, nb * wb / best / 1e6))
        print(lines[-1], flush=True)
    for nf in (8, 512, 4096, 65536):
        bits = rng.integers(0, 2**32, (nf, 264), dtype=np.uint64).astype(np.uint32)
        best = 1e9
        for rep in range(6):
            t0 = time.perf_counter()
            crc = q.crc32_frames(bits)
            dt = time.perf_counter() - t0
            if rep:
                best = min(best, dt)
        lines.append("| `qldpc_crc32_frames` (K = 8 448 bits) | %d | %.3f | %.1f | %.1f |" % (nf, best * 1e3, nf * 32 / best / 1e6, nf * 8448 / best / 1e6))
        print(lines[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
