#!/usr/bin/env python3
"""BASELINE config 5 measurement: blind reconciliation over the ecd2 packet formats, two parties in loop-back
(host/driver_blind.cpp: BlindAlice / BlindBob of host/qldpc_blind.hpp, the handlers the ecd2 patch registers).
Runs on the GPU box; compiles the driver, writes synthetic stream-3 key blocks and prints a markdown table.

    python tools/blind_bench.py [--out gpurun_out/config5_blind.md]
"""
import argparse
import importlib
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HOST = os.path.join(ROOT, "qcrypto-ldpc_b200", "host")


def write_keys(q, path, n_blocks, workbits, qber, seed):
    rng = np.random.default_rng(seed)
    words = (workbits + 31) // 32
    A = rng.integers(0, 2, (n_blocks, words * 32)).astype(np.uint8)
    A[:, workbits:] = 0
    e = (rng.random((n_blocks, words * 32)) < qber).astype(np.uint8)
    e[:, workbits:] = 0
    pa, pb = q.pack_bits(A), q.pack_bits(A ^ e)
    with open(path, "wb") as f:
        f.write(struct.pack("<iif", n_blocks, workbits, qber))
        for b in range(n_blocks):
            f.write(pa[b].astype("<u4").tobytes())
            f.write(pb[b].astype("<u4").tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5_blind.md"))
    ap.add_argument("--code", default="NR_1_1_384.qc")
    args = ap.parse_args()
    q = importlib.import_module("qcrypto-ldpc_b200")
    tmp = tempfile.mkdtemp(prefix="blind")
    exe = os.path.join(tmp, "driver_blind")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", HOST, os.path.join(HOST, "driver_blind.cpp"), "-o", exe, q.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
    rows = []
    for n_blocks, qber in ((64, 0.03), (512, 0.01), (512, 0.03), (512, 0.05), (512, 0.08), (2048, 0.03)):
        keys = os.path.join(tmp, "k.bin")
        write_keys(q, keys, n_blocks, 65535, qber, seed=n_blocks + int(qber * 1000))
        best = None
        for _ in range(3):   # the first run of a process pays module load; keep the best wall time
            p = subprocess.run([exe, q.data_path(args.code), keys, os.path.join(tmp, "c.bin")], capture_output=True, text=True)
            assert p.returncode == 0, (p.returncode, p.stderr[-500:])
            d = json.loads(p.stdout.strip().splitlines()[-1])
            if best is None or d["wall_s"] < best["wall_s"]:
                best = d
        rows.append(best)
        print(json.dumps(best), flush=True)
    lines = ["# Config 5 — ecd2-integrated blind reconciliation, two parties in loop-back (B200, `host/driver_blind.cpp`), round 2",
             "",
             "`tools/blind_bench.py`: blocks of 65 535 bits (ecd2's block cap, `processblock_mgmt.c:94`) = 8 frames of K = 8 448 key bits,",
             "defaults f_start 1.25, 2 extra parity block rows per NACK round, int8 layered NMS 6/8, max 20 iterations, early termination;",
             "all pending blocks are decoded in one launch per rate. Wall time covers the whole exchange (NR encoding, packet building and",
             "validation on both sides, host<->device copies, decoding, NACK rounds, CRC confirmation); best of three runs.",
             "Leakage counts every parity bit sent, every revealed bit and the 32 CRC bits per frame, on both sides identically.",
             "",
             "| blocks | QBER | initial rows | rounds histogram | leak bits (Alice = Bob) | efficiency f | packets A→B / B→A | bytes A→B / B→A | wall s | reconciled key bit/s |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    for d in rows:
        assert d["blocks_differ"] == 0 and d["leak_bits"] == d["leak_bits_bob"] and d["crc_mismatches"] == 0
        lines.append("| %d | %g %% | %d | %s | %d | %.3f | %d / %d | %d / %d | %.4f | %.3g |" % (
            d["blocks"], 100 * d["qber"], d["initial_rows"], json.dumps(d["round_hist"]).replace('"', ''), d["leak_bits"],
            d["efficiency"], d["packets_a2b"], d["packets_b2a"], d["bytes_a2b"], d["bytes_b2a"], d["wall_s"],
            d["reconciled_key_bits_per_s"]))
    lines += ["",
              "Every block ends identical on both sides. Reference comparator (SURVEY §6, probed): Cascade + BICONF in `ecd2` reconciles one",
              "40 000-bit block at QBER 3 % in 8.2 s with 293 packets each way = 4.9e3 key bit/s. The same handlers inside two patched `ecd2`",
              "daemons joined by FIFOs: `tests/test_gpu_ecd2.py`."]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write("\n".join(lines) + "\n")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
