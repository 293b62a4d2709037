#!/usr/bin/env python3
"""Differential soak of the three int8 layered implementations on the GPU (run on the GPU box): the streamed headline kernel
(layered_i8s.cu), the previous-generation kernel (layered_i8.cu, QLDPC_FLAG_LI8_RESIDENT) and the generic layered kernel
(layered_generic.cu, taken when posteriors are requested) must agree on decoded bits, syndrome flag and iteration count for
every frame -- random and saturating LLRs, with and without syndrome, both rules, several QBERs.  Prints one line per case.

    python tools/soak_i8.py [--frames 32768] [--code NR_1_1_384.qc]"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32768)
    ap.add_argument("--code", default="NR_1_1_384.qc")
    args = ap.parse_args()
    q = importlib.import_module("qcrypto-ldpc_b200")
    code = q.Code.from_qc_file(q.data_path(args.code))
    N, K, F = code.n, code.k, args.frames
    bad = 0
    case = 0
    for rule, kw in ((q.RULE_NMS, dict(norm_factor=0.75)), (q.RULE_OMS, dict(offset=2.0)), (q.RULE_NMS, dict(norm_factor=0.875))):
        for early in (True, False):
            for kind in ("bsc3", "bsc8", "random", "saturating"):
                case += 1
                rng = np.random.default_rng(1000 + case)
                mk = lambda **extra: q.Decoder(code, schedule=q.SCHED_LAYERED, rule=rule, dtype=q.DTYPE_I8, max_iter=8,
                                               early_stop=early, out_mode=q.OUT_ALL, **kw, **extra)
                a, b = mk(), mk(flags=q.FLAG_LI8_RESIDENT)
                if kind.startswith("bsc"):
                    qber = 0.03 if kind == "bsc3" else 0.08
                    x = rng.integers(0, 2, (F, N), dtype=np.uint8)
                    y = x ^ (rng.random((F, N), dtype=np.float32) < qber)
                    syn = a.syndrome(q.pack_bits(x))
                    llr = np.where(y, -14, 14).astype(np.int8)
                elif kind == "random":
                    llr = rng.integers(-40, 41, (F, N)).astype(np.int8)
                    syn = None
                else:
                    llr = rng.choice(np.array([-128, -127, -64, -1, 0, 1, 63, 127], np.int8), (F, N))
                    syn = rng.integers(0, 2**32, (F, a.syn_words), dtype=np.uint64).astype(np.uint32)
                ra = a.decode(llr, syn)
                rb = b.decode(llr, syn)
                n_gen = min(F, 4096)                       # the generic kernel is the slow one: a subset
                rc = a.decode(llr[:n_gen], None if syn is None else syn[:n_gen], want_posterior=True)
                same_ab = all((u == v).all() for u, v in zip(ra[:3], rb[:3]))
                same_ac = all((u[:n_gen] == v).all() for u, v in zip(ra[:3], rc[:3]))
                bad += (not same_ab) + (not same_ac)
                print("case %2d rule %d early %d %-10s kernels %s / %s: i8s==i8 %s, i8s==generic %s, ok %.3f, mean iters %.2f" %
                      (case, rule, early, kind, a.kernel_name, b.kernel_name, same_ab, same_ac, ra[1].mean(), ra[2].mean()), flush=True)
                a.close(); b.close()
    print("MISMATCHING CASES: %d" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
