#!/usr/bin/env python3
"""Generates qcrypto-ldpc_b200/data/qkd_irregular_n65536_r34.qc: the IRREGULAR long-block QKD code of BASELINE config 3.

    python tests/golden/make_qkd65536_irregular.py            # needs /root/reference (this container only)

Construction = the reference's PSD-PEG (errorcorrection/ldpc_examples/psd-peg.py:186-455) under a seeded `random`.  The
script builds regular codes as shipped; its author marks the one line that fixes the degrees ("For now var node degrees
are constant, but can be easily customized", :12, and the commented `vn_deg.append(random.randint(3, 10))`, :226): that
line is the ONLY thing replaced here, at load time, by a lookup into the degree list below.  Nothing of the script is
copied into this repository.

Code: base matrix 16 x 64, lifting size Z = 1024 -> N = 65 536 key bits per frame, M = 16 384 syndrome bits (rate 3/4:
at QBER 3 % the unpunctured code already sits at efficiency f = 0.25 / h(0.03) = 1.286; lower QBERs are reached by
puncturing, higher ones by shortening, tools/qber_sweep.py).  Variable degrees (ascending, the PEG order): 12 x 2, 36 x 3,
10 x 7, 6 x 14 -- average 4.47, check degrees 17 / 18.  Chosen by density evolution on the BSC (population dynamics, 2e5
messages): BP threshold 3.66 % against 2.84 % for the regular (3, 12) code of the same rate (Shannon limit 4.17 %)."""
import io
import os
import random
import sys
from contextlib import redirect_stdout

REF = "/root/reference/errorcorrection/ldpc_examples/psd-peg.py"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
COLS, ROWS, Z = 64, 16, 1024
# psd-peg.py takes "Gamma == -1" as its root marker (:347); an accumulated shift that happens to equal -1 sends it down the
# wrong branch and it indexes p[] with a variable id (IndexError).  Seeds are tried in order until a run gets through; the
# seed that did is printed.
SEEDS = range(20261019, 20261119)
DEGREES = [2] * 12 + [3] * 36 + [7] * 10 + [14] * 6
FIXED_DEGREE_LINE = "    vn_deg.append(vn_degree_param)\n"


def main():
    assert len(DEGREES) == COLS
    src = open(REF).read()
    assert src.count(FIXED_DEGREE_LINE) == 1, "psd-peg.py changed: the degree line is not where it was"
    src = src.replace(FIXED_DEGREE_LINE, "    vn_deg.append(QLDPC_DEGREES[j])\n")
    argv = sys.argv
    sys.argv = [REF, str(COLS), str(ROWS), "0", str(Z), "0", "1"]
    code = compile(src, REF, "exec")
    seed_used = None
    try:
        for seed in SEEDS:
            random.seed(seed)
            buf = io.StringIO()
            try:
                with redirect_stdout(buf):
                    exec(code, {"__name__": "__main__", "QLDPC_DEGREES": DEGREES})
                seed_used = seed
                break
            except IndexError:
                continue
    finally:
        sys.argv = argv
    assert seed_used is not None
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    hdr = lines[0].split()
    assert int(hdr[1]) == ROWS and int(hdr[2]) == Z, hdr
    rows = [[int(x) for x in l.split()] for l in lines[1:1 + ROWS]]
    a = [[(s % Z) if s >= 0 else -1 for s in r[:COLS]] for r in rows]
    assert [sum(1 for r in a if r[c] >= 0) for c in range(COLS)] == DEGREES
    rdeg = [sum(1 for s in r if s >= 0) for r in a]
    out = os.path.join(ROOT, "qcrypto-ldpc_b200", "data", "qkd_irregular_n65536_r34.qc")
    with open(out, "w") as f:
        f.write("%d %d %d\n\n" % (COLS, ROWS, Z))
        for r in a:
            f.write(" ".join("%d" % s for s in r) + "\n")
    print("wrote", out, "seed", seed_used, "edges per lane:", sum(rdeg), "check degrees:", sorted(set(rdeg)))


if __name__ == "__main__":
    main()
