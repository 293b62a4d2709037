#!/usr/bin/env python3
"""Generates qcrypto-ldpc_b200/data/qkd_psdpeg_n65536.qc with the REFERENCE's own PSD-PEG construction
(errorcorrection/ldpc_examples/psd-peg.py, run unmodified under a seeded `random`): base matrix 16 x 32,
variable degree 3, lifting size Z = 2048 -> N = 65536 key bits per frame, M = 32768 syndrome bits (rate 1/2).

    python tests/golden/make_qkd65536.py            # needs /root/reference (this container only)

psd-peg.py prints [A | I] (the identity block is the systematic parity part the AFF3CT from-QC encoder wants,
SURVEY.md A.2); the file written here keeps A only: in syndrome decoding Alice sends s = A x_A.  Shifts are
reduced mod Z on load (psd-peg.py can print shifts >= Z)."""
import io
import os
import random
import runpy
import sys
from contextlib import redirect_stdout

REF = "/root/reference/errorcorrection/ldpc_examples/psd-peg.py"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
COLS, ROWS, DV, Z, SEED = 32, 16, 3, 2048, 20261018


def main():
    random.seed(SEED)
    argv = sys.argv
    sys.argv = [REF, str(COLS), str(ROWS), str(DV), str(Z), "0", "1"]
    buf = io.StringIO()
    try:
        with redirect_stdout(buf):
            runpy.run_path(REF, run_name="__main__")
    finally:
        sys.argv = argv
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    hdr = lines[0].split()
    assert int(hdr[1]) == ROWS and int(hdr[2]) == Z, hdr
    rows = [[int(x) for x in l.split()] for l in lines[1:1 + ROWS]]
    assert all(len(r) == int(hdr[0]) for r in rows)
    a = [[(s % Z) if s >= 0 else -1 for s in r[:COLS]] for r in rows]
    ident = [r[COLS:] for r in rows]
    assert all(ident[i][j] == (0 if i == j else -1) for i in range(ROWS) for j in range(ROWS)), "expected [A | I]"
    assert all(sum(1 for r in a if r[c] >= 0) == DV for c in range(COLS))
    out = os.path.join(ROOT, "qcrypto-ldpc_b200", "data", "qkd_psdpeg_n65536.qc")
    with open(out, "w") as f:
        f.write("%d %d %d\n\n" % (COLS, ROWS, Z))
        for r in a:
            f.write(" ".join("%d" % s for s in r) + "\n")
    print("wrote", out, "edges per lane:", sum(1 for r in a for s in r if s >= 0))


if __name__ == "__main__":
    main()
