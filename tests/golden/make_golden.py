#!/usr/bin/env python3
"""Regenerates the committed fixtures from the read-only reference checkout.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Writes
  tests/golden/kat_pegreg504x1008.json   the reference's one known-answer vector
        ("main.cpp (alist)":443-462: data[504] -> encoded[1008], llrs[1008] -> decoded[504])
  tests/golden/matrix_facts.json         structural facts of the reference's matrices (SURVEY 8c pin 3)
  qcrypto-ldpc_b200/data/*.alist|*.qc    parity-check matrices the tests / bench load
        (data files of the reference: BOOT/matrices/H/*, ML/base_matrices/NR_*.txt re-emitted as .qc)
  qcrypto-ldpc_b200/data/wifi_n1944_r12.qc   SURVEY Appendix F table ([RECALL]; not from the reference)
"""
import json
import os
import re
import shutil

REF = "/root/reference/errorcorrection/ldpc_examples"
BOOT = REF + "/my_project_with_aff3ct/examples/bootstrap"
ML = REF + "/matlab_code_Base_matrices/matlab_code & Base_matrices"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DATA = os.path.join(ROOT, "qcrypto-ldpc_b200", "data")


def kat():
    src = open(BOOT + "/src/variants (copy out as main.cpp to use)/main.cpp (alist)").read()
    out = {}
    for name, typ in (("data", "int"), ("encoded", "int"), ("llrs", "float"), ("decoded", "int")):
        m = re.search(r"std::vector<%s>\s+%s\s*\{([^}]*)\}" % (typ, name), src)
        vals = [float(x) if typ == "float" else int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()]
        out[name] = vals
    assert len(out["data"]) == 504 and len(out["encoded"]) == 1008
    assert len(out["llrs"]) == 1008 and len(out["decoded"]) == 504
    out["source"] = "main.cpp (alist):443-462 (block comment), H = BOOT/matrices/H/PEGReg504x1008.alist"
    json.dump(out, open(os.path.join(HERE, "kat_pegreg504x1008.json"), "w"))


def nr_txt_to_qc(name, Z):
    rows = [[int(x) for x in l.split()] for l in open("%s/base_matrices/%s.txt" % (ML, name)) if l.strip()]
    with open(os.path.join(DATA, name + ".qc"), "w") as f:
        f.write("%d %d %d\n\n" % (len(rows[0]), len(rows), Z))
        for r in rows:
            f.write(" ".join(str(x) for x in r) + "\n")
    return rows


def matrices():
    os.makedirs(DATA, exist_ok=True)
    for n in ("PEGReg504x1008.alist", "20.alist", "1998.5.3.2665.alist", "NR_1_0_2.qc", "NR_1_1_192.qc",
              "NR_2_3_112.qc", "NR_1_7_30.qc", "test.qc", "test2.qc"):
        shutil.copyfile(BOOT + "/matrices/H/" + n, os.path.join(DATA, n))
    # the generator matrix the reference's KAT-E was made with ("main.cpp (alist)":333,443-455)
    shutil.copyfile(BOOT + "/matrices/G/PEGReg504x1008.alist", os.path.join(DATA, "G_PEGReg504x1008.alist"))
    facts = {}
    for name, Z in (("NR_1_1_384", 384), ("NR_1_1_24", 24), ("NR_2_6_52", 52), ("NR_1_7_240", 240), ("NR_1_0_256", 256)):
        rows = nr_txt_to_qc(name, Z)
        facts[name] = {"rows": len(rows), "cols": len(rows[0]), "edges": sum(x >= 0 for r in rows for x in r),
                       "row_degrees": [sum(x >= 0 for x in r) for r in rows]}
    # every BG1 file has 316 edges, every BG2 file 197 (SURVEY 8c pin 3)
    counts = {}
    for fn in sorted(os.listdir(ML + "/base_matrices")):
        rows = [[int(x) for x in l.split()] for l in open(ML + "/base_matrices/" + fn) if l.strip()]
        counts[fn] = [len(rows), len(rows[0]), sum(x >= 0 for r in rows for x in r)]
    facts["all_nr_files"] = counts
    json.dump(facts, open(os.path.join(HERE, "matrix_facts.json"), "w"), indent=0)
    # rate-1/2 N=1944 Z=81 ([RECALL] of the IEEE 802.11n table, not in the reference)
    wifi = """57 -1 -1 -1 50 -1 11 -1 50 -1 79 -1  1  0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
 3 -1 28 -1  0 -1 -1 -1 55  7 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1 -1
30 -1 -1 -1 24 37 -1 -1 56 14 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1
62 53 -1 -1 53 -1 -1  3 35 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1
40 -1 -1 20 66 -1 -1 22 28 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1
 0 -1 -1 -1  8 -1 42 -1 50 -1 -1  8 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1
69 79 79 -1 -1 -1 56 -1 52 -1 -1 -1  0 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1
65 -1 -1 -1 38 57 -1 -1 72 -1 27 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1
64 -1 -1 -1 14 52 -1 -1 30 -1 -1 32 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1
-1 45 -1 70  0 -1 -1 -1 77  9 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1
 2 56 -1 57 35 -1 -1 -1 -1 -1 12 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0
24 -1 61 -1 60 -1 -1 27 51 -1 -1 16  1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0
"""
    with open(os.path.join(DATA, "wifi_n1944_r12.qc"), "w") as f:
        f.write("24 12 81\n\n" + wifi)


if __name__ == "__main__":
    kat()
    matrices()
    print("golden fixtures written")
