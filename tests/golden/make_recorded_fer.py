#!/usr/bin/env python3
"""Extracts the FER / BER tables the reference RECORDED for its own LDPC decoders into tests/golden/recorded_fer.json.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_recorded_fer.py

Sources (paths relative to /root/reference/errorcorrection):
  ldpc_examples/matlab_code_Base_matrices/matlab_code & Base_matrices/sim_results.m:1-13
      res11 / res12: BPSK_nrldpc_sim_RM_FP.m (integer layered offset min-sum, rmax 3, 20 iterations, rate 1/2) on
      NR_1_1_24 and NR_2_6_52; columns EbNodB FER BER Nblkerrs Nbiterrs Nblocks (the script's `disp` line, :109)
  README_LDPC.md:937-974
      "5G Standard (NR_1_0_2.qc ...)": the (5g-qc) driver, flooding SPA, 10 iterations, 100 frames per QBER step,
      parity bits punctured from the end to a target efficiency of 1.0 (expansion_factor argument 1)
"""
import json
import os
import re

EC = "/root/reference/errorcorrection"
ML = EC + "/ldpc_examples/matlab_code_Base_matrices/matlab_code & Base_matrices"
HERE = os.path.dirname(os.path.abspath(__file__))


def matlab_table(src, name):
    m = re.search(r"%s\s*=\s*\[(.*?)\];" % name, src, re.S)
    rows = []
    for line in m.group(1).split(";"):
        v = line.split()
        if len(v) == 6:
            rows.append({"ebno_db": float(v[0]), "fer": float(v[1]), "ber": float(v[2]), "frame_errors": int(v[3]),
                         "bit_errors": int(v[4]), "frames": int(v[5])})
    return rows


def readme_table(lines):
    rows = []
    punct = None
    for ln in lines:
        m = re.search(r"Punct\. Bits \(Round down\): (\d+)\| Reconcil\. Effic\.: ([0-9.e+-]+)\|.*Code rate: ([0-9.]+)", ln)
        if m:
            punct = (int(m.group(1)), float(m.group(2)), float(m.group(3)))
            continue
        m = re.match(r"-\s+([0-9.]+) \|\|\s+(\d+) \|\s+(\d+) \|\s*(\d+) \|\s*([0-9.e+-]+) \|\s*([0-9.e+-]+) \|\|", ln)
        if m and punct:
            rows.append({"qber": float(m.group(1)), "frames": int(m.group(2)), "bit_errors": int(m.group(3)),
                         "frame_errors": int(m.group(4)), "ber": float(m.group(5)), "fer": float(m.group(6)),
                         "punctured_bits": punct[0], "efficiency": punct[1], "code_rate": punct[2]})
            punct = None
    return rows


def main():
    src = open(ML + "/sim_results.m").read()
    readme = open(EC + "/README_LDPC.md").read().split("\n")
    start = next(i for i, l in enumerate(readme) if l.startswith("5G Standard (NR_1_0_2.qc"))
    out = {
        "sim_results_m": {
            "source": "ML/sim_results.m:1-13 (BPSK_nrldpc_sim_RM_FP.m, Rate = 1/2, rmax = 3, MaxItrs = 20)",
            "NR_1_1_24": matlab_table(src, "res11"),
            "NR_2_6_52": matlab_table(src, "res12"),
        },
        "readme_ldpc_nr_1_0_2": {
            "source": "errorcorrection/README_LDPC.md:%d-%d ((5g-qc) driver, flooding SPA, n_ite 10, 100 frames per QBER)"
                      % (start + 1, start + 38),
            "rows": readme_table(readme[start:start + 40]),
        },
    }
    assert len(out["sim_results_m"]["NR_1_1_24"]) == 4 and len(out["sim_results_m"]["NR_2_6_52"]) == 4
    assert len(out["readme_ldpc_nr_1_0_2"]["rows"]) == 12
    json.dump(out, open(os.path.join(HERE, "recorded_fer.json"), "w"), indent=1)
    print("wrote recorded_fer.json")


if __name__ == "__main__":
    main()
