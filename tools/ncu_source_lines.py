#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one kernel of an .ncu-rep (captured with --import-source on):
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv;  python tools/ncu_source_lines.py src.csv [N]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr = rows[2]
    i_sass, i_inst, i_samp = 3, hdr.index("Instructions Executed"), hdr.index("# Samples")
    cur, agg, src = None, {}, {}
    for r in rows[3:]:
        if len(r) <= i_inst:
            continue
        if r[0] != "":
            if r[0].isdigit():
                cur = int(r[0]); src[cur] = r[1]
            continue
        if cur is None:
            continue
        a = agg.setdefault(cur, [0, 0, {}])
        a[0] += num(r[i_inst]); a[1] += num(r[i_samp])
        w = r[i_sass].split()
        op = (w[1] if w[0].startswith('@') else w[0]).split('.')[0]
        a[2][op] = a[2].get(op, 0) + num(r[i_inst])
    tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
    print("warp instructions", tot, "samples", tots)
    for ln, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        ops = sorted(a[2].items(), key=lambda x: -x[1])[:6]
        print("%4d %5.1f%% inst %5.1f%% samp  %-72s %s" % (ln, 100 * a[0] / tot, 100 * a[1] / max(tots, 1), src[ln].strip()[:72],
                                                        " ".join("%s:%.1f" % (o, 100 * c / tot) for o, c in ops)))


if __name__ == "__main__":
    main()
