#!/usr/bin/env python
"""Aggregates `ncu -i rep --page source --csv --print-source cuda,sass` by CUDA source line:
share of warp-state samples and of executed instructions per line, with the dominant stall reasons.

    ncu -i march.ncu-rep --page source --csv --print-source cuda,sass > march_source.csv
    python tools/ncu_hot_lines.py march_source.csv [top_n] > profiles/march_hot_lines_r02.txt
"""
import csv
import collections
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path, newline="", errors="replace")))
    hdr = None
    for k, r in enumerate(rows):
        if any(c.strip() in ("Source", "# Samples", "Warp Stall Sampling (All Samples)", "Instructions Executed") for c in r):
            hdr = k
            break
    if hdr is None:
        print("no source table found in", path)
        return
    names = [c.strip() for c in rows[hdr]]
    col = {n: i for i, n in enumerate(names)}

    def pick(*cands):
        for c in cands:
            for n, i in col.items():
                if n.startswith(c):
                    return i
        return None
    c_src = pick("Source")
    c_smp = pick("Warp Stall Sampling (All", "# Samples", "Sampling Data (All")
    c_ins = pick("Instructions Executed", "# Instructions Executed")
    stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") or n.lower().startswith("warp stall") is False and n.startswith("Stall")]
    agg = collections.OrderedDict()
    tot_s = tot_i = 0.0
    cur = None
    for r in rows[hdr + 1:]:
        if len(r) <= max(x for x in (c_src, c_smp, c_ins) if x is not None):
            continue
        src = r[c_src].strip() if c_src is not None else ""

        def num(i):
            try:
                return float(r[i].replace(",", "")) if i is not None and r[i] != "" else 0.0
            except ValueError:
                return 0.0
        s, n = num(c_smp), num(c_ins)
        key = src[:110]
        a = agg.setdefault(key, [0.0, 0.0, collections.Counter()])
        a[0] += s
        a[1] += n
        for nme, i in stall_cols:
            v = num(i)
            if v:
                a[2][nme.replace("stall_", "")] += v
        tot_s += s
        tot_i += n
    print("total samples %d instr %d" % (tot_s, tot_i))
    for key, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        stalls = " ".join("%s=%d" % (k, v) for k, v in st.most_common(3))
        print("%5.2f%% smp %5.2f%% ins %s | %s" % (100 * s / max(tot_s, 1), 100 * n / max(tot_i, 1), key, stalls))


if __name__ == "__main__":
    main()
