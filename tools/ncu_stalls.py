#!/usr/bin/env python3
"""Per CUDA source line: which stall reasons the samples of an ncu --set full --import-source on capture fall into.

    python tools/ncu_stalls.py gpurun_out/prof_transform.ncu-rep [reason] [top_n]
reason: long_sb, short_sb, barrier, wait, math, mio, lg, ...  (default: all reasons, lines ranked by samples)
"""
import csv
import subprocess
import sys


def main(rep, reason=None, top=30):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True).stdout.decode("utf-8", "replace")
    rows = list(csv.reader(txt.splitlines()))
    agg, cur, i, cols = {}, None, 0, None
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            hdr = rows[i + 2]
            cols = {h[6:]: k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            isamp = hdr.index("# Samples")
            i += 3
            continue
        if cur and cols and len(r) > max(cols.values()) and r[0].strip().isdigit() and r[2] == "-":
            f = lambda x: int(float(x)) if x not in ("", "-") else 0
            a = agg.setdefault((cur, int(r[0])), [0, {}, r[1].strip()[:90]])
            a[0] += f(r[isamp])
            for name, k in cols.items():
                a[1][name] = a[1].get(name, 0) + f(r[k])
        i += 1
    tot = sum(v[0] for v in agg.values()) or 1
    totals = {}
    for v in agg.values():
        for k, n in v[1].items():
            totals[k] = totals.get(k, 0) + n
    print("samples %d; by reason: %s" % (tot, ", ".join("%s %.1f%%" % (k, 100*n/tot) for k, n in sorted(totals.items(), key=lambda kv: -kv[1]) if n)))
    key = (lambda kv: -kv[1][1].get(reason, 0)) if reason else (lambda kv: -kv[1][0])
    for k, v in sorted(agg.items(), key=key)[:top]:
        top3 = sorted(v[1].items(), key=lambda kv: -kv[1])[:3]
        print("%-18s %5d %5.1f%%  %-40s %s" % (k[0], k[1], 100*v[0]/tot, " ".join("%s:%d" % t for t in top3 if t[1]), v[2]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].isdigit() else None,
         int(sys.argv[-1]) if sys.argv[-1].isdigit() else 30)
