#!/usr/bin/env python3
"""Per CUDA source line: stall samples and executed instructions from an ncu report captured with
--import-source on (kernels built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof_entropy.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True).stdout.decode("utf-8", "replace")
    rows = list(csv.reader(txt.splitlines()))
    agg, cur, i = {}, None, 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            hdr = rows[i + 2]
            isamp, iex, ithr = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            i += 3
            continue
        if cur and len(r) > ithr and r[0].strip().isdigit() and r[2] == "-":
            f = lambda x: int(float(x)) if x not in ("", "-") else 0
            k = (cur, int(r[0]))
            a = agg.setdefault(k, [0, 0, 0, r[1].strip()[:100]])
            a[0] += f(r[isamp]); a[1] += f(r[iex]); a[2] += f(r[ithr])
        i += 1
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print("total stall samples %d, warp instructions %d" % (tot, toti))
    print("%-20s %5s %7s %7s %6s  %s" % ("file", "line", "samp%", "instr%", "thr/w", "source"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-20s %5d %6.1f%% %6.1f%% %6.1f  %s" % (k[0], k[1], 100*v[0]/tot, 100*v[1]/toti, v[2]/max(1, v[1]), v[3]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
