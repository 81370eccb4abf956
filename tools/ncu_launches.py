#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, average, share.

    python tools/ncu_launches.py profiles/r01_launches_bench_steps2.csv > profiles/r01_launches_summary.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        ms = float(r[iv].replace(",", ""))*{"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[iu], 1e-6)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print("# ncu launch list summary\n")
    print("Command: `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline` under")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv` (raw list: %s)." % path.split("/")[-1])
    print("Per-launch times are cold-cache and serialised; compare shares, not absolutes.\n")
    print("| kernel | launches | avg ms | share of GPU time |\n|---|---|---|---|")
    for k, a in agg.items():
        print("| `%s` | %d | %.3f | %.1f %% |" % (k, a[0], a[1]/a[0], 100*a[1]/tot))


if __name__ == "__main__":
    main(sys.argv[1])
