#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall samples / executed instructions / shared-memory wavefronts
between consecutive BAR.SYNC instructions (= kernel phases), first kernel instance only."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    icf, iw = hdr.index('L1 Wavefronts Shared Excessive'), hdr.index('L1 Wavefronts Shared')
    acc = [[0, 0, 0, 0, 0]]
    f = lambda x: int(float(x)) if x not in ('', '-') else 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            break          # next kernel instance
        if r[isamp] == '# Samples':
            break
        acc[-1][0] += f(r[isamp]); acc[-1][1] += f(r[iex]); acc[-1][2] += 1; acc[-1][3] += f(r[icf]); acc[-1][4] += f(r[iw])
        if 'BAR.SYNC' in r[ia]:
            acc.append([0, 0, 0, 0, 0])
    tot = sum(a[0] for a in acc) or 1
    toti = sum(a[1] for a in acc) or 1
    print("segment  stall-samples  instr-executed  sass  smem-wavefronts  excess-wavefronts")
    for i, a in enumerate(acc):
        print("%3d   %5.1f%%   %5.1f%% (%d)   %5d   %d   %d" % (i, 100*a[0]/tot, 100*a[1]/toti, a[1], a[2], a[4], a[3]))


if __name__ == "__main__":
    main(sys.argv[1])
