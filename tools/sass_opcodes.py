#!/usr/bin/env python3
"""SASS opcode counts of the built library (cuobjdump -sass): which instructions the product's kernels are made of.
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "minimp3_b200", "libminimp3_b200.so")
MARKERS = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "LDGSTS", "UTMALDG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(full):
    m = re.search(r"(k_\w+)(<[^>]*>)?", full)
    if not m:
        return full
    t = m.group(2) or ""
    t = t.replace("(bool)0", "false").replace("(bool)1", "true")
    return m.group(1) + t


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    names = demangle(list(per))
    print("# SASS opcode counts of `minimp3_b200/libminimp3_b200.so` (`cuobjdump -sass`, sm_100a cubins; tools/sass_opcodes.py)\n")
    print("Own kernels only (the CUB radix-sort kernels of the once-per-plan lane order are left out).  What to look for: `UBLKCP` = TMA bulk copy\n"
          "(`cp.async.bulk`), `SYNCS` = mbarrier, `FFMA2` / `FADD2` / `FMUL2` = Blackwell packed FP32; no `UTC*MMA` / `LDTM` / `STTM` in the product:\n"
          "the tensor path was tried as a stand-alone experiment (see DESIGN.md section 6), not adopted.\n")
    total = collections.Counter()
    for f, c in per.items():
        d = names.get(f, f)
        if "cub::" in d or "cub" in f[:12]:
            continue
        total.update(c)
        top = ", ".join("%s %d" % kv for kv in c.most_common(14))
        print("* `%s` -- %d instructions: %s" % (short(d), sum(c.values()), top))
    print("\nAll own kernels together: " + ", ".join("%s %d" % kv for kv in total.most_common(40)))
    print("\nMarkers: " + ", ".join("%s %d" % (m, total.get(m, 0)) for m in MARKERS))


if __name__ == "__main__":
    sys.exit(main())
