#!/usr/bin/env python3
"""Turn an .ncu-rep (one `ncu --set full` capture) into a small markdown summary for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_transform.ncu-rep transform > profiles/r01_transform.md
Also updates profiles/traffic.json[<name>] = dram bytes read + written per launch (GB) for the names bench.py reads
(transform, entropy, entropy_mips).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v*{"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main(rep, name):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, first = rows[0], rows[1], rows[2]
    print("# ncu --set full: %s (first captured launch)\n" % first[hdr.index("Kernel Name")])
    print("Captured with `ncu --set full --clock-control none --import-source on` on the bench command")
    print("(`python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline [--config ...]`, tools/profile.sh).\n")
    print("| metric | value | unit |\n|---|---|---|")
    vals = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals[k] = (first[i], units[i])
            print("| %s | %s | %s |" % (k, first[i], units[i]))
    stalls = [(h, first[i]) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    stalls = sorted(((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v.replace(",", "")))
                     for h, v in stalls), key=lambda x: -x[1])
    print("\nWarp stall reasons (warps stalled per issued instruction): " +
          ", ".join("%s %.2f" % s for s in stalls if s[1] >= 0.05))
    try:
        rd = to_bytes(*vals["dram__bytes_read.sum"])
        wr = to_bytes(*vals["dram__bytes_write.sum"])
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        t = json.load(open(tp)) if os.path.exists(tp) else {}
        t[name] = (rd + wr)/1e9
        t[name + "_note"] = "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum), from %s" % os.path.basename(rep)
        if name in ("transform", "entropy", "entropy_mips"):
            json.dump(t, open(tp, "w"), indent=1, sort_keys=True)
        print("\nDRAM traffic per launch: %.3f GB read + %.3f GB written = %.3f GB" % (rd/1e9, wr/1e9, (rd + wr)/1e9))
    except Exception as e:
        print("\n(no dram counters: %s)" % e)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_src_%s.csv" % name
    open(tmp, "w").write(src)
    print("\nPhases (SASS between consecutive BAR.SYNC; tools/ncu_phases.py):\n\n```")
    sys.stdout.flush()
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), tmp])
    print("```")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
