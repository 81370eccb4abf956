#!/bin/bash
# every BASELINE.json config through bench.py on one GPU: one JSON line per config -> gpurun_out/bench_configs.jsonl
# usage (on the GPU box): bash tools/bench_configs.sh [steps] [configs...]
mkdir -p gpurun_out
steps=${1:-20}; shift
cfgs=${@:-2 2b 3 4 5 mips}
out=gpurun_out/bench_configs.jsonl
: > $out
for c in $cfgs; do
  echo "== config $c" >&2
  python bench.py --config $c --steps $steps --warmup 5 --e2e-steps 3 2>gpurun_out/bench_cfg_$c.err | tail -1 >> $out || echo "{\"config_id\": \"$c\", \"failed\": true}" >> $out
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_configs.jsonl"):
    try: d = json.loads(l)
    except Exception: print("unparsed:", l[:200]); continue
    if d.get("failed"): print(d); continue
    print("cfg %-4s value %7.1f G  %.3f ms/step  kern %s  roof %.3f (%s)  e2e %s  scope_ii %s  malloc %s" % (
        d["config"]["config_id"], d["value"]/1e9, d["ms_per_step"], {k: round(v, 3) for k, v in d["kernels_ms"].items()},
        d["roofline"]["frac"], d["roofline"]["kernel"],
        d.get("e2e") and round(d["e2e"]["value"]/1e9, 1), d.get("scope_ii") and round(d["scope_ii"]["value"]/1e9, 1),
        d.get("e2e_malloc") and round(d["e2e_malloc"]["value"]/1e9, 1)))
PY
