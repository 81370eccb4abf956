#!/bin/bash
# Run under gpurun on one B200: plain run first (must exit 0), then the ncu launch list and one
# --set full capture per kernel of the same command.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_l3_transform -s 3 -c 2 -f -o gpurun_out/prof_transform $CMD > gpurun_out/ncu_transform.log 2>&1
echo "transform capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_l3_entropy -s 3 -c 2 -f -o gpurun_out/prof_entropy $CMD > gpurun_out/ncu_entropy.log 2>&1
echo "entropy capture rc=$?"
ls -la gpurun_out
