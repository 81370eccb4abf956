#!/bin/bash
# Run under gpurun on one B200: plain run first (must exit 0), then the ncu launch list and one
# --set full capture per kernel of the same command.  Outputs land in gpurun_out/.
#   bash tools/profile.sh [transform|entropy|launches|all|kernel:<name regex>] [bench args...]
set -u
mkdir -p gpurun_out
what=${1:-all}; shift || true
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-400
if [ "$what" = launches ] || [ "$what" = all ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
fi
if [ "$what" = transform ] || [ "$what" = all ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_l3_transform -s 3 -c 1 -f -o gpurun_out/prof_transform $CMD > gpurun_out/ncu_transform.log 2>&1
  echo "transform capture rc=$?"
fi
if [ "$what" = entropy ] || [ "$what" = all ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_l3_entropy -s 3 -c 1 -f -o gpurun_out/prof_entropy $CMD > gpurun_out/ncu_entropy.log 2>&1
  echo "entropy capture rc=$?"
fi
case "$what" in kernel:*)
  k=${what#kernel:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k capture rc=$?";;
esac
ls -la gpurun_out | tail -8
