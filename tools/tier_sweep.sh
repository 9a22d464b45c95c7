#!/bin/bash
# GPU box: sweep of the traversal tiers' hand-over thresholds on a bench workload.  bash tools/tier_sweep.sh <tag> <workload> <res> <spp> "<big list>" "<huge list>"
TAG=$1; WL=$2; RES=$3; SPP=$4
for b in $5; do for h in $6; do
  echo "== WT_BIG_TESTED=$b WT_HUGE_TESTED=$h" >> gpurun_out/${TAG}_sweep_${WL}.log
  WT_BIG_TESTED=$b WT_HUGE_TESTED=$h PROBE_REPS=2 timeout 300 python tools/probe.py $WL $RES $SPP 0 2>&1 | grep "^render 1" >> gpurun_out/${TAG}_sweep_${WL}.log
done; done
cat gpurun_out/${TAG}_sweep_${WL}.log
