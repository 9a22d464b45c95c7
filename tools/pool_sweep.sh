#!/bin/bash
# GPU box: sweep of pool size x number of sub-pools on a bench workload.  bash tools/pool_sweep.sh <tag> <workload> <res> <spp> "<pools>" "<subpools>"
TAG=$1; WL=$2; RES=$3; SPP=$4
for p in $5; do for n in $6; do
  echo "== pool=$p WT_SUBPOOLS=$n" >> gpurun_out/${TAG}_poolsweep_${WL}.log
  WT_SUBPOOLS=$n PROBE_REPS=2 timeout 300 python tools/probe.py $WL $RES $SPP 0 $p 2>&1 | grep "^render 1" >> gpurun_out/${TAG}_poolsweep_${WL}.log
done; done
cat gpurun_out/${TAG}_poolsweep_${WL}.log
