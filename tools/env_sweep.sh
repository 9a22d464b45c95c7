#!/bin/bash
# GPU box: A/B of environment-tunable settings on a bench workload.  bash tools/env_sweep.sh <tag> <workload> <res> <spp> <pool> "VAR=a VAR=b ..." (each entry one run; "A=1,B=2" sets several)
TAG=$1; WL=$2; RES=$3; SPP=$4; POOL=$5
for cfg in $6; do
  echo "== $cfg" >> gpurun_out/${TAG}_envsweep_${WL}.log
  env $(echo $cfg | tr ',' ' ') PROBE_REPS=2 timeout 300 python tools/probe.py $WL $RES $SPP 0 $POOL 2>&1 | grep "^render 1" | cut -c1-80 >> gpurun_out/${TAG}_envsweep_${WL}.log
done
cat gpurun_out/${TAG}_envsweep_${WL}.log
