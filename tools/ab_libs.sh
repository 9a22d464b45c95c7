#!/bin/bash
# GPU box: A/B of two builds of the library on the four workloads (second render of tools/probe.py).  bash tools/ab_libs.sh <tag> <variant .so>
TAG=$1; VAR=$2
for cfg in "cornell 1440 1 0 1048576" "etoile 720 16 0" "sponza 1920 1 0" "double_slits 1440 16 0"; do
  for lib in "" "$VAR"; do
    echo "== $cfg lib=${lib:-default}" >> gpurun_out/${TAG}_ab.log
    WT_B200_LIB=$lib PROBE_REPS=3 timeout 300 python tools/probe.py $cfg 2>&1 | grep "^render [12]" | cut -c1-75 >> gpurun_out/${TAG}_ab.log
  done
done
cat gpurun_out/${TAG}_ab.log
