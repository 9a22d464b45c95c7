#!/bin/bash
# GPU box: A/B of builds of the library (second / third render of tools/probe.py).  bash tools/ab_libs.sh <tag> "<probe args>;<probe args>..." <lib .so> [<lib .so> ...]   ("" = the in-tree build)
TAG=$1; CFGS=$2; shift 2
IFS=';' read -ra CL <<< "$CFGS"
for cfg in "${CL[@]}"; do
  for lib in "$@"; do
    echo "== $cfg lib=${lib:-default}" >> gpurun_out/${TAG}_ab.log
    WT_B200_LIB=$lib PROBE_REPS=3 timeout 300 python tools/probe.py $cfg 2>&1 | grep "^render [12]" | cut -c1-75 >> gpurun_out/${TAG}_ab.log
  done
done
cat gpurun_out/${TAG}_ab.log
