#!/bin/bash
# session R: with the instruction-fetch stall gone, re-measure occupancy (6 / 8 blocks per SM) and a skip of the ranked push when <= 1 child passes
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
V=wave_tracer_b200/_variants
for v in base mb6 mb8 skip1; do
  if [ $v = base ]; then unset WT_B200_LIB; else export WT_B200_LIB=$V/libwt_$v.so; fi
  $B --workload etoile > gpurun_out/r_etoile_$v.json 2> gpurun_out/r_err.log
  $B --workload cornell --steps 3 > gpurun_out/r_cornell_$v.json 2>> gpurun_out/r_err.log
  $B > gpurun_out/r_bdpt_$v.json 2>> gpurun_out/r_err.log
done
for f in gpurun_out/r_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -5 gpurun_out/r_err.log
