#!/bin/bash
mkdir -p gpurun_out
python tools/etoile_diag.py > gpurun_out/etoile_diag.log 2>&1; tail -30 gpurun_out/etoile_diag.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bdpt_timed.json 2> gpurun_out/bench_bdpt.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-kernel-timing > gpurun_out/bench_bdpt_untimed.json 2>> gpurun_out/bench_bdpt.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bdpt_timed2.json 2>> gpurun_out/bench_bdpt.err
cap() {  # name kernel-regex skip bench-args...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/full_$name \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_full_$name.log 2>&1
  ncu -i gpurun_out/full_$name.ncu-rep --page raw --csv > gpurun_out/full_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$name.ncu-rep --page source --csv > gpurun_out/full_${name}_source.csv 2>/dev/null
  rm -f gpurun_out/full_$name.ncu-rep
}
cap et_gtraverse 'k_gtraverse' 4 --workload etoile --spp-per-step 4
cap et_shade2 'k_shade' 4 --workload etoile --spp-per-step 4
cap bd_resolve 'k_bd_resolve' 12 --spp-per-step 4
cap bd_fsd_sample2 'k_bd_fsd_sample' 8 --spp-per-step 4
du -sm gpurun_out
