#!/bin/bash
# session Q: ncu --set full captures (raw + source pages) of the current kernels, e2e breakdown, full GPU test run
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/e2e_diag.py etoile > gpurun_out/e2e_diag.log 2>&1; python tools/e2e_diag.py double_slits >> gpurun_out/e2e_diag.log 2>&1
cat gpurun_out/e2e_diag.log | tail -8
cap() {  # name kernel-regex skip bench-args...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/full_$name \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_full_$name.log 2>&1
  ncu -i gpurun_out/full_$name.ncu-rep --page raw --csv > gpurun_out/full_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$name.ncu-rep --page source --csv > gpurun_out/full_${name}_source.csv 2>/dev/null
  rm -f gpurun_out/full_$name.ncu-rep
}
cap q_et_gtraverse 'k_gtraverse' 4 --workload etoile --spp-per-step 4
cap q_et_shade 'k_shade' 4 --workload etoile --spp-per-step 4
cap q_co_gtraverse 'k_bd_gtraverse' 6 --workload cornell --spp-per-step 1
cap q_bd_gtraverse 'k_bd_gtraverse' 12 --spp-per-step 4
cap q_bd_resolve 'k_bd_resolve' 12 --spp-per-step 4
cap q_bd_shade 'k_bd_shade' 12 --spp-per-step 4
cap q_bd_connect3 'k_bd_connect' 63 --spp-per-step 4     # five class launches per iteration: 5*12+3 = class 3 of iteration 12
cap q_bd_connect4 'k_bd_connect' 64 --spp-per-step 4
ls -la gpurun_out/full_q_* | awk '{print $5, $9}'
