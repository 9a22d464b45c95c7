#!/bin/bash
# One gpurun call: GPU parity tests, bench lines for the three workloads, ncu launch list, ncu --set full captures exported to CSV on the box
# (the .ncu-rep files are deleted: gpurun_out/ must stay under 64 MiB).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_bdpt.json 2> gpurun_out/bench_bdpt.err
python bench.py --workload etoile --steps 4 --warmup 3 > gpurun_out/bench_etoile.json 2> gpurun_out/bench_etoile.err
python bench.py --workload cornell --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cornell.json 2> gpurun_out/bench_cornell.err
python bench.py --integrator plt_path --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_path.json 2> gpurun_out/bench_path.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_bdpt.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch.log 2>&1
cap() {  # name kernel-regex skip bench-args...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/full_$name \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_full_$name.log 2>&1
  ncu -i gpurun_out/full_$name.ncu-rep --page raw --csv > gpurun_out/full_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$name.ncu-rep --page source --csv > gpurun_out/full_${name}_source.csv 2>/dev/null
  rm -f gpurun_out/full_$name.ncu-rep
}
cap bd_shade k_bd_shade 12 --spp-per-step 4
cap bd_connect4 'k_bd_connect<4>' 12 --spp-per-step 4
cap bd_connect3 'k_bd_connect<3>' 12 --spp-per-step 4
cap bd_gtraverse k_bd_gtraverse 12 --spp-per-step 4
cap bd_fsd_sample k_bd_fsd_sample 8 --spp-per-step 4
cap et_traverse 'k_traverse' 6 --workload etoile --spp-per-step 4
cap et_shade 'k_shade' 6 --workload etoile --spp-per-step 4
du -sm gpurun_out; ls -la gpurun_out
