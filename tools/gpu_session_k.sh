#!/bin/bash
# session K: k_bd_resolve with eight lanes per walker
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "bdpt or golden" > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log
tail -4 gpurun_out/pytest_gpu_k.log
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
$B > gpurun_out/k_bdpt.json 2> gpurun_out/k_err.log
$B --workload cornell --steps 3 > gpurun_out/k_cornell.json 2>> gpurun_out/k_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_bdpt.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch.log 2>&1
for f in gpurun_out/k_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3; done
python tools/ncu_launch_summary.py gpurun_out/launches_bdpt.csv 2>/dev/null | head -8
tail -5 gpurun_out/k_err.log
