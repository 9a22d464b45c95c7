#!/bin/bash
# session J: persistent pinned counters/events (no per-render cudaMallocHost), gaussian surface profile parity, cornell launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
$B > gpurun_out/j_bdpt.json 2> gpurun_out/j_err.log
$B --workload etoile > gpurun_out/j_etoile.json 2>> gpurun_out/j_err.log
$B --workload cornell --steps 3 > gpurun_out/j_cornell.json 2>> gpurun_out/j_err.log
$B --integrator plt_path > gpurun_out/j_path.json 2>> gpurun_out/j_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_cornell.csv \
    python bench.py --workload cornell --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_cornell.log 2>&1
for f in gpurun_out/j_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3; done
python tools/ncu_launch_summary.py gpurun_out/launches_cornell.csv | head -12
tail -5 gpurun_out/j_err.log
