#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bdpt.json 2> gpurun_out/bench_bdpt.err
python bench.py --workload etoile --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_etoile.json 2> gpurun_out/bench_etoile.err
python bench.py --workload cornell --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cornell.json 2> gpurun_out/bench_cornell.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_etoile.csv \
    python bench.py --workload etoile --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch_et.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_bdpt.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch.log 2>&1
du -sm gpurun_out
