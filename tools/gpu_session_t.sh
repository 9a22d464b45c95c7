#!/bin/bash
# session T: validation of the final build -- full GPU test run, smoke, both bench arms, the other workloads, ncu launch lists
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/t_reference.json 2> gpurun_out/t_err.log
python bench.py > gpurun_out/t_bdpt.json 2>> gpurun_out/t_err.log
python bench.py --workload etoile > gpurun_out/t_etoile.json 2>> gpurun_out/t_err.log
python bench.py --workload cornell --steps 3 --no-cpu-baseline > gpurun_out/t_cornell.json 2>> gpurun_out/t_err.log
python bench.py --integrator plt_path --no-cpu-baseline > gpurun_out/t_path.json 2>> gpurun_out/t_err.log
python bench.py --sampler sobolld --no-cpu-baseline > gpurun_out/t_bdpt_sobolld.json 2>> gpurun_out/t_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_bdpt.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_etoile.csv \
    python bench.py --workload etoile --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch_et.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_cornell.csv \
    python bench.py --workload cornell --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_cornell.log 2>&1
for f in gpurun_out/t_*.json; do python tools/show_bench.py $f 2>/dev/null | head -5; done
tail -5 gpurun_out/t_err.log
