#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_bdpt.json 2> gpurun_out/bench_bdpt.err
python bench.py --workload etoile --steps 4 --warmup 3 > gpurun_out/bench_etoile.json 2> gpurun_out/bench_etoile.err
python bench.py --workload cornell --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cornell.json 2> gpurun_out/bench_cornell.err
python bench.py --integrator plt_path --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_path.json 2> gpurun_out/bench_path.err
python bench.py --sampler sobolld --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bdpt_sobolld.json 2> gpurun_out/bench_bdpt_sobolld.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_bdpt.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_etoile.csv \
    python bench.py --workload etoile --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch_et.log 2>&1
du -sm gpurun_out
