#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list, ncu --set full captures of the BDPT kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_bdpt.json 2> gpurun_out/bench_bdpt.err; tail -c 600 gpurun_out/bench_bdpt.json
python bench.py --integrator plt_path --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_path.json 2> gpurun_out/bench_path.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_bdpt.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_launch.log 2>&1
for k in k_bd_connect k_bd_gtraverse k_bd_shade k_bd_fsd_sample k_bd_resolve; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 3 -f -o gpurun_out/full_$k \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp-per-step 4 > gpurun_out/ncu_full_$k.log 2>&1
done
ls -la gpurun_out
