#!/bin/bash
# session I: occupancy variants of the group-traversal kernels (WT_GT_MINB = min blocks/SM), clock-sampler diagnostic, pool-size sweeps
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
V=wave_tracer_b200/_variants
for v in 5 1 6 8; do
  export WT_B200_LIB=$V/libwt_minb$v.so
  $B > gpurun_out/i_bdpt_v$v.json 2> gpurun_out/i_err.log
  $B --workload etoile --pool 4194304 > gpurun_out/i_etoile_v$v.json 2>> gpurun_out/i_err.log
  $B --workload cornell --steps 3 > gpurun_out/i_cornell_v$v.json 2>> gpurun_out/i_err.log
done
export WT_B200_LIB=$V/libwt_minb5.so
WT_BENCH_NO_CLOCKS=1 $B > gpurun_out/i_bdpt_noclk.json 2>> gpurun_out/i_err.log
WT_BENCH_NO_CLOCKS=1 $B --workload etoile > gpurun_out/i_etoile_noclk.json 2>> gpurun_out/i_err.log
$B --workload etoile > gpurun_out/i_etoile_pool20.json 2>> gpurun_out/i_err.log
$B --workload etoile --pool 2097152 > gpurun_out/i_etoile_pool21.json 2>> gpurun_out/i_err.log
$B --workload etoile --pool 8388608 > gpurun_out/i_etoile_pool23.json 2>> gpurun_out/i_err.log
$B --pool 65536 > gpurun_out/i_bdpt_pool16.json 2>> gpurun_out/i_err.log
$B --pool 131072 > gpurun_out/i_bdpt_pool17.json 2>> gpurun_out/i_err.log
$B --integrator plt_path --pool 4194304 > gpurun_out/i_path_pool22.json 2>> gpurun_out/i_err.log
$B --integrator plt_path > gpurun_out/i_path.json 2>> gpurun_out/i_err.log
$B --workload cornell --steps 3 --pool 131072 > gpurun_out/i_cornell_pool17.json 2>> gpurun_out/i_err.log
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for f in gpurun_out/i_*.json; do python tools/show_bench.py $f 2>/dev/null | head -1; done
tail -5 gpurun_out/i_err.log
