#!/bin/bash
# session O: code-size variants of the group traversal (rolled cone-edge loop, shared-memory rank), plt_bdpt pool sizes re-measured without the host overheads
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
V=wave_tracer_b200/_variants
for v in base roll rank both; do
  if [ $v = base ]; then unset WT_B200_LIB; else export WT_B200_LIB=$V/libwt_$v.so; fi
  $B --workload etoile > gpurun_out/o_etoile_$v.json 2> gpurun_out/o_err.log
  $B --workload cornell --steps 3 > gpurun_out/o_cornell_$v.json 2>> gpurun_out/o_err.log
  $B > gpurun_out/o_bdpt_$v.json 2>> gpurun_out/o_err.log
done
export WT_B200_LIB=$V/libwt_both.so
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
tail -3 gpurun_out/pytest_gpu_o.log
unset WT_B200_LIB
$B --pool 524288 > gpurun_out/o_bdpt_pool19.json 2>> gpurun_out/o_err.log
$B --pool 1048576 > gpurun_out/o_bdpt_pool20.json 2>> gpurun_out/o_err.log
for f in gpurun_out/o_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -5 gpurun_out/o_err.log
