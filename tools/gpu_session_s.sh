#!/bin/bash
# session S: gaussian2d Ige out of line + rolled (k_bd_resolve 5088 -> 2672 SASS instructions), A/B
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
for v in base ige; do
  if [ $v = base ]; then unset WT_B200_LIB; else export WT_B200_LIB=wave_tracer_b200/_variants/libwt_$v.so; fi
  $B > gpurun_out/s_bdpt_$v.json 2> gpurun_out/s_err.log
  $B --workload cornell --steps 3 > gpurun_out/s_cornell_$v.json 2>> gpurun_out/s_err.log
done
python -m pytest tests -m gpu -q -k "bdpt or golden" > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s.log
tail -3 gpurun_out/pytest_gpu_s.log
for f in gpurun_out/s_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -5 gpurun_out/s_err.log
