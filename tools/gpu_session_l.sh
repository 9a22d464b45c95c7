#!/bin/bash
# session L: separating-axis rejection in front of the cone-triangle test, A/B against the build without it (_variants/libwt_noqr.so)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
for v in qr noqr; do
  if [ $v = noqr ]; then export WT_B200_LIB=wave_tracer_b200/_variants/libwt_noqr.so; fi
  $B > gpurun_out/l_bdpt_$v.json 2> gpurun_out/l_err.log
  $B --workload etoile > gpurun_out/l_etoile_$v.json 2>> gpurun_out/l_err.log
  $B --workload cornell --steps 3 > gpurun_out/l_cornell_$v.json 2>> gpurun_out/l_err.log
done
for f in gpurun_out/l_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -5 gpurun_out/l_err.log
