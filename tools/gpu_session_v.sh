#!/bin/bash
# session V: XML-scene mismatch bisect; warp-uniform traversal loop re-measured now that instruction fetch no longer bounds the kernel
mkdir -p gpurun_out
python tools/xml_diag.py > gpurun_out/xml_diag.log 2>&1; cat gpurun_out/xml_diag.log | cut -c1-330
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
for v in base wu; do
  if [ $v = base ]; then unset WT_B200_LIB; else export WT_B200_LIB=wave_tracer_b200/_variants/libwt_$v.so; fi
  $B --workload etoile > gpurun_out/v_etoile_$v.json 2> gpurun_out/v_err.log
  $B --workload cornell --steps 3 > gpurun_out/v_cornell_$v.json 2>> gpurun_out/v_err.log
  $B > gpurun_out/v_bdpt_$v.json 2>> gpurun_out/v_err.log
done
export WT_B200_LIB=wave_tracer_b200/_variants/libwt_wu.so
python -m pytest tests -m gpu -q -k "traverse or cone or etoile or golden" > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v.log; tail -3 gpurun_out/pytest_gpu_v.log
for f in gpurun_out/v_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -3 gpurun_out/v_err.log
