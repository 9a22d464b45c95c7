#!/bin/bash
# session N: k_bd_resolve split into the cheap part + k_bd_flux on a compacted list, A/B against _variants/libwt_noflux.so
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "bdpt or golden" > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n.log
tail -4 gpurun_out/pytest_gpu_n.log
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
for v in flux noflux; do
  if [ $v = noflux ]; then export WT_B200_LIB=wave_tracer_b200/_variants/libwt_noflux.so; fi
  $B > gpurun_out/n_bdpt_$v.json 2> gpurun_out/n_err.log
  $B --workload cornell --steps 3 > gpurun_out/n_cornell_$v.json 2>> gpurun_out/n_err.log
done
for f in gpurun_out/n_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -5 gpurun_out/n_err.log
