#!/bin/bash
# session P: more code-size variants.  v1 = single ranked-push site in the node step; raynode = v1 + rolled 2x4 child loop in the per-thread ray
# traversal; v2 (the in-tree library) = v1 + bd_connect compiled per strategy class; v3 = v2 + rolled collect_edges / clip_triangle_z + out-of-line RNG draw
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
V=wave_tracer_b200/_variants
for v in v1 raynode v2 v3; do
  if [ $v = v2 ]; then unset WT_B200_LIB; else export WT_B200_LIB=$V/libwt_$v.so; fi
  $B > gpurun_out/p_bdpt_$v.json 2> gpurun_out/p_err.log
  $B --workload etoile > gpurun_out/p_etoile_$v.json 2>> gpurun_out/p_err.log
  $B --workload cornell --steps 3 > gpurun_out/p_cornell_$v.json 2>> gpurun_out/p_err.log
  $B --integrator plt_path > gpurun_out/p_path_$v.json 2>> gpurun_out/p_err.log
done
export WT_B200_LIB=$V/libwt_v3.so
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_p.log
tail -3 gpurun_out/pytest_gpu_p.log
for f in gpurun_out/p_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -5 gpurun_out/p_err.log
