#!/bin/bash
# session W: block size of the one-thread-per-item kernels (WT_BLOCK_T = 128 / 64 / 32); XML-scene parity with the re-parametrised fixture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "xml_scene" > gpurun_out/pytest_gpu_w0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_w0.log; tail -2 gpurun_out/pytest_gpu_w0.log
B="python bench.py --no-cpu-baseline --steps 4 --warmup 3"
for bt in 128 64 32; do
  export WT_BLOCK_T=$bt
  $B --workload etoile > gpurun_out/w_etoile_$bt.json 2> gpurun_out/w_err.log
  $B --workload cornell --steps 3 > gpurun_out/w_cornell_$bt.json 2>> gpurun_out/w_err.log
  $B > gpurun_out/w_bdpt_$bt.json 2>> gpurun_out/w_err.log
  $B --integrator plt_path > gpurun_out/w_path_$bt.json 2>> gpurun_out/w_err.log
done
export WT_BLOCK_T=32
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_w.log; tail -3 gpurun_out/pytest_gpu_w.log
for f in gpurun_out/w_*.json; do python tools/show_bench.py $f 2>/dev/null | head -3 | grep -v roofline; done
tail -3 gpurun_out/w_err.log
