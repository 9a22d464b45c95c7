#!/bin/bash
# session H: ray-range culling A/B (flags 32 = WTGPU_RENDER_NO_RAY_CULL), pool-size sweeps
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "ray culling|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -8
B="python bench.py --no-cpu-baseline"
$B --steps 4 --warmup 3 > gpurun_out/h_bdpt.json 2> gpurun_out/h_bdpt.err
$B --steps 4 --warmup 3 --flags 32 > gpurun_out/h_bdpt_nocull.json 2>> gpurun_out/h_bdpt.err
$B --steps 4 --warmup 3 --pool 524288 > gpurun_out/h_bdpt_pool19.json 2>> gpurun_out/h_bdpt.err
$B --steps 4 --warmup 3 --pool 1048576 > gpurun_out/h_bdpt_pool20.json 2>> gpurun_out/h_bdpt.err
$B --workload etoile --steps 4 --warmup 3 > gpurun_out/h_etoile.json 2> gpurun_out/h_etoile.err
$B --workload etoile --steps 4 --warmup 3 --flags 32 > gpurun_out/h_etoile_nocull.json 2>> gpurun_out/h_etoile.err
$B --workload etoile --steps 4 --warmup 3 --pool 4194304 > gpurun_out/h_etoile_pool22.json 2>> gpurun_out/h_etoile.err
$B --workload cornell --steps 3 --warmup 3 > gpurun_out/h_cornell.json 2> gpurun_out/h_cornell.err
$B --workload cornell --steps 3 --warmup 3 --flags 32 > gpurun_out/h_cornell_nocull.json 2>> gpurun_out/h_cornell.err
$B --integrator plt_path --steps 4 --warmup 3 > gpurun_out/h_path.json 2> gpurun_out/h_path.err
$B --integrator plt_path --steps 4 --warmup 3 --flags 32 > gpurun_out/h_path_nocull.json 2>> gpurun_out/h_path.err
for f in gpurun_out/h_*.json; do python tools/show_bench.py $f | head -2; done
