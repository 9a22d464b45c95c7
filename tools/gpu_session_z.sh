#!/bin/bash
# session Z: the round-end sequence once more on the final commit (oracle relinked against the dynamic libstdc++)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench.json 2> gpurun_out/z_err.log; python tools/show_bench.py gpurun_out/z_bench.json | head -3; tail -2 gpurun_out/z_err.log
