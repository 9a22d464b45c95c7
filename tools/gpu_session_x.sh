#!/bin/bash
# session X: what the driver runs at round end, on the committed build
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/x_reference.json 2> gpurun_out/x_err.log
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/x_bench.json 2>> gpurun_out/x_err.log
python tools/show_bench.py gpurun_out/x_bench.json gpurun_out/x_reference.json
tail -3 gpurun_out/x_err.log
