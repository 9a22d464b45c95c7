#!/bin/bash
# session Y: the C++ host example on the GPU + the full GPU suite on the rebuilt library
mkdir -p gpurun_out
./examples/host_render 192 8 > gpurun_out/host_render.log 2>&1; echo "rc=$?" >> gpurun_out/host_render.log; cat gpurun_out/host_render.log
python -m pytest tests -x -q -m gpu -s -k "host_example" > gpurun_out/pytest_gpu_y0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_y0.log; grep -E "host_render:|passed|failed|Error|assert" gpurun_out/pytest_gpu_y0.log | head
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
