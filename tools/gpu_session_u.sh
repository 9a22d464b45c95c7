#!/bin/bash
# session U (2 GPUs): final build under torchrun, both arms; + the XML-scene GPU parity test
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "xml_scene" > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_u.log; tail -3 gpurun_out/pytest_gpu_u.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$T bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/u_bdpt_n2.json 2> gpurun_out/u_err.log
$T bench.py --gpus 2 --steps 1 --warmup 0 --impl reference > gpurun_out/u_reference_n2.json 2>> gpurun_out/u_err.log
$T bench.py --gpus 2 --steps 4 --warmup 3 --workload etoile --no-cpu-baseline > gpurun_out/u_etoile_n2.json 2>> gpurun_out/u_err.log
for f in gpurun_out/u_bdpt_n2.json gpurun_out/u_etoile_n2.json; do python tools/show_bench.py $f | head -3; done
tail -1 gpurun_out/u_reference_n2.json | cut -c1-200
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/u_err.log | tail -5
