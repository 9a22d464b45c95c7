#!/bin/bash
# session M (2 GPUs): bench.py under torchrun, both arms
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/m_bdpt_n2.json 2> gpurun_out/m_err.log
$T bench.py --gpus 2 --steps 1 --warmup 0 --impl reference > gpurun_out/m_reference_n2.json 2>> gpurun_out/m_err.log
$T bench.py --gpus 2 --steps 3 --warmup 3 --workload etoile > gpurun_out/m_etoile_n2.json 2>> gpurun_out/m_err.log
for f in gpurun_out/m_*_n2.json; do tail -1 $f | cut -c1-600; echo; done
tail -5 gpurun_out/m_err.log
