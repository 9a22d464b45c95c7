#!/bin/bash
# One parametrised GPU session (replaces round 1's one-shot gpu_session_[a-z].sh).  Run on the B200 box through gpurun, from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <tag> <step> [<step> ...]'
# Every step writes into gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   tests              pytest -m gpu, with the printed parity numbers (-s)
#   tests:<expr>       the same, restricted with -k <expr>
#   smoke              __graft_entry__.smoke()
#   bench:<workload>   python bench.py --workload <workload> (short: --steps 3 --warmup 3)
#   launches:<workload>  ncu launch list of the bench command (gpu__time_duration.sum), 400 launches
#   ncu:<workload>:<kernel-regex>   one `ncu --set full` capture of the first launches matching the regex
#   py:<script>        python <script>
set -u
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
for step in "$@"; do
  kind=${step%%:*}; arg=${step#*:}
  case $kind in
    tests)
      if [ "$arg" = "tests" ]; then timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1
      else timeout 1500 python -m pytest tests -m gpu -q -s -k "$arg" > gpurun_out/${TAG}_pytest_gpu_k.log 2>&1; fi
      tail -5 gpurun_out/${TAG}_pytest_gpu*.log ;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log ;;
    bench) timeout 900 python bench.py --workload ${arg%%,*} --steps 3 --warmup 3 $(echo "$arg" | cut -s -d, -f2- | tr ',' ' ') > gpurun_out/${TAG}_bench_${arg%%,*}.json 2> gpurun_out/${TAG}_bench_${arg%%,*}.err; tail -c 600 gpurun_out/${TAG}_bench_${arg%%,*}.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${arg}.csv python bench.py --workload $arg --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_${arg}.log 2>&1 ;;
    ncu)
      wl=${arg%%:*}; rx=${arg#*:}
      timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$rx" -c 3 -o gpurun_out/${TAG}_ncu_${wl} -f python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_${wl}.log 2>&1 ;;
    py) timeout 1200 python $arg > gpurun_out/${TAG}_$(basename ${arg%% *} .py).log 2>&1; tail -5 gpurun_out/${TAG}_$(basename ${arg%% *} .py).log ;;
    *) echo "unknown step $step" ;;
  esac
done
