#!/bin/bash
# Standard GPU session (run through gpurun): tests, smoke, bench; everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_session.sh <tag> [tests|notests] [bench args...]'
tag=${1:-s}; what=${2:-tests}; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
if [ "$what" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -150 > gpurun_out/${tag}_pytest_gpu.log
  tail -3 gpurun_out/${tag}_pytest_gpu.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 1800 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
