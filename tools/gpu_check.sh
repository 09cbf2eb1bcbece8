#!/bin/bash
# One gpurun call: GPU parity tests + the default bench line (+ optional extras); everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_check.sh <tag> [pytest-args...]
tag=${1:-r03}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rf --durations=15 "$@" > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_gpu_tests.log
tail -n 40 gpurun_out/${tag}_gpu_tests.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"
head -c 1500 gpurun_out/${tag}_bench.json
tail -n 5 gpurun_out/${tag}_bench.err
