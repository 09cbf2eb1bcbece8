#!/bin/bash
# ncu --set full capture (with source) of selected kernels on one workload:  tools/gpu_ncu.sh <tag> <workload> <mesh> <kernel-regex> [count]
tag=$1; wl=$2; mesh=$3; rx=$4; cnt=${5:-3}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 2 -c $cnt -o gpurun_out/${tag}_prof -f \
    python bench.py --workload $wl --mesh $mesh --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-yardstick > gpurun_out/${tag}_ncu.log 2>&1
tail -n 3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}_prof.ncu-rep
