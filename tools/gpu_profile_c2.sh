#!/bin/bash
# profiles for the headline (C2): launch list of the bench command, then one ncu --set full capture of its three kernels
tag=${1:-r03p}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_C2.csv \
    python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-yardstick > gpurun_out/${tag}_launches_C2.log 2>&1
tail -n 2 gpurun_out/${tag}_launches_C2.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bands_sub|eig_small_hq|sweep_ahc_mma" -s 9 -c 3 -o gpurun_out/${tag}_c2 -f \
    python bench.py --steps 1 --warmup 3 --no-configs --no-cpu-baseline --no-yardstick > gpurun_out/${tag}_ncu_C2.log 2>&1
tail -n 2 gpurun_out/${tag}_ncu_C2.log
ls -la gpurun_out/${tag}_c2.ncu-rep
