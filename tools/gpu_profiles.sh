#!/bin/bash
# One gpurun call that refreshes everything under profiles/ that comes from ncu: the launch list of the headline bench command
# and one `--set full` capture (with source) of the kernels of C2, C3, C4 and C5.   usage: tools/gpu_profiles.sh <tag>
tag=${1:-r03v}; part=${2:-all}      # part: small (C2, C3, C4), c5, all — one call may bring back at most 64 MiB
mkdir -p gpurun_out
B="--no-configs --no-cpu-baseline --no-yardstick"
if [ "$part" != "c5" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_C2.csv \
    python bench.py --steps 2 --warmup 3 $B > gpurun_out/${tag}_launches_C2.log 2>&1
tail -n 1 gpurun_out/${tag}_launches_C2.log | head -c 300; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bands_sub|eig_small_hq|sweep_ahc_mma" -s 9 -c 3 -o gpurun_out/${tag}_c2 -f \
    python bench.py --steps 1 --warmup 3 $B > gpurun_out/${tag}_ncu_C2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bands_sub|sweep_photo_mma|sweep_ahc_mma" -s 3 -c 3 -o gpurun_out/${tag}_c4 -f \
    python bench.py --workload C4 --mesh 64 --steps 1 --warmup 3 $B > gpurun_out/${tag}_ncu_C4.log 2>&1
fi
if [ "$part" != "small" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fourier_dense|tridiag|tql|zrot|eig_rot_large" -s 7 -c 5 -o gpurun_out/${tag}_c5 -f \
    python bench.py --workload C5 --mesh 32 --steps 1 --warmup 3 $B > gpurun_out/${tag}_ncu_C5.log 2>&1
fi
if [ "$part" != "c5" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bands_sub|mep_sums|eig_small_hq" -s 6 -c 3 -o gpurun_out/${tag}_c3 -f \
    python bench.py --workload C3 --mesh 128 --steps 1 --warmup 3 $B > gpurun_out/${tag}_ncu_C3.log 2>&1
fi
ls -la gpurun_out/${tag}_*.ncu-rep
