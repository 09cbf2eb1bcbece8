#!/bin/bash
# C5 (general path) quick loop: general-path parity tests + bench at mesh 48
tag=${1:-r03c5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "general_path or c5 or band_sums or long_lines" > gpurun_out/${tag}_tests.log 2>&1
tail -n 5 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --workload C5 --mesh 48 --steps 3 --no-cpu-baseline --no-yardstick > gpurun_out/${tag}_bench_C5_48.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_C5_48.json'))
print('C5 48^3 value %.4g ms %.2f'%(d['value'], d['ms_per_step']))
for k,v in d['roofline']['kernels'].items(): print('  %-10s %.2f ms share %.3f tf %s'%(k, v['ms_per_pass'], v['share'], v['tflops']))
PY
