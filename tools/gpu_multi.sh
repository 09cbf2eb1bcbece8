#!/bin/bash
# multi-GPU check (gpurun --gpus N): NCCL parity tests, in-library all-reduce, bench under torchrun at N and N=1 for reference
tag=${1:-r03m}; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "nccl or multi_context or strided" > gpurun_out/${tag}_tests.log 2>&1
tail -n 8 gpurun_out/${tag}_tests.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 5 --warmup 3 \
   > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
echo "bench exit $?"; tail -n 3 gpurun_out/${tag}_bench_${n}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_${n}gpu.json'))
print('C2 N=%d value %.4g ms %.3f parity %s stages %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity_check'), d['stages']))
for e in d.get('configs',[]):
    print('  ', e.get('workload','?')[:30], e.get('error') or ('%.4g ms %.3f ar %s'%(e['value'], e['ms_per_step'], e['stages']['allreduce_ms'])))
PY
