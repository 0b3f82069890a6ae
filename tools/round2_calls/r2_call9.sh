#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_multi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s >> gpurun_out/r2_multi.log 2>&1
tail -25 gpurun_out/r2_multi.log | cut -c1-400
for ov in 1 0 1 0; do
  B200_AR_OVERLAP=$ov timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_bench9_$ov.err | tail -1 > gpurun_out/r2_bench9_$ov.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench9_$ov.json')); print('overlap=$ov n=', d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'loss', d['final_loss'], d['roofline']['classes'].get('allreduce_nccl'))" 2>&1 | tail -1
  tail -3 gpurun_out/r2_bench9_$ov.err | cut -c1-300
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n=1', round(d['value']), round(d['ms_per_step'],3))"
