#!/bin/bash
# 4-GPU box: on-device DDP / SyncBN test (2 ranks), then the contract bench at N=4 and N=2 exactly as the driver
# launches it; wall time and exit code of every torchrun are logged (teardown must not hang).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_multi4.log
nvidia-smi --query-gpu=index,name --format=csv > $L 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s >> $L 2>&1
echo "pytest rc=$?" | tee -a $L
for n in 4 2; do
  t0=$(date +%s)
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.out 2> gpurun_out/r2_bench_n$n.err
  rc=$?
  t1=$(date +%s)
  echo "N=$n rc=$rc wall=$((t1-t0))s" | tee -a $L
  tail -1 gpurun_out/r2_bench_n$n.out | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=', d['n_gpus'], round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'loss', d['final_loss'], d['roofline']['classes'].get('allreduce_nccl'))" 2>&1 | tail -1 | tee -a $L
  grep -i -E "warning.*capture|fall|error" gpurun_out/r2_bench_n$n.err | head -3 | cut -c1-300
done
