#!/bin/bash
# 2-GPU validation: on-device DDP / SyncBN test, then the contract bench with the overlapped and the flat all-reduce;
# wall time and exit code of every torchrun are logged (teardown must not hang).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2_multi11.log
nvidia-smi --query-gpu=index,name --format=csv > $L 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s >> $L 2>&1
echo "pytest rc=$?" >> $L
tail -12 $L | cut -c1-300
for ov in 1 0; do
  t0=$(date +%s)
  B200_AR_OVERLAP=$ov timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$ov bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench11_$ov.out 2> gpurun_out/r2_bench11_$ov.err
  rc=$?
  t1=$(date +%s)
  echo "overlap=$ov rc=$rc wall=$((t1-t0))s" | tee -a $L
  tail -1 gpurun_out/r2_bench11_$ov.out | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n=', d['n_gpus'], round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'loss', d['final_loss'], d['roofline']['classes'].get('allreduce_nccl'))" 2>&1 | tail -1 | tee -a $L
  tail -2 gpurun_out/r2_bench11_$ov.err | cut -c1-300
done
