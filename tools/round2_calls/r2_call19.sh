#!/bin/bash
# igemm shared-memory budget 200 vs 224 KB (3 instead of 2 stages for 256-wide tiles): layers + step
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B200_IGEMM_SMEM_KB=224 B200_IGEMM_DEBUG=1 timeout 200 python tools/layer_bench.py l3_1x1_1024_256 2>&1 | grep -E "igemm\]" | sort | uniq -c | head -4
for kb in 200 224; do
  echo "-- B200_IGEMM_SMEM_KB=$kb"
  for l in l1_1x1_64_256 l2_1x1_512_128 l3_1x1_256_1024 l3_1x1_1024_256 l4_1x1_512_2048 l4_1x1_2048_512 l2_ds_256_512_s2; do
    B200_IGEMM_SMEM_KB=$kb timeout 200 python tools/layer_bench.py $l 2>/dev/null | tail -1 | cut -c1-330
  done
done
for kb in 200 224 200 224; do
  B200_IGEMM_SMEM_KB=$kb timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); c=d['roofline']['classes']; print('kb=$kb', round(d['value']), round(d['ms_per_step'],3), d['final_loss'], {k: round(v['ms'],3) for k,v in c.items() if k.startswith('conv')})" 2>&1 | tail -1
done
B200_IGEMM_SMEM_KB=224 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv or fused_bn" 2>&1 | tail -2
