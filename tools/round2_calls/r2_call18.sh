#!/bin/bash
# owned-n-tile (weight-stationary) igemm walk: op tests, engine parity, per-layer and step A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_parity.py -m gpu -q -x -k "conv or fused_bn or resnet50 or resnet18 or resnext50 or bottleneck or mobilenet_v2" > gpurun_out/r2_pytest18.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest18.log | cut -c1-300
B200_IGEMM_DEBUG=1 timeout 200 python tools/layer_bench.py l3_1x1_256_1024 2>&1 | grep -E "igemm\]" | sort | uniq -c | head -8
for own in 0 1; do
  echo "-- B200_IGEMM_OWN_NTILE=$own"
  for l in l2_1x1_128_512 l2_1x1_512_128 l3_1x1_256_1024 l3_1x1_1024_256 l2_ds_256_512_s2; do
    B200_IGEMM_OWN_NTILE=$own timeout 200 python tools/layer_bench.py $l 2>/dev/null | tail -1 | cut -c1-400
  done
done
for own in 0 1 0 1; do
  B200_IGEMM_OWN_NTILE=$own timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); c=d['roofline']['classes']; print('own=$own', round(d['value']), round(d['ms_per_step'],3), d['final_loss'], {k: round(v['ms'],3) for k,v in c.items() if k.startswith('conv')})" 2>&1 | tail -1
done
