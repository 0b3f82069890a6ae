#!/bin/bash
# sliding-window depthwise kernels: op tests + MobileNet parity, MobileNet-v2 bench A/B; fused-SGD microbenchmark;
# compute-sanitizer pass
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -q -s -k "depthwise or mobilenet" > gpurun_out/r2_pytest15.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest15.log | cut -c1-300
for dw in 0 1 0 1; do
  B200_DW3X3=$dw timeout 300 python bench.py --model mobilenet_v2 --batch 512 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('mbv2 dw3x3=$dw', round(d['value']), round(d['ms_per_step'],3), d['final_loss'])" 2>&1 | tail -1
done
timeout 120 python tools/sgd_bench.py 2>&1 | tail -6
bash tools/sanitize.sh r02 2>&1 | tail -12
