#!/bin/bash
# C5 (Mix&Match) spot sizes with the final kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for sz in 128 288 192; do
  timeout 70 python bench.py --size $sz --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('size=$sz', round(d['value']), round(d['ms_per_step'],3))" 2>&1 | tail -1 | tee -a gpurun_out/r02_c5_sizes.log
done
