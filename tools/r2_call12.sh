#!/bin/bash
# programmatic dependent launch (B200_PDL) + fused stem bn/relu/maxpool (B200_FUSE_STEM_POOL): full GPU suite with both
# on, then A/B of the contract bench over the four combinations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_PDL=1 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest12.log 2>&1
echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest12.log | cut -c1-300
grep -E "unit features|MobileNet" gpurun_out/r2_pytest12.log | cut -c1-400 | head -30
for cfg in "0 0" "1 0" "0 1" "1 1" "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  B200_PDL=$1 B200_FUSE_STEM_POOL=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_ab12.err | tail -1 > gpurun_out/r2_ab12_$1$2.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_ab12_$1$2.json')); print('pdl=$1 stem=$2', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['final_loss'], {k: round(v['ms'],3) for k,v in d['roofline']['classes'].items() if k.startswith('conv')})" 2>&1 | tail -1
  tail -2 gpurun_out/r2_ab12.err | cut -c1-300
done
for m in "--model resnext --depth 101 --batch 128" "--model mobilenet_v2 --batch 512"; do
  for pdl in 0 1; do
    B200_PDL=$pdl timeout 300 python bench.py $m --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$m pdl=$pdl', round(d['value']), round(d['ms_per_step'],3), d['final_loss'])" 2>&1 | tail -1
  done
done
