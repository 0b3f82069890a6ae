#!/bin/bash
# programmatic dependent launch: full GPU suite with B200_PDL=1, then A/B of the contract bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_PDL=1 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest12.log 2>&1
echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest12.log | cut -c1-300
grep -E "unit features|MobileNet" gpurun_out/r2_pytest12.log | cut -c1-400 | head -30
for pdl in 0 1 0 1; do
  B200_PDL=$pdl timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_pdl_$pdl.err | tail -1 > gpurun_out/r2_pdl_$pdl.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_pdl_$pdl.json')); print('pdl=$pdl', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['final_loss'], {k: round(v['ms'],3) for k,v in d['roofline']['classes'].items() if k.startswith('conv')})" 2>&1 | tail -1
  tail -2 gpurun_out/r2_pdl_$pdl.err | cut -c1-300
done
for m in "--model resnext --depth 101 --batch 128" "--model mobilenet_v2 --batch 512"; do
  for pdl in 0 1; do
    B200_PDL=$pdl timeout 300 python bench.py $m --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$m pdl=$pdl', round(d['value']), round(d['ms_per_step'],3), d['final_loss'])" 2>&1 | tail -1
  done
done
