#!/bin/bash
# programmatic dependent launch (B200_PDL), fused stem bn/relu/maxpool (B200_FUSE_STEM_POOL), high-priority main chain
# (B200_MAIN_PRIORITY), row-quad activation masks (B200_BN_ACT_MASK): full GPU suite with PDL on, then one-at-a-time A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B200_PDL=1 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest12.log 2>&1
echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest12.log | cut -c1-300
grep -E "unit features|MobileNet" gpurun_out/r2_pytest12.log | cut -c1-400 | head -30
B200_BN_ACT_MASK=1 timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_parity.py -m gpu -q -x -k "resnet50 or bottleneck or resnet18 or fused_ce" > gpurun_out/r2_pytest12_mask.log 2>&1
echo "pytest(mask) rc=$?"; tail -3 gpurun_out/r2_pytest12_mask.log | cut -c1-300
run() {   # label, env assignments...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_ab12.err | tail -1 > gpurun_out/r2_ab12_$label.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_ab12_$label.json')); print('$label', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['final_loss'], {k: round(v['ms'],3) for k,v in d['roofline']['classes'].items() if k.startswith('conv')})" 2>&1 | tail -1
  tail -2 gpurun_out/r2_ab12.err | cut -c1-300
}
BASE="B200_PDL=0 B200_FUSE_STEM_POOL=0 B200_MAIN_PRIORITY=0 B200_BN_ACT_MASK=0"
for rep in 1 2; do
  run base_$rep $BASE
  run pdl_$rep $BASE B200_PDL=1
  run stem_$rep $BASE B200_FUSE_STEM_POOL=1
  run prio_$rep $BASE B200_MAIN_PRIORITY=1
  run mask_$rep $BASE B200_BN_ACT_MASK=1
  run all_$rep B200_PDL=1 B200_FUSE_STEM_POOL=1 B200_MAIN_PRIORITY=1 B200_BN_ACT_MASK=1
done
for m in "--model resnext --depth 101 --batch 128" "--model mobilenet_v2 --batch 512"; do
  for pdl in 0 1; do
    B200_PDL=$pdl B200_MAIN_PRIORITY=$pdl timeout 300 python bench.py $m --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$m pdl+prio=$pdl', round(d['value']), round(d['ms_per_step'],3), d['final_loss'])" 2>&1 | tail -1
  done
done
