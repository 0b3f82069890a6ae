#!/bin/bash
# sliding-window max pool + parallel depthwise-wgrad final reduce: op tests, engine tests, benches
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -q -s -k "maxpool or stem or depthwise or mobilenet or resnet50_imagenet or resnet18 or resnext50" > gpurun_out/r2_pytest17.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest17.log | cut -c1-300
run() {   # label, bench args, env...
  local label=$1; local args=$2; shift; shift
  env "$@" timeout 300 python bench.py $args --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/r2_ab17.err | tail -1 > gpurun_out/r2_ab17_$label.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_ab17_$label.json')); c=d['roofline']['classes']
print('$label', round(d['value']), round(d['ms_per_step'],3), d['final_loss'])
print('   ', {k: round(v['ms'],2) for k,v in sorted(c.items(), key=lambda kv:-kv[1]['ms'])})" 2>&1 | tail -2
  tail -2 gpurun_out/r2_ab17.err | cut -c1-300
}
run rn50_slide0 "" B200_POOL_SLIDE=0
run rn50_slide1 "" B200_POOL_SLIDE=1
run rn50_slide0b "" B200_POOL_SLIDE=0
run rn50_slide1b "" B200_POOL_SLIDE=1
run mbv2 "--model mobilenet_v2 --batch 512" X=0
