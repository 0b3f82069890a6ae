#!/bin/bash
# SGD clearing via memset, depthwise wgrad on the side stream: full GPU suite + benches of the three models
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest16.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest16.log | cut -c1-300
run() {   # label, bench args
  local label=$1; local args=$2
  timeout 300 python bench.py $args --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_ab16.err | tail -1 > gpurun_out/r2_ab16_$label.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_ab16_$label.json')); c=d['roofline']['classes']
print('$label', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, d['final_loss'])
print('   ', {k: round(v['ms'],2) for k,v in sorted(c.items(), key=lambda kv:-kv[1]['ms'])})" 2>&1 | tail -2
  tail -2 gpurun_out/r2_ab16.err | cut -c1-300
}
run rn50_1 ""
run mbv2 "--model mobilenet_v2 --batch 512 --no-e2e"
run resnext "--model resnext --depth 101 --batch 128 --no-e2e"
run rn50_2 ""
timeout 120 python tools/sgd_bench.py 2>&1 | tail -4
