#!/bin/bash
# final code: full GPU suite, contract bench (copied to profiles/r02_bench_default.json), smoke
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest20.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest20.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02_smoke.log 2>&1
tail -2 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().split('\n')[-1]); r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e']['uint8_input']['value']), 'frac', round(r['frac'],3), 'conv_all', round(r['conv_all']['frac_of_tensor_peak'],3), 'hbm_all', round(r['hbm_all']['frac_of_hbm_peak'],3), 'traffic', r['traffic'], d['clocks'])
print({k: round(v['ms'],2) for k,v in sorted(r['classes'].items(), key=lambda kv:-kv[1]['ms'])})"
for m in "--model resnext --depth 101 --batch 128" "--model mobilenet_v2 --batch 512"; do
  timeout 300 python bench.py $m --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$m', round(d['value']), round(d['ms_per_step'],3), d['final_loss'])" 2>&1 | tail -1
done
