#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest7.log 2>&1
grep -E "T1 |T2 |MBv1|MBv2|duplicates|adapt_grad|full-size|^   [a-z0-9.]+ +cos|uint8|evaluate:|passed|failed|^FAILED|^E  " gpurun_out/r2_pytest7.log | cut -c1-520 | head -120
for kt in 4 1; do B200_WGRAD_KT=$kt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench7_$kt.err | tail -1 > gpurun_out/r2_bench7_$kt.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench7_$kt.json')); print('kt=$kt', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['windows_ms_per_step'], 'u8', round(d['e2e']['uint8_input'].get('value',0)), 'loss', d['final_loss']); print({k:v['ms'] for k,v in d['roofline']['classes'].items()})"; done
timeout 200 python tools/layer_bench.py "" wgrad 2>&1 | cut -c1-120 > gpurun_out/r2_layer_wgrad7.log; cat gpurun_out/r2_layer_wgrad7.log
