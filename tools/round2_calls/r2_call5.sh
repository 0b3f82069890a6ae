#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest5.log 2>&1
grep -E "^F?\.*T1|^F?\.*T2|T1 logits|T2 logits|MBv2|duplicates|adapt_grad|full-size|^   [a-z0-9.]+ +cos|uint8|evaluate:|passed|failed|^FAILED" gpurun_out/r2_pytest5.log | cut -c1-420
for kt in 4 1; do B200_WGRAD_KT=$kt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench5_$kt.err | tail -1 > gpurun_out/r2_bench5_$kt.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench5_$kt.json')); print('kt=$kt', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['windows_ms_per_step'], 'u8', d['e2e'].get('uint8_input'), 'loss', d['final_loss'], 'h2d', round(d['e2e']['h2d_gbs_measured'],1)); print({k:v['ms'] for k,v in d['roofline']['classes'].items()})"; done
