#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for c in g32_128_56 g32_256_28s2 g32_512_14 g32_1024_7 g8_256_20x12; do timeout 90 python tools/conv_diag.py $c 2>&1 | grep DIAG | cut -c1-600; done > gpurun_out/r2_grouped_diag.log 2>&1
cat gpurun_out/r2_grouped_diag.log
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -q -s -k "grouped or resnext or mobilenet or conv_fprop_dgrad or softmax or squeeze or folded" > gpurun_out/r2_pytest8.log 2>&1
grep -E "T1 |T2 |MobileNet|passed|failed|^FAILED|^E  " gpurun_out/r2_pytest8.log | cut -c1-400 | head -60
timeout 300 python bench.py --model resnext --depth 101 --batch 128 --steps 10 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/r2_bench8_rnx.err | tail -1 > gpurun_out/r2_bench8_rnx.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench8_rnx.json')); print('resnext101', round(d['value']), round(d['ms_per_step'],3), d['final_loss']); print({k:v['ms'] for k,v in d['roofline']['classes'].items()})"
timeout 300 python bench.py --model mobilenet_v2 --batch 512 --steps 10 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/r2_bench8_mb.err | tail -1 > gpurun_out/r2_bench8_mb.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench8_mb.json')); print('mobilenet_v2', round(d['value']), round(d['ms_per_step'],3), d['final_loss']); print({k:v['ms'] for k,v in d['roofline']['classes'].items()})"
