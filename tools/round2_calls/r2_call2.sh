#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in A B C D; do timeout 120 python tools/capture_diag.py $v 2>&1 | grep CAPTURE; done > gpurun_out/r2_capture_diag.log
cat gpurun_out/r2_capture_diag.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest2.log 2>&1
grep -E "^T1|^T2|MBv2|duplicates|adapt_grad|full-size|^   [a-z0-9.]+ +cos|passed|failed|^FAILED" gpurun_out/r2_pytest2.log | cut -c1-400
for kt in 1 4; do echo "-- B200_WGRAD_KT=$kt"; B200_WGRAD_KT=$kt timeout 300 python tools/layer_bench.py 1x1 wgrad 2>&1 | cut -c1-200; B200_WGRAD_KT=$kt timeout 100 python tools/layer_bench.py ds_ wgrad 2>&1 | cut -c1-200; done > gpurun_out/r2_wgrad_kt.log 2>&1
cat gpurun_out/r2_wgrad_kt.log
for kt in 1 4 1 4; do B200_WGRAD_KT=$kt timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kt=$kt', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['e2e']['windows_ms_per_step'], d['final_loss'], d['e2e']['host_enqueue_ms_per_step'])"; done 2>&1 | tee gpurun_out/r2_bench_kt.log
