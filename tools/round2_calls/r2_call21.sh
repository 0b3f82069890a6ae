#!/bin/bash
# ring-buffered host->device prefetch: Trainer tests + the contract bench twice (e2e windows)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -m gpu -q -x -k "trainer or Trainer or lazy or uint8 or regime or duplicates or evaluate or fused_ce or resnet20" > gpurun_out/r2_pytest21.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest21.log | cut -c1-300
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_e2e_$i.json 2> gpurun_out/r02_bench_e2e_$i.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_e2e_$i.json').read().strip().split('\n')[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['windows_ms_per_step'], 'u8', round(d['e2e']['uint8_input']['value']))"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().split('\n')[-1])
print('contract: value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['windows_ms_per_step'], 'u8', round(d['e2e']['uint8_input']['value']), d['cpu_baseline'])"
