#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_parity.py -m gpu -q -s -k "conv_fprop_dgrad or fused_bn or resnet50_imagenet_against or resnet20 or graph or full_size or grouped or mobilenet_v1 or eval_with_folded" > gpurun_out/r2_pytest10.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_pytest10.log | cut -c1-300 | head -20
for v in 0 1; do echo "-- B200_IGEMM_EPI2=$v"; B200_IGEMM_EPI2=$v timeout 300 python tools/layer_bench.py l1_1x1 2>&1 | cut -c1-330; B200_IGEMM_EPI2=$v timeout 200 python tools/layer_bench.py l2_1x1 2>&1 | cut -c1-330; done > gpurun_out/r2_epi2_layers.log 2>&1
cat gpurun_out/r2_epi2_layers.log
for v in 0 1 0 1; do B200_IGEMM_EPI2=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('epi2=$v', round(d['value']), round(d['ms_per_step'],3), d['final_loss'], {k:v['ms'] for k,v in d['roofline']['classes'].items() if k.startswith('conv')})"; done 2>&1 | tee gpurun_out/r2_epi2_bench.log
