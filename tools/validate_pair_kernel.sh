#!/bin/bash
# One-call validation of the opt-in CTA-pair kernel (csrc/conv_pair.cu), meant to be the FIRST gpurun call of the next
# round:   gpurun --timeout 900 -- 'bash tools/validate_pair_kernel.sh'   ->  gpurun_out/pair_validation.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/pair_validation.log
: > $LOG
echo "== probe (cta_group::2 semantics)" >> $LOG
timeout 60 python tools/pair_probe.py >> $LOG 2>&1
echo "== conv diag, wide 1x1 cases, B200_IGEMM_PAIR=1" >> $LOG
for c in p1_1024_256_14 p1_512_2048_7 p1_256_128_ragged p1_512_512_res p1_256_1024_14; do
  B200_IGEMM_PAIR=1 timeout 90 python tools/conv_diag.py $c 2>&1 | grep -E "DIAG|rror" | tail -2 | cut -c1-500 >> $LOG
  echo "exit=$? case=$c" >> $LOG
done
echo "== fused BN statistics + engine parity with the pair kernel" >> $LOG
B200_IGEMM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -q -x \
  -k "fused_bn or resnet50_imagenet_against or eval_with_folded or graph" 2>&1 | tail -3 >> $LOG
echo "== layer bench 1x1: default vs pair" >> $LOG
for v in 0 1; do
  echo "-- B200_IGEMM_PAIR=$v" >> $LOG
  B200_IGEMM_PAIR=$v timeout 200 python tools/layer_bench.py 1x1 2>&1 | grep -E "l3_|l4_|l2_1x1_512|l2_1x1_128" | cut -c1-330 >> $LOG
done
echo "== bench A/B" >> $LOG
for v in 0 1 0 1; do
  B200_IGEMM_PAIR=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pair=$v', d['value'], d['ms_per_step'], d['final_loss'])" >> $LOG 2>&1
done
cat $LOG
