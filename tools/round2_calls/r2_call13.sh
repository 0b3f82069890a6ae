#!/bin/bash
# after the 32-bit index fix: op tests + MobileNet-v1, then A/B of stem fusion / priority / masks on the three models
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -q -s -k "bn_ or maxpool or stem or mobilenet_v1 or resnet50_imagenet or resnet18" > gpurun_out/r2_pytest13.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest13.log | cut -c1-300
grep -E "  unit features|teacher-forced" gpurun_out/r2_pytest13.log | cut -c1-300 | head
run() {   # label, bench args, env assignments...
  local label=$1; local args=$2; shift; shift
  env "$@" timeout 300 python bench.py $args --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/r2_ab13.err | tail -1 > gpurun_out/r2_ab13_$label.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_ab13_$label.json')); print('$label', round(d['value']), round(d['ms_per_step'],3), d['final_loss'])" 2>&1 | tail -1
  tail -2 gpurun_out/r2_ab13.err | cut -c1-300
}
BASE="B200_PDL=0 B200_FUSE_STEM_POOL=0 B200_MAIN_PRIORITY=0 B200_BN_ACT_MASK=0"
for rep in 1 2; do
  run rn50_base_$rep "" $BASE
  run rn50_stem_$rep "" $BASE B200_FUSE_STEM_POOL=1
  run rn50_all3_$rep "" $BASE B200_FUSE_STEM_POOL=1 B200_MAIN_PRIORITY=1 B200_BN_ACT_MASK=1
  run rn50_prio_mask_$rep "" $BASE B200_MAIN_PRIORITY=1 B200_BN_ACT_MASK=1
done
for m in "resnext:--model resnext --depth 101 --batch 128" "mbv2:--model mobilenet_v2 --batch 512"; do
  name=${m%%:*}; args=${m#*:}
  run ${name}_base "$args" $BASE
  run ${name}_prio "$args" $BASE B200_MAIN_PRIORITY=1
  run ${name}_mask "$args" $BASE B200_BN_ACT_MASK=1
  run ${name}_pdl "$args" $BASE B200_PDL=1
  run ${name}_all3 "$args" $BASE B200_FUSE_STEM_POOL=1 B200_MAIN_PRIORITY=1 B200_BN_ACT_MASK=1
done
