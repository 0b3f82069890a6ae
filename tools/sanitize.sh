#!/bin/bash
# compute-sanitizer over the convolution / BN op tests: memcheck, then racecheck.  Summaries go to
# gpurun_out/<tag>_sanitizer_{memcheck,racecheck}.log (copy the tails into profiles/).  usage: tools/sanitize.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
SEL='test_conv_fprop_dgrad_wgrad or test_fused_bn_statistics or test_bn_forward_backward or test_stem_bn_relu_maxpool_fused or test_grouped_conv_window_mode'
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
    python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL" > $OUT/${TAG}_sanitizer_$tool.full.log 2>&1
  echo "rc=$?" >> $OUT/${TAG}_sanitizer_$tool.full.log
  { echo "# compute-sanitizer --tool $tool, tests/test_gpu_ops.py -k \"$SEL\""; \
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Error|hazard" $OUT/${TAG}_sanitizer_$tool.full.log | sort | uniq -c | sort -rn | head -40; } \
    > $OUT/${TAG}_sanitizer_$tool.log
  tail -5 $OUT/${TAG}_sanitizer_$tool.log
done
