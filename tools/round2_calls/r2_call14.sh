#!/bin/bash
# new defaults (row-quad masks, high-priority main chain, fused stem forward): full GPU suite, default bench, a few A/Bs,
# then the round-2 profiling pass
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest14.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest14.log | cut -c1-300
run() {   # label, bench args, env assignments...
  local label=$1; local args=$2; shift; shift
  env "$@" timeout 300 python bench.py $args --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_ab14.err | tail -1 > gpurun_out/r2_ab14_$label.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_ab14_$label.json')); print('$label', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, d['final_loss'])" 2>&1 | tail -1
  tail -2 gpurun_out/r2_ab14.err | cut -c1-300
}
run default_1 "" X=0
run stem0 "" B200_FUSE_STEM_POOL=0
run var1 "" B200_BN_BWD_VARIANT=1
run var2 "" B200_BN_BWD_VARIANT=2
run var3 "" B200_BN_BWD_VARIANT=3
run default_2 "" X=0
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench14_contract.json 2> gpurun_out/r2_bench14_contract.err
tail -c 1500 gpurun_out/r2_bench14_contract.json
bash tools/profile_round2.sh r02 > gpurun_out/r2_profile14.log 2>&1
tail -12 gpurun_out/r2_profile14.log
