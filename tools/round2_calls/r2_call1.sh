#!/bin/bash
# round-2 call 1: full GPU suite (calibrates the survey bounds), ncu single-CTA vs CTA-pair on layer3/4 1x1, quick bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$" | tail -150 > gpurun_out/r2_pytest1.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:conv_igemm -o gpurun_out/r2_single_l3_256_1024 python tools/layer_bench.py l3_1x1_256_1024 fprop --once > gpurun_out/r2_ncu1.log 2>&1
B200_IGEMM_PAIR=1 timeout 200 $NCU -k regex:conv_pair -o gpurun_out/r2_pair_l3_256_1024 python tools/layer_bench.py l3_1x1_256_1024 fprop --once > gpurun_out/r2_ncu2.log 2>&1
B200_IGEMM_PAIR=1 timeout 200 $NCU -k regex:conv_pair -o gpurun_out/r2_pair_l4_2048_512 python tools/layer_bench.py l4_1x1_2048_512 fprop --once > gpurun_out/r2_ncu3.log 2>&1
timeout 200 $NCU -k regex:conv_wgrad_kernel -o gpurun_out/r2_wgrad_l3_256_1024 python tools/layer_bench.py l3_1x1_256_1024 wgrad --once > gpurun_out/r2_ncu4.log 2>&1
timeout 200 $NCU -k regex:conv_wgrad_kernel -o gpurun_out/r2_wgrad_l1_64_256 python tools/layer_bench.py l1_1x1_64_256 wgrad --once > gpurun_out/r2_ncu5.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 1500 gpurun_out/r2_bench1.json
grep -E "passed|failed" gpurun_out/r2_pytest1.log | tail -3
ls -la gpurun_out | grep r2_
