#!/bin/bash
# Round-2 profiling pass (one GPU): (1) ncu launch list of one eager bench step with device time AND DRAM bytes per launch,
# (2) `ncu --set full` captures of the kernels the summary cites.  usage: tools/profile_round2.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
B200_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 4 --no-e2e --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:conv_wgrad_kernel -o $OUT/${TAG}_wgrad_l3_256_1024 python tools/layer_bench.py l3_1x1_256_1024 wgrad --once > $OUT/${TAG}_ncu_a.log 2>&1
timeout 200 $NCU -k regex:conv_wgrad_kernel -o $OUT/${TAG}_wgrad_l1_64_256 python tools/layer_bench.py l1_1x1_64_256 wgrad --once > $OUT/${TAG}_ncu_b.log 2>&1
timeout 200 $NCU -k regex:conv_igemm -o $OUT/${TAG}_igemm_l3_256_1024 python tools/layer_bench.py l3_1x1_256_1024 fprop --once > $OUT/${TAG}_ncu_c.log 2>&1
timeout 200 $NCU -k regex:conv_halo_kernel -o $OUT/${TAG}_halo_l2_3x3 python tools/layer_bench.py l2_3x3_128_128 fprop --once > $OUT/${TAG}_ncu_d.log 2>&1
timeout 300 $NCU -k regex:bn_ -o $OUT/${TAG}_bn python tools/ncu_bn_once.py > $OUT/${TAG}_ncu_e.log 2>&1
ls -la $OUT | grep $TAG
