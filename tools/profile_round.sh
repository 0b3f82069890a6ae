#!/bin/bash
# Profiling pass of one round (run under gpurun on ONE GPU): ncu launch list of bench.py (eager launches, no graph,
# so that every kernel is listed once per step) + `--set full` captures of the top kernels on ResNet-50 layer shapes.
# usage: tools/profile_round.sh <tag>      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_*.ncu-rep
cd "$(dirname "$0")/.."
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
B200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 4 --no-e2e --no-cpu-baseline \
  > $OUT/${TAG}_launches_bench.log 2>&1
for L in l2_3x3_128_128 l3_1x1_256_1024 l1_1x1_256_64 stem_halo_4x4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ -o $OUT/${TAG}_conv_$L -f \
    python tools/layer_bench.py $L --once > $OUT/${TAG}_ncu_$L.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_ -o $OUT/${TAG}_bn -f \
  python tools/ncu_bn_once.py > $OUT/${TAG}_ncu_bn.log 2>&1
ls -la $OUT | grep $TAG
