#!/bin/bash
cd "$(dirname "$0")/../.."
L=gpurun_out/r2_wgrad_debug2.log; : > $L
for lib in gpurun_tmp/old/libb200conv.so convnet/pytorch_b200/libb200conv.so; do
  for c in c3_64_64 c3s2_128_128_56 stem_s2d; do
    echo "-- $lib $c" >> $L
    B200_LIB_PATH=$PWD/$lib B200_DIAG_TAPS=1 B200_WGRAD_DEBUG=1 timeout 90 python tools/conv_diag.py $c 2>&1 | grep -E "DIAG|wgrad cfg" | cut -c1-900 >> $L
  done
done
cat $L
