#!/bin/bash
# Runs every conv diagnostic case in its own process with a timeout; writes gpurun_out/conv_diag.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/conv_diag.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version --format=csv >> $LOG 2>&1
for c in $(python tools/conv_diag.py); do
  if [ -n "$1" ] && [[ "$c" != $1 ]]; then continue; fi
  timeout 90 python tools/conv_diag.py $c 2>&1 | grep -E "DIAG|Error|error" | tail -3 >> $LOG
  echo "exit=$? case=$c" >> $LOG
done
cat $LOG
