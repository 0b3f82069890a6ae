#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python tools/parity_diag.py 50 224 32 oracle > gpurun_out/r2_parity_50_224.log 2>&1
timeout 600 python tools/parity_diag.py 50 64 32 oracle > gpurun_out/r2_parity_50_64.log 2>&1
timeout 600 python tools/parity_diag.py 18 128 32 oracle > gpurun_out/r2_parity_18_128.log 2>&1
tail -45 gpurun_out/r2_parity_50_224.log gpurun_out/r2_parity_50_64.log gpurun_out/r2_parity_18_128.log | grep -v Warning
