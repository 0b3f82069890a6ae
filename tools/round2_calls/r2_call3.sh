#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2_wgrad_debug.log; : > $L
for kt in 1 4; do
  for c in c3_64_64 res_relu c3s2_128_128_56 c3_512_512_7 stem_s2d halo_8_256 halo_stem_67 p1_512_2048_7; do
    echo "-- KT=$kt $c" >> $L
    B200_WGRAD_KT=$kt timeout 90 python tools/conv_diag.py $c 2>&1 | grep DIAG | python -c "import sys,json; d=json.loads(sys.stdin.read()[5:]); print({k:(round(v[0],6) if isinstance(v,list) else v) for k,v in d.items()})" >> $L 2>&1
  done
done
echo "-- sanitizer memcheck res_relu" >> $L
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/conv_diag.py res_relu 2>&1 | grep -E "ERROR SUMMARY|Invalid|DIAG|at 0x|by thread" | head -20 >> $L
echo "-- sanitizer racecheck res_relu" >> $L
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python tools/conv_diag.py res_relu 2>&1 | grep -E "RACECHECK SUMMARY|hazard|DIAG|ERROR" | head -20 >> $L
cat $L
