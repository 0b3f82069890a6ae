#!/bin/bash
# end-of-round evidence on one GPU: smoke(), launch list of one step (time + DRAM bytes per launch), contract bench,
# reference arm.  Outputs under gpurun_out/ (copy into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02_smoke.log 2>&1
tail -2 gpurun_out/r02_smoke.log
B200_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 1 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv gpurun_out/r02_launches.md "r02 final" gpurun_out/r02_traffic.json | tail -3
cp gpurun_out/r02_traffic.json profiles/r02_traffic.json      # the contract run below reports roofline.traffic from it
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -c 2500 gpurun_out/r02_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -c 700 gpurun_out/r02_bench_reference.json
