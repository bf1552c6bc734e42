#!/bin/bash
# One short GPU session: source-level ncu capture of the two heaviest kernels of the headline workload (first k_setup and
# first k_raster of the timed region), then the default bench line. Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:k_raster|k_setup' -c 2 \
  -f -o gpurun_out/r01_hot python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/ncu_hot.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_hot.log
timeout 120 python bench.py --steps 300 --cpu-seconds 6 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"
tail -c 600 gpurun_out/bench_final.json
