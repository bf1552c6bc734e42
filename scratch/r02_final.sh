#!/bin/bash
# Record session of round 2 (one GPU): smoke + the whole -m gpu suite, the default bench line as the driver runs it, the
# reference arm, launch list, DRAM bytes per kernel of one step (bunny and crates), a source-level capture of k_raster,
# the pass timeline, and the other workloads.
#   gpurun --timeout 1100 -- 'bash scratch/r02_final.sh <tag>'
tag=${1:-final}
mkdir -p gpurun_out
( timeout 60 python __graft_entry__.py --smoke; echo "smoke rc $?"
  timeout 600 python -m pytest tests -m gpu -q -rf 2>&1 | tail -30 ) > gpurun_out/${tag}_tests.txt 2>&1
timeout 240 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference rc $?"
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_ -c 52 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --kernel-only > /dev/null 2>&1; echo "launch list rc $?"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum
timeout 120 ncu --metrics $M --clock-control none --profile-from-start off -k regex:k_ -c 14 --csv --log-file gpurun_out/${tag}_dram_per_kernel.csv \
  python bench.py --steps 1 --warmup 3 --kernel-only > /dev/null 2>&1; echo "dram rc $?"
timeout 120 ncu --metrics $M --clock-control none --profile-from-start off -k regex:k_ -c 16 --csv --log-file gpurun_out/${tag}_dram_per_kernel_crates.csv \
  python bench.py --workload crates --frames 8 --steps 1 --warmup 3 --kernel-only > /dev/null 2>&1; echo "dram crates rc $?"
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:k_raster' -c 1 \
  -f -o gpurun_out/${tag}_raster python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/${tag}_ncu.log 2>&1; echo "capture rc $?"
RF_DEBUG_PASS=1 timeout 100 python bench.py --steps 6 --warmup 3 --kernel-only > /dev/null 2> gpurun_out/${tag}_timeline.txt; echo "timeline rc $?"
for w in crates sprites small_tris; do timeout 100 python bench.py --workload $w --steps 50 --cpu-seconds 2 > gpurun_out/${tag}_bench_$w.json 2>> gpurun_out/${tag}_bench.err; echo "$w rc $?"; done
tail -12 gpurun_out/${tag}_tests.txt; tail -c 1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
