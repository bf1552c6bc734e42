#!/bin/bash
# One GPU iteration of round 2: parity suite, default bench line, launch list, DRAM bytes per kernel of one step.
#   gpurun --timeout 900 -- 'bash scratch/r02_step.sh <tag>'
tag=${1:-step}
mkdir -p gpurun_out
( timeout 60 python __graft_entry__.py --smoke; echo "smoke rc $?"
  timeout 600 python -m pytest tests -m gpu -q -rf 2>&1 | tail -30 ) > gpurun_out/${tag}_tests.txt 2>&1
timeout 150 python bench.py --steps 100 --cpu-seconds 4 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_ -c 52 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --kernel-only > /dev/null 2>&1; echo "launch list rc $?"
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum \
  --clock-control none --profile-from-start off -k regex:k_ -c 14 --csv --log-file gpurun_out/${tag}_dram_per_kernel.csv \
  python bench.py --steps 1 --warmup 3 --kernel-only > /dev/null 2>&1; echo "dram rc $?"
for w in crates sprites small_tris; do timeout 100 python bench.py --workload $w --steps 20 --cpu-seconds 2 --kernel-only > gpurun_out/${tag}_bench_$w.json 2>> gpurun_out/${tag}_bench.err; echo "$w rc $?"; done
tail -12 gpurun_out/${tag}_tests.txt; tail -c 1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
