#!/bin/bash
# First GPU call of round 2 (≈ 8 GPU-minutes): the parity suite with the tests added at the end of round 1 (lattice ties, bands of
# any height, find_draw combinations — so far only run under emulation), the default bench line of the committed tree (the
# find_draw division of k_vertex/k_assemble is unmeasured), the launch list, DRAM bytes of every kernel of one step (to check the
# traffic table of DESIGN §9), and source-level captures of k_assemble and k_bin_sort_warp (none yet).
#   gpurun --timeout 900 -- 'bash scratch/r02_first.sh'
mkdir -p gpurun_out
( timeout 60 python __graft_entry__.py --smoke; echo "smoke rc $?"
  timeout 600 python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 ) > gpurun_out/r02_tests.txt 2>&1
timeout 120 python bench.py --steps 300 --cpu-seconds 6 > gpurun_out/r02_bench_bunny_1gpu.json 2> gpurun_out/r02_bench.err; echo "bench rc $?"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_ -c 48 --csv \
  --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 --kernel-only > /dev/null 2>&1; echo "launch list rc $?"
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum \
  --clock-control none --profile-from-start off -k regex:k_ -c 14 --csv --log-file gpurun_out/r02_dram_per_kernel.csv \
  python bench.py --steps 1 --warmup 3 --kernel-only > /dev/null 2>&1; echo "dram rc $?"
timeout 170 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:k_assemble|k_bin_sort_warp' -c 2 \
  -f -o gpurun_out/r02_assemble_sort python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/r02_ncu_hot.log 2>&1; echo "capture rc $?"
cat gpurun_out/r02_tests.txt; tail -c 700 gpurun_out/r02_bench_bunny_1gpu.json
