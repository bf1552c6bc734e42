#!/bin/bash
# Last GPU session of round 1: parity of the working-tree library, A/B of the second batch of changes (rs = previous commit,
# n1 = staged k_assemble output + staged k_setup input + one barrier less, n2 = n1 + L2 prefetch hints in k_raster),
# the default bench line and the ncu launch list of the final code.
mkdir -p gpurun_out
( timeout 40 python __graft_entry__.py --smoke; echo "smoke rc $?"
  timeout 60 python -m pytest tests/test_gpu_1_configs.py tests/test_gpu_2_api.py tests/test_gpu_3_adversarial.py -x -q -k "bunny_x16 or sprites_10k or random_soup or widest or crates_169 or hello_tri or indexed or line_prim or fuzz or heaviest or deep_tile or context_flags or depth_sort or growth" 2>&1 | tail -3 ) > gpurun_out/ab3_tests.txt 2>&1
one() {  # variant workload frames steps
  RF_B200_LIB=$PWD/retrofire_b200/_variants/$1.so timeout 40 python bench.py --workload $2 --frames $3 --steps $4 --kernel-only 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$1', '$2', round(j['ms_per_step'], 4), 'ms/step', round(j['frames_per_s']), 'fps', 'raster_ms', round(j['roofline']['kernel_ms_avg'], 4), json.dumps(j['roofline']['kernel_time_share']))
"
}
for rep in 1 2; do for v in rs n1 n2; do one $v bunny 128 200; done; done > gpurun_out/ab3.txt 2>&1
cat gpurun_out/ab3_tests.txt gpurun_out/ab3.txt
timeout 60 python bench.py --steps 300 --cpu-seconds 5 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc $?"
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_ -c 48 --csv --log-file gpurun_out/r01_final2_launches.csv python bench.py --steps 4 --warmup 3 --kernel-only > /dev/null 2>&1; echo "ncu rc $?"
