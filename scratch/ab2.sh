#!/bin/bash
# One GPU session: parity of the working-tree library on the device, then A/B timing of the tuning builds
# (retrofire_b200/_variants/*.so: base = previous commit, r = rasteriser changes only, rs = rasteriser + staged k_setup stores).
mkdir -p gpurun_out
( timeout 60 python __graft_entry__.py --smoke; echo "smoke rc $?"
  timeout 100 python -m pytest tests/test_gpu_1_configs.py tests/test_gpu_2_api.py tests/test_gpu_3_adversarial.py -x -q -k "bunny_x16 or sprites_10k or random_soup or crates_169 or hello_tri or indexed or line_prim or fuzz or heaviest or deep_tile or context_flags" 2>&1 | tail -3 ) > gpurun_out/ab_tests.txt 2>&1
one() {  # variant workload frames steps
  RF_B200_LIB=$PWD/retrofire_b200/_variants/$1.so timeout 60 python bench.py --workload $2 --frames $3 --steps $4 --kernel-only 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$1', '$2', round(j['ms_per_step'], 4), 'ms/step', round(j['frames_per_s']), 'fps', 'raster_ms', round(j['roofline']['kernel_ms_avg'], 4), json.dumps(j['roofline']['kernel_time_share']))
"
}
for rep in 1 2; do for v in base r rs; do one $v bunny 128 200; done; done > gpurun_out/ab.txt 2>&1
for v in base rs; do one $v crates 8 60; done >> gpurun_out/ab.txt 2>&1
for v in base rs; do one $v sprites 64 60; done >> gpurun_out/ab.txt 2>&1
cat gpurun_out/ab_tests.txt gpurun_out/ab.txt
