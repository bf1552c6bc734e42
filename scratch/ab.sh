#!/bin/bash
# A/B timing of tuning builds (retrofire_b200/_variants/*.so, see build.build_variant) inside ONE GPU session:
#   gpurun -- bash scratch/ab.sh "bunny:128:300 crates:8:100" base avg4 ...
WL="$1"; shift
for rep in 1 2; do
  for v in "$@"; do
    for w in $WL; do
      IFS=: read n f s <<< "$w"
      RF_B200_LIB=$PWD/retrofire_b200/_variants/$v.so python bench.py --workload $n --frames $f --steps $s --kernel-only 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$v', '$n', round(j['ms_per_step'], 4), 'ms/step', round(j['frames_per_s']), 'fps', 'raster', round(j['roofline']['kernel_ms_avg'], 4), 'shares', {k[2:]: v for k, v in j['roofline']['kernel_time_share'].items() if v >= 0.02})
"
    done
  done
done
