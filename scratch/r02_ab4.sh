#!/bin/bash
# One GPU session: the whole -m gpu suite on the working-tree library, then A/B timing of tuning builds
#   gpurun --timeout 700 -- 'bash scratch/r02_ab4.sh <tag> "bunny:128:300 crates:8:100 sprites:64:100 small_tris:4:100" v1 v2 ...'
tag=$1; WL="$2"; shift 2
mkdir -p gpurun_out
( timeout 60 python __graft_entry__.py --smoke; echo "smoke rc $?"
  timeout 300 python -m pytest tests -m gpu -q -rf 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.txt 2>&1
bash scratch/ab.sh "$WL" "$@" > gpurun_out/${tag}_ab.txt 2>&1
timeout 120 python bench.py --workload crates --frames 128 --steps 5 --kernel-only 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('crates128', round(j['ms_per_step'], 4), 'ms/step', round(j['frames_per_s']), 'fps')
" >> gpurun_out/${tag}_ab.txt
tail -6 gpurun_out/${tag}_tests.txt; cat gpurun_out/${tag}_ab.txt
