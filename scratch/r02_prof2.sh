#!/bin/bash
# ncu capture of arbitrary kernels + A/B:  bash scratch/r02_prof2.sh <tag> <kernel regex> <count> "<workloads>" v1 v2 ...
tag=$1; KR="$2"; CNT=$3; WL="$4"; shift 4
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KR" -c $CNT \
  -f -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?"
bash scratch/ab.sh "$WL" "$@" | tee gpurun_out/${tag}_ab.txt
