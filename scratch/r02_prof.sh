#!/bin/bash
# Source-level ncu capture of k_raster (first launch of the timed region) + A/B of tuning variants in one session.
#   gpurun --timeout 900 -- 'bash scratch/r02_prof.sh <tag> "<workloads>" v1 v2 ...'
tag=$1; WL="$2"; shift 2
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:k_raster' -c 1 \
  -f -o gpurun_out/${tag}_raster python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?"
bash scratch/ab.sh "$WL" "$@" | tee gpurun_out/${tag}_ab.txt
