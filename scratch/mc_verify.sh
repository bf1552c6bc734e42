#!/bin/bash
# After the k_setup poison fix: the flow that failed (bunny batch, then a 32-frame crates batch in one context) under memcheck, three
# times; then the -m gpu suite.
for i in 1 2 3; do
  timeout 300 compute-sanitizer --tool memcheck --print-limit 2 python bench.py --frames 8 --steps 2 --warmup 3 --cpu-seconds 0.2 > gpurun_out/mcv_$i.log 2>&1
  echo "memcheck run $i rc=$? $(grep -a 'ERROR SUMMARY' gpurun_out/mcv_$i.log | head -1) $(grep -a 'RetrofireError:' gpurun_out/mcv_$i.log | head -1 | cut -c1-160)"
done
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
