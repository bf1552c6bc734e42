#!/bin/bash
# Multi-GPU session: the sort-first tests, then the default bench line as the driver launches it at N ranks.
#   gpurun --gpus N --timeout 900 -- 'bash scratch/r02_multi.sh <tag> N'
tag=$1; N=$2
mkdir -p gpurun_out
( nvidia-smi -L; timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -12 ) > gpurun_out/${tag}_tests.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 3 \
  > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 1 \
  > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference rc $?"
cat gpurun_out/${tag}_tests.txt; tail -c 2500 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
