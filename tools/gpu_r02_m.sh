#!/bin/bash
# round 2, call m (8 GPUs, last GPU minutes of the round): c4 on the final code (value + parity, no e2e leg), then the
# 8-GPU c4 tuple test of tests/test_gpu_multi.py
set -u
mkdir -p gpurun_out
O=gpurun_out/r02m
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --no-e2e --steps 5 > ${O}_bench_c4_n8.json 2> ${O}_bench_c4_n8.err; echo "bench rc=$?"; tail -c 200 ${O}_bench_c4_n8.err | tail -1
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k c4 > ${O}_pytest_c4_8gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_c4_8gpu.log; tail -3 ${O}_pytest_c4_8gpu.log
