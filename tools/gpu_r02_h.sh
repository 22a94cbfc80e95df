#!/bin/bash
# round 2, call h (8 GPUs): the metric's configuration c4 (No=100 Nv=1000) with in-bench parity and e2e, c2 (scaling
# limiter), c5, and the NCCL transport on c4
set -u
mkdir -p gpurun_out
O=gpurun_out/r02h
N=8
nvidia-smi --query-gpu=name,memory.total --format=csv > ${O}_box.txt; free -g >> ${O}_box.txt; nproc >> ${O}_box.txt
run_bench() { # tag, extra args
  tag=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > ${O}_bench_$tag.json 2> ${O}_bench_$tag.err; echo "bench $tag rc=$?"; tail -c 300 ${O}_bench_$tag.err | tail -2
}
run_bench c4_n8
run_bench c2_n8 --config c2 --no-e2e
run_bench c5_n8 --config c5 --no-e2e --steps 6
run_bench c4_n8_nccl --transport 1 --no-e2e --no-parity --steps 5
