#!/bin/bash
# round 2, call d (2 GPUs): sharded engine on the current code -- GPU suite (multi-GPU cases run), bench c3 at N=2
# with both transports, C++ API on two ranks
set -u
mkdir -p gpurun_out
O=gpurun_out/r02d
nvidia-smi --query-gpu=name,memory.total --format=csv > ${O}_box.txt; nvidia-smi topo -m >> ${O}_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_2gpu.log; tail -6 ${O}_pytest_2gpu.log
run_bench() { # tag, extra args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 "$@" > ${O}_bench_$tag.json 2> ${O}_bench_$tag.err; echo "bench $tag rc=$?"; tail -c 300 ${O}_bench_$tag.err
}
run_bench c3_n2
run_bench c3_n2_nccl --transport 1 --no-e2e
run_bench c2_n2 --config c2 --no-e2e
