#!/bin/bash
# round 2, call g (4 GPUs): sharded engine at world 4 (both transports, complex, ingest, C++ API on 4 ranks),
# bench c3 and c5 at N=4
set -u
mkdir -p gpurun_out
O=gpurun_out/r02g
N=4
nvidia-smi --query-gpu=name,memory.total --format=csv > ${O}_box.txt; free -g >> ${O}_box.txt; nproc >> ${O}_box.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_api.py -m gpu -x -q \
  -k "4-nccl or 4-p2p or (complex and p2p) or (ingest and p2p) or (several_ranks and 4)" > ${O}_pytest_4gpu.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest_4gpu.log; tail -5 ${O}_pytest_4gpu.log
run_bench() { # tag, extra args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > ${O}_bench_$tag.json 2> ${O}_bench_$tag.err; echo "bench $tag rc=$?"; tail -c 300 ${O}_bench_$tag.err | tail -2
}
run_bench c3_n4
run_bench c5_n4 --config c5 --no-e2e
