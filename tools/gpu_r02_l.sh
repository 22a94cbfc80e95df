#!/bin/bash
# round 2, call l (2 GPUs): the final code on the sharded path -- world-2 tests (P2P transport, C++ API on 2 ranks)
# and the default bench line at N=2
set -u
mkdir -p gpurun_out
O=gpurun_out/r02l
N=2
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_api.py tests/test_gpu_slices.py -m gpu -x -q \
  -k "2-p2p or (ingest and p2p) or (complex and p2p) or (several_ranks and 2) or two_ranks or slices" > ${O}_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest_2gpu.log; tail -4 ${O}_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 > ${O}_bench_c3_n2.json 2> ${O}_bench_c3_n2.err; echo "bench rc=$?"; tail -c 300 ${O}_bench_c3_n2.err | tail -2
