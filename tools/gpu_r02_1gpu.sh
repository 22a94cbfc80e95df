#!/bin/bash
# round 2, 1-GPU call: GPU test suite, the default bench line (N=1 -> c3 if it fits), ncu evidence for c2 and c3
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/r02b_box.txt; free -g >> gpurun_out/r02b_box.txt; nproc >> gpurun_out/r02b_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_1gpu.log
tail -5 gpurun_out/r02b_pytest_1gpu.log
timeout 900 python bench.py > gpurun_out/r02b_bench_n1_default.json 2> gpurun_out/r02b_bench_n1_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02b_bench_n1_default.err
timeout 600 python bench.py --config c2 > gpurun_out/r02b_bench_n1_c2.json 2> gpurun_out/r02b_bench_n1_c2.err; echo "bench c2 rc=$?"
tools/prof.sh r02b_c2 40 400 2100
tools/prof.sh r02b_c3 64 640 600
