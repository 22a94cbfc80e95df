#!/bin/bash
# round 2, call i (1 GPU): 256-thread bulk-copy reduction -- parity, timings, ncu of the reduction on c2 / c3 / c4 shapes
set -u
mkdir -p gpurun_out
O=gpurun_out/r02i
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzz_experimental.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -3 ${O}_pytest.log
echo "== c2" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py 40 400 4200 >> ${O}_ab.log 2>&1
echo "== c3" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py 64 640 1020 >> ${O}_ab.log 2>&1
echo "== c4 shapes solo" >> ${O}_ab.log; timeout 600 python tools/dev_perf_solo.py 100 1000 8 156 >> ${O}_ab.log 2>&1
grep -E "^==|run |rror" ${O}_ab.log
tools/prof.sh r02i_c2 40 400 2100
PROF_CMD="python tools/dev_perf_solo.py 100 1000 8 117" tools/prof.sh r02i_c4
