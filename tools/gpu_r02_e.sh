#!/bin/bash
# round 2, call e (1 GPU): class chain (one Tijk cube per tuple) -- parity suite, whole-run timings, ncu
set -u
mkdir -p gpurun_out
O=gpurun_out/r02e
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zcomplex.py tests/test_gpu_zzz_experimental.py tests/test_gpu_host_api.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
echo "== c2" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py 40 400 4200 >> ${O}_ab.log 2>&1
echo "== c2 sync reduce" >> ${O}_ab.log; ATRIP_B200_REDUCE=sync timeout 300 python tools/dev_perf.py 40 400 4200 >> ${O}_ab.log 2>&1
echo "== c3" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py 64 640 1200 >> ${O}_ab.log 2>&1
echo "== c5-shaped (No=32 Nv=480)" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py 32 480 8000 >> ${O}_ab.log 2>&1
echo "== c4 shapes solo" >> ${O}_ab.log; timeout 600 python tools/dev_perf_solo.py 100 1000 8 238 >> ${O}_ab.log 2>&1
grep -E "^==|run |rror" ${O}_ab.log
tools/prof.sh r02e_c2 40 400 3200
PROF_CMD="python tools/dev_perf_solo.py 100 1000 8 238" tools/prof.sh r02e_c4
