#!/bin/bash
# round 2, call f (1 GPU): class chain with whole-round groups and mid-item publish -- timings, ncu
set -u
mkdir -p gpurun_out
O=gpurun_out/r02f
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -3 ${O}_pytest.log
echo "== c2" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py 40 400 4200 >> ${O}_ab.log 2>&1
echo "== c2 publish at chunk 1" >> ${O}_ab.log; ATRIP_B200_PUBLISH_CHUNK=1 timeout 300 python tools/dev_perf.py 40 400 4200 >> ${O}_ab.log 2>&1
echo "== c2 group 8 (no whole-round lag: waits in the epilogue)" >> ${O}_ab.log; ATRIP_B200_GROUP=8 timeout 300 python tools/dev_perf.py 40 400 4200 >> ${O}_ab.log 2>&1
echo "== c3" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py 64 640 1200 >> ${O}_ab.log 2>&1
echo "== c3 group 3" >> ${O}_ab.log; ATRIP_B200_GROUP=3 timeout 300 python tools/dev_perf.py 64 640 1200 >> ${O}_ab.log 2>&1
echo "== c4 shapes solo" >> ${O}_ab.log; timeout 600 python tools/dev_perf_solo.py 100 1000 8 234 >> ${O}_ab.log 2>&1
echo "== c4 shapes solo group 1" >> ${O}_ab.log; ATRIP_B200_GROUP=1 timeout 600 python tools/dev_perf_solo.py 100 1000 8 234 >> ${O}_ab.log 2>&1
grep -E "^==|run |rror" ${O}_ab.log
tools/prof.sh r02f_c2 40 400 3150
PROF_CMD="python tools/dev_perf_solo.py 100 1000 8 234" tools/prof.sh r02f_c4
