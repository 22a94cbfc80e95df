#!/bin/bash
# round 2, call c (1 GPU): setmaxnreg contraction variants (parity + A/B timing), reduction variants, c4 kernel shapes solo
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zcomplex.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -3 ${O}_pytest.log
for cfg in "40 400 4200" "64 640 1020"; do
  set -- $cfg
  echo "== $1 $2 setmaxnreg variants" >> ${O}_ab.log; timeout 300 python tools/dev_perf.py $1 $2 $3 >> ${O}_ab.log 2>&1
  echo "== $1 $2 uniform registers (ATRIP_B200_NO_SETMAXNREG=1)" >> ${O}_ab.log; ATRIP_B200_NO_SETMAXNREG=1 timeout 300 python tools/dev_perf.py $1 $2 $3 >> ${O}_ab.log 2>&1
  echo "== $1 $2 reduce async" >> ${O}_ab.log; ATRIP_B200_REDUCE=async timeout 300 python tools/dev_perf.py $1 $2 $3 >> ${O}_ab.log 2>&1
  echo "== $1 $2 reduce async-rev" >> ${O}_ab.log; ATRIP_B200_REDUCE=async-rev timeout 300 python tools/dev_perf.py $1 $2 $3 >> ${O}_ab.log 2>&1
done
echo "== c4 shapes solo (rank 0 of 8), setmaxnreg" >> ${O}_ab.log; timeout 600 python tools/dev_perf_solo.py 100 1000 8 156 >> ${O}_ab.log 2>&1
echo "== c4 shapes solo, uniform registers" >> ${O}_ab.log; ATRIP_B200_NO_SETMAXNREG=1 timeout 600 python tools/dev_perf_solo.py 100 1000 8 156 >> ${O}_ab.log 2>&1
echo "== c4 shapes solo, reduce async" >> ${O}_ab.log; ATRIP_B200_REDUCE=async timeout 600 python tools/dev_perf_solo.py 100 1000 8 156 >> ${O}_ab.log 2>&1
cat ${O}_ab.log | grep -E "^==|run "
PROF_CMD="python tools/dev_perf_solo.py 100 1000 8 117" tools/prof.sh r02c_c4
tools/prof.sh r02c_c3 64 640 600
ATRIP_B200_REDUCE=async PROF_CMD="python tools/dev_perf.py 40 400 2100" tools/prof.sh r02c_c2async
