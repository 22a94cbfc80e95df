#!/bin/bash
# GPU box: ncu evidence for the two hot kernels (one GPU; never under torchrun).
#   tools/prof.sh <tag> [No Nv tuples]      -> gpurun_out/<tag>_*.{csv,ncu-rep}
#   PROF_CMD='...' tools/prof.sh <tag>      profile another command (e.g. tools/dev_perf_solo.py for the c4 shapes)
set -u
TAG=${1:-prof}; NO=${2:-40}; NV=${3:-400}; NT=${4:-2100}
mkdir -p gpurun_out
CMD=${PROF_CMD:-"python tools/dev_perf.py $NO $NV $NT"}
# every launch with its device time (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
# the two kernels, full set, third launch of each
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 2 -c 1 \
    -f -o gpurun_out/${TAG}_contract $CMD > gpurun_out/${TAG}_contract.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:reduce_(async_)?kernel' -s 2 -c 1 \
    -f -o gpurun_out/${TAG}_reduce $CMD > gpurun_out/${TAG}_reduce.log 2>&1
ls -la gpurun_out/${TAG}_*
