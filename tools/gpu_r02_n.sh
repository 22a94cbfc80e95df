#!/bin/bash
# round 2, call n (1 GPU, last minutes): ncu captures of the c5 kernel shapes (No=32 Nv=1200) as rank 0 of 4
set -u
mkdir -p gpurun_out
PROF_CMD="python tools/dev_perf_solo.py 32 1200 4 3552" tools/prof.sh r02n_c5
