#!/bin/bash
# developer A/B: build a second copy of the device library with extra -D flags
#   tools/ab_build.sh <suffix> -DFOO ...   -> atrip_b200/csrc/libatrip_b200_<suffix>.so
#   ATRIP_B200_LIB=$PWD/atrip_b200/csrc/libatrip_b200_<suffix>.so python tools/dev_perf.py ...
set -e
S=$1; shift
cd "$(dirname "$0")/../atrip_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall \
  "$@" -shared -Xlinker -soname=libatrip_b200_$S.so -o libatrip_b200_$S.so engine.cu -ldl
