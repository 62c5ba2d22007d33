#!/bin/bash
# tools/sweep_builds.sh -- build tuning variants of the CUDA library (compile-time knobs of kernels.cuh) into
# ambivert_b200/align/tuning/ ; run one with AMBIVERT_B200_LIB=<path> python bench.py ...
#   usage: tools/sweep_builds.sh name "-DABV_UNROLL=4 -DABV_SWITCH_UNROLL=1" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../ambivert_b200/csrc"
mkdir -p ../align/tuning
while [ $# -ge 2 ]; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden $2 -shared -o ../align/tuning/lib_$1.so align_abi.cu &
  shift 2
done
wait
ls -la ../align/tuning/
