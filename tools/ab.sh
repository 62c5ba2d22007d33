#!/bin/bash
# tools/ab.sh -- A/B of tuning builds on the GPU box: for each ambivert_b200/align/tuning/lib_<name>.so given, the reduced-panel
# timing of tools/quick_mm.py in global and glocal mode (one JSON line each, appended to gpurun_out/ab.jsonl).
#   usage (inside gpurun): tools/ab.sh [-q queries] name1 name2 ...
Q=20000
if [ "$1" = "-q" ]; then Q=$2; shift 2; fi
mkdir -p gpurun_out
for n in "$@"; do
  for m in global glocal; do
    AMBIVERT_B200_LIB=ambivert_b200/align/tuning/lib_$n.so python tools/quick_mm.py --queries $Q --mode $m 2>&1 | tail -1 | tee -a gpurun_out/ab.jsonl
  done
done
